import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fiesta_b200
from tests import scenes
m = fiesta_b200.ESDFMap((-1.6, -1.6, 0.0), 0.1, (3.2, 3.2, 3.2))
m.SetParameters(*scenes.PARAMS_TOGGLE)
allv = scenes.all_voxels(m.grid_size)
m.SetOccupancyBatchVox(allv, np.zeros(len(allv), np.uint8))
print("upd occ", m.UpdateOccupancy(True)); m.UpdateESDF(); print(m.stats())
m.SetOccupancyBatchVox(scenes.pillar(12, 12), np.ones(25, np.uint8))
print("upd occ", m.UpdateOccupancy(True)); m.UpdateESDF(); print(m.stats())
D = m.export_distance(); print("finite", ((D >= 0) & (D < 10000)).sum(), D[(D >= 0) & (D < 10000)].sum())
