"""Regression aid for FAST-mode kernel work: replays a seeded LIDAR workload and prints a digest of the resulting
closest-obstacle / distance / occupancy arrays, so two builds of the library can be compared bit for bit.
    python scripts/fast_checksum.py [--workload lidar256] [--frames 12]"""
import argparse, hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, fiesta_b200
from tests import scenes

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="lidar256")
ap.add_argument("--frames", type=int, default=12)
a = ap.parse_args()
w = bench.WORKLOADS[a.workload]
frames = bench.make_frames(a.workload, a.frames)
m = fiesta_b200.ESDFMap(w["origin"], w["res"], w["size"], mode="fast")
m.SetParameters(*scenes.PARAMS_DEFAULT)
for f, (pts, T) in enumerate(frames):
    m.RaycastFrame(pts, T, w["min_len"], w["max_len"])
    if m.CheckUpdate():
        m.SetOriginalRange(); m.UpdateOccupancy(True); m.UpdateESDF()
    s = m.stats()
    h = hashlib.sha1()
    for arr in (m.export_closest_obstacle(), m.export_distance(), m.export_occupancy()):
        h.update(np.ascontiguousarray(arr).tobytes())
    print(f, h.hexdigest()[:16], "changed", s["voxels_changed"], "visits", s["tile_visits"], "gens", s["generations"],
          "rounds", s["raycast_rounds"], "ms", round(s["ms_raycast"], 3), round(s["ms_esdf_wavefront"], 3), flush=True)
