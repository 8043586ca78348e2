"""Run under torchrun on >= 2 GPUs (tests/test_gpu_shard.py launches it): the x-slab sharded FAST UpdateESDF, ghost layers
exchanged with NCCL, against (a) the CPU oracle and (b) the unsharded FAST map on the same inputs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fiesta_b200  # noqa: E402
from fiesta_b200 import shard  # noqa: E402
from oracle import pyoracle  # noqa: E402
from tests import scenes  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
origin, res, size = (-3.2, -3.2, -1.6), 0.1, (6.4, 6.4, 3.2)
m = fiesta_b200.ESDFMap(origin, res, size, device=local, mode="fast")
m.SetParameters(*scenes.PARAMS_TOGGLE)
info = m.set_shard(rank, world)
bufs = shard.HaloBuffers(info.layer_words, torch.device("cuda", local))
ref = fiesta_b200.ESDFMap(origin, res, size, device=local, mode="fast")      # unsharded FAST map on the same GPU
ref.SetParameters(*scenes.PARAMS_TOGGLE)
ora = pyoracle.OracleMap(origin, res, size) if rank == 0 else None
if ora:
    ora.SetParameters(*scenes.PARAMS_TOGGLE)
rng = np.random.default_rng(3)
gs = m.grid_size
G = int(np.prod(gs))


def gather_distance():
    """Every rank contributes the x-layers it owns."""
    D = torch.from_numpy(m.export_distance().reshape(gs)).cuda()
    full = torch.zeros_like(D)
    full[info.x_begin:info.x_end] = D[info.x_begin:info.x_end]
    dist.all_reduce(full)
    return full.cpu().numpy().reshape(-1)


allv = scenes.all_voxels(gs)
rounds_total = 0
for r in range(5):
    if r == 0:
        vox, occ = allv, np.zeros(len(allv), np.uint8)
    else:
        n = 400
        vox = np.stack([rng.integers(0, gs[i], n) for i in range(3)], -1).astype(np.int32)
        occ = (rng.random(n) < 0.5).astype(np.uint8)
    for mm in (m, ref) + ((ora,) if ora else ()):
        mm.SetOccupancyBatchVox(vox, occ)
        mm.UpdateOccupancy(True)
    ref.UpdateESDF()
    if ora:
        ora.UpdateESDF()
    rounds_total += shard.sharded_update_esdf(m, bufs, rank, world)
    D = gather_distance()
    Dref = ref.export_distance()
    assert np.array_equal(D, Dref), ("sharded vs single-GPU", r, int((D != Dref).sum()))
    if ora:
        Do = ora.export_distance()
        assert np.array_equal(D, Do), ("sharded vs oracle", r, int((D != Do).sum()))
if rank == 0:
    print("SHARD_OK world", world, "exchange rounds", rounds_total)
dist.destroy_process_group()
