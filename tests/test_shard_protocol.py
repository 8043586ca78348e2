"""CPU test (gloo, world_size 2) of the multi-GPU exchange protocol in fiesta_b200/shard.py.

The GPU map is replaced by a mock with the same four calls that relaxes a 1-D distance field with reach 2 inside its slab;
the sharded result gathered over the ranks must equal the unsharded relaxation, and the loop must terminate on all ranks
together.  This covers the host-side logic of the N>1 path: who sends which layers to whom, ghost ingestion, termination.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from fiesta_b200 import shard

N, INF = 96, 10**6
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
seeds = [5, 40, 41, 90]

def relax_full():
    d = np.full(N, INF, np.int64); d[seeds] = 0
    ch = True
    while ch:
        ch = False
        for i in range(N):
            for s in (-2, -1, 1, 2):
                j = i + s
                if 0 <= j < N and d[j] + abs(s) < d[i]: d[i] = d[j] + abs(s); ch = True
    return d

class Mock:
    """1-D stand-in for the sharded map: owns [x0, x1), keeps 2 ghost cells per internal face in a full-size array."""
    def __init__(self):
        self.x0, self.x1 = N * rank // world, N * (rank + 1) // world
        self.d = np.full(N, INF, np.int64)
        self.bufs = None
    def _relax(self):
        n = 0; ch = True
        while ch:
            ch = False
            for i in range(self.x0, self.x1):
                for s in (-2, -1, 1, 2):
                    j = i + s
                    if 0 <= j < N and (self.x0 - 2 <= j < self.x1 + 2) and self.d[j] + abs(s) < self.d[i]:
                        self.d[i] = self.d[j] + abs(s); ch = True; n += 1
        return n
    def UpdateESDF(self):
        for s in seeds: self.d[s] = 0          # replicated inputs
        self._relax()
    def shard_pack(self, lo, hi):
        if not isinstance(lo, int): lo[:2] = torch.from_numpy(self.d[self.x0:self.x0 + 2].astype(np.int32))
        if not isinstance(hi, int): hi[:2] = torch.from_numpy(self.d[self.x1 - 2:self.x1].astype(np.int32))
    def shard_ingest(self, lo, hi):
        n = 0
        if not isinstance(lo, int):
            new = lo[:2].numpy().astype(np.int64); n += int((new != self.d[self.x0 - 2:self.x0]).sum()); self.d[self.x0 - 2:self.x0] = new
        if not isinstance(hi, int):
            new = hi[:2].numpy().astype(np.int64); n += int((new != self.d[self.x1:self.x1 + 2]).sum()); self.d[self.x1:self.x1 + 2] = new
        return n
    def shard_relax(self):
        return self._relax()

m = Mock()
bufs = shard.HaloBuffers(2, "cpu")
rounds = shard.sharded_update_esdf(m, bufs, rank, world, ptr=lambda t: t)
# `ptr` hands the tensors themselves to the mock; rank-edge arguments arrive as 0 -> None
own = torch.from_numpy(m.d[m.x0:m.x1].copy())
parts = [torch.zeros(N * (r + 1) // world - N * r // world, dtype=torch.int64) for r in range(world)]
dist.all_gather(parts, own)
full = torch.cat(parts).numpy()
assert np.array_equal(full, relax_full()), (rank, full, relax_full())
assert rounds >= 2
print("rank", rank, "ok rounds", rounds)
dist.destroy_process_group()
'''


def test_two_rank_halo_protocol(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok rounds" in o, o
