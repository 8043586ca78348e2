"""Offline study with the CPU model of the FAST-mode UpdateESDF (oracle/fast_model.c): how many tile visits does the exit
test (a visiting tile checks exactly whether its changes can improve a neighbour's border before queueing it) save on the
LIDAR workload, and does it -- or pulling from every queued voxel -- change any record?  CPU only.
    python tests/tools/fast_model_study.py [--workload lidar256] [--frames 12]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
from oracle import pyoracle
from tests import scenes

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="lidar256")
ap.add_argument("--frames", type=int, default=12)
a = ap.parse_args()
w = bench.WORKLOADS[a.workload]
frames = bench.make_frames(a.workload, a.frames)
ora = pyoracle.OracleMap(w["origin"], w["res"], w["size"])
ora.SetParameters(*scenes.PARAMS_DEFAULT)
l_occ = float(np.log(scenes.PARAMS_DEFAULT[4] / (1 - scenes.PARAMS_DEFAULT[4])))
base, full, ex = (pyoracle.FastModel(ora.grid_size, ora.resolution, l_occ) for _ in range(3))
t16, t16x = (pyoracle.FastModel(ora.grid_size, ora.resolution, l_occ, tile=16) for _ in range(2))
tot = dict(base=0, exit=0, act=0, sup=0)
agg = {k: dict(generations=0, full_visits=0, iterations=0, evaluations=0, candidates=0, diff=0) for k in ("8", "8x", "16", "16x")}
for f, (pts, T) in enumerate(frames):
    ora.RaycastFrame(pts, T, w["min_len"], w["max_len"])
    if not ora.CheckUpdate():
        continue
    ora.SetOriginalRange(); ora.UpdateOccupancy(True)
    dist, occ = ora.export_distance(), ora.export_occupancy()
    s0 = base.update(dist, occ, 0)
    s1 = full.update(dist, occ, pyoracle.FastModel.FULL_PULL)
    s2 = ex.update(dist, occ, pyoracle.FastModel.EXIT_TEST)
    s3 = t16.update(dist, occ, 0)
    s4 = t16x.update(dist, occ, pyoracle.FastModel.EXIT_TEST)
    ora.UpdateESDF()
    c0, d0 = base.export(); c1, d1 = full.export(); c2, d2 = ex.export()
    same = np.array_equal(c0, c1) and np.array_equal(c0, c2)
    R = ora.export_distance()
    fin = (R >= 0) & (R < 10000)
    print("frame %2d  N_ref %8d  changed %8d  gens %3d->%3d  full visits %7d -> %7d (%.0f %%)  activations %7d suppressed %7d  variants identical %s  dist != reference %.3f %% of %d"
          % (f, ora.stats()["expansions"], s0["changed"], s0["generations"], s2["generations"], s0["full_visits"], s2["full_visits"],
             100.0 * s2["full_visits"] / max(1, s0["full_visits"]), s2["activations"], s2["suppressed"], same,
             100.0 * (d0 != R)[fin].mean() if fin.any() else 0.0, int(fin.sum())), flush=True)
    assert same and s0["full_visits"] == s1["full_visits"]
    tot["base"] += s0["full_visits"]; tot["exit"] += s2["full_visits"]; tot["act"] += s2["activations"]; tot["sup"] += s2["suppressed"]
    c3, d3 = t16.export(); c4, d4 = t16x.export()
    for k, st, dd in (("8", s0, d0), ("8x", s2, d2), ("16", s3, d3), ("16x", s4, d4)):
        for q in ("generations", "full_visits", "iterations", "evaluations", "candidates"):
            agg[k][q] += st[q]
        agg[k]["diff"] += int((dd != d0).sum())
print("total full visits %d -> %d (%.1f %%), %d of %d neighbour activations suppressed" %
      (tot["base"], tot["exit"], 100.0 * tot["exit"] / max(1, tot["base"]), tot["sup"], tot["act"]))
print("tile / exit test | generations | full visits | local iterations | voxel evaluations | candidates | distances != 8^3 result")
for k in ("8", "8x", "16", "16x"):
    g = agg[k]
    print("%4s | %6d | %8d | %8d | %10d | %11d | %d" % (k, g["generations"], g["full_visits"], g["iterations"], g["evaluations"], g["candidates"], g["diff"]))
