"""CPU tests of oracle/fast_model.c, the sequential model of the product's FAST-mode UpdateESDF.

The model is driven with the occupancy state of the reference build (oracle/_ref, else the C port) and compared with that
oracle's UpdateESDF result; its two study switches (every queued voxel pulls / exit test before queueing a neighbour)
must not change a single record.  The GPU kernels are compared with the same model in tests/test_gpu_fast_model.py.
"""
import numpy as np
import pytest

from tests import scenes


def logit(p):
    return float(np.log(p / (1.0 - p)))


def model_for(pyoracle, ora, params):
    return pyoracle.FastModel(ora.grid_size, ora.resolution, logit(params[4]))


def step(ora, models, vox, occ):
    """One SetOccupancy batch -> UpdateOccupancy -> UpdateESDF on the oracle, mirrored into every (model, flags) pair."""
    ora.SetOccupancyBatchVox(vox, occ)
    if not ora.CheckUpdate():
        return []
    ora.UpdateOccupancy(True)
    dist, occs = ora.export_distance(), ora.export_occupancy()    # after UpdateOccupancy, before UpdateESDF
    stats = [m.update(dist, occs, flags) for m, flags in models]
    ora.UpdateESDF()
    for m, _ in models:
        assert m.fresh_left() == 0
    return stats


def check_against_oracle(model, ora, tag):
    cobs, dist = model.export()
    R, S = ora.export_distance(), ora.export_closest_obstacle()
    dm = dist != R
    cm = (cobs != S).any(axis=1)
    assert dm.sum() == 0, (tag, int(dm.sum()))
    assert (cm & dm).sum() == 0, tag                      # a different obstacle is only ever an exact distance tie
    return int((cm & ~dm).sum())


def same(a, b):
    ca, da = a.export()
    cb, db = b.export()
    return np.array_equal(ca, cb) and np.array_equal(da, db)


def test_pillar_replay_model_equals_reference(oracle_built):
    """observe all, 25 pillars, delete in reverse; distances equal the reference bit for bit."""
    params = scenes.PARAMS_TOGGLE
    ora = oracle_built.OracleMap((-6.4, -6.4, 0.0), 0.2, (12.8, 12.8, 12.8))
    ora.SetParameters(*params)
    base, full, exit_ = (model_for(oracle_built, ora, params) for _ in range(3))
    models = [(base, 0), (full, oracle_built.FastModel.FULL_PULL), (exit_, oracle_built.FastModel.EXIT_TEST)]
    allv = scenes.all_voxels(ora.grid_size)
    step(ora, models, allv, np.zeros(len(allv), np.uint8))
    sites = scenes.pillar_sites()
    seq = [(x, y, 1) for x, y in sites] + [(x, y, 0) for x, y in reversed(sites)]
    saved = 0
    for k, (x, y, o) in enumerate(seq):
        st = step(ora, models, scenes.pillar(x, y), np.full(25, o, np.uint8))
        check_against_oracle(base, ora, "step %d" % k)
        assert same(base, full) and same(base, exit_), k
        assert st[0]["full_visits"] == st[1]["full_visits"]          # restricting the pull does not change the schedule
        assert st[2]["full_visits"] <= st[0]["full_visits"]
        saved += st[0]["full_visits"] - st[2]["full_visits"]
    _, dist = base.export()
    assert ((dist >= 0) & (dist < 10000)).sum() == 0                   # everything deleted again
    assert saved > 0                                                   # the exit test does prune echo visits


@pytest.mark.parametrize("G,res,observed", [(40, 0.1, 1.0), (33, 0.125, 1.0), (40, 0.1, 0.6)])
def test_random_insert_delete(oracle_built, G, res, observed):
    """Mixed insert/delete rounds; fully observed grids must match the reference's distances exactly, partially observed
    ones (unknown voxels are barriers, results become order dependent) only have to agree between the model variants."""
    rng = np.random.default_rng(11)
    params = scenes.PARAMS_TOGGLE
    size = ((G - 0.5) * res,) * 3
    ora = oracle_built.OracleMap((-1.0, -2.0, 0.5), res, size)
    ora.SetParameters(*params)
    base, full, exit_ = (model_for(oracle_built, ora, params) for _ in range(3))
    models = [(base, 0), (full, oracle_built.FastModel.FULL_PULL), (exit_, oracle_built.FastModel.EXIT_TEST)]
    gs = ora.grid_size
    allv = scenes.all_voxels(gs)
    if observed < 1.0:
        allv = allv[rng.random(len(allv)) < observed]
    step(ora, models, allv, np.zeros(len(allv), np.uint8))
    for r in range(6):
        n = 500
        vox = np.stack([rng.integers(0, gs[i], n) for i in range(3)], -1).astype(np.int32)
        occ = (rng.random(n) < 0.5).astype(np.uint8)
        step(ora, models, vox, occ)
        if observed == 1.0:
            check_against_oracle(base, ora, "round %d" % r)
        assert same(base, full) and same(base, exit_), r
