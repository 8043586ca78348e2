"""GPU parity of FAST mode against its CPU model (oracle/fast_model.c): the tile wavefront is schedule independent, so
the kernels have to reproduce the model's closest obstacles and distances bit for bit, ties included -- on scripted
insert/delete replays, on partially observed grids (where FAST may differ from the reference) and on
ray-cast LIDAR frames.  The model is driven with the GPU map's own occupancy state."""
import numpy as np
import pytest

from tests import scenes

pytestmark = pytest.mark.gpu


def logit(p):
    return float(np.log(p / (1.0 - p)))


def esdf_step(dev, model):
    """UpdateOccupancy -> (model update from the device's occupancy state) -> UpdateESDF; returns both tile statistics."""
    if not dev.CheckUpdate():
        return None
    dev.SetOriginalRange()
    dev.UpdateOccupancy(True)
    st = model.update(dev.export_distance(), dev.export_occupancy())
    dev.UpdateESDF()
    return st, dev.stats()


def assert_identical(dev, model, tag):
    cobs, dist = model.export()
    C, D = dev.export_closest_obstacle(), dev.export_distance()
    assert np.array_equal(D, dist), (tag, int((D != dist).sum()))
    assert np.array_equal(C, cobs), (tag, int((C != cobs).any(axis=1).sum()))


@pytest.mark.parametrize("G,res,observed", [(40, 0.1, 1.0), (33, 0.125, 0.6)])
def test_insert_delete_replay_equals_model(oracle_built, G, res, observed):
    import fiesta_b200
    rng = np.random.default_rng(5)
    params = scenes.PARAMS_TOGGLE
    size = ((G - 0.5) * res,) * 3
    dev = fiesta_b200.ESDFMap((-1.0, -2.0, 0.5), res, size, mode="fast")
    dev.SetParameters(*params)
    model = oracle_built.FastModel(dev.grid_size, res, logit(params[4]))
    gs = dev.grid_size
    allv = scenes.all_voxels(gs)
    if observed < 1.0:
        allv = allv[rng.random(len(allv)) < observed]
    dev.SetOccupancyBatchVox(allv, np.zeros(len(allv), np.uint8))
    esdf_step(dev, model)
    for r in range(6):
        n = 600
        vox = np.stack([rng.integers(0, gs[i], n) for i in range(3)], -1).astype(np.int32)
        occ = (rng.random(n) < 0.5).astype(np.uint8)
        dev.SetOccupancyBatchVox(vox, occ)
        st, sd = esdf_step(dev, model)
        assert_identical(dev, model, "round %d" % r)
        assert sd["voxels_changed"] == st["changed"] and sd["generations"] == st["generations"], (st, sd)


def test_lidar_frames_equal_model(oracle_built):
    """Ray-cast frames into a 128^3 map: partially observed space, moving obstacles (inserts and deletes every frame)."""
    import fiesta_b200
    params = scenes.PARAMS_DEFAULT
    origin, res, size = (-3.2, -3.2, -3.2), 0.05, (6.4, 6.4, 6.4)
    dev = fiesta_b200.ESDFMap(origin, res, size, mode="fast")
    dev.SetParameters(*params)
    model = oracle_built.FastModel(dev.grid_size, res, logit(params[4]))
    sc = scenes.Scene((3.0, 3.0, 1.5), 14, 5, seed=21)
    changed = 0
    for f, (p, yaw) in enumerate(scenes.pose_walk(8, seed=22, clamp=1.0)):
        pts, T = scenes.lidar_frame(sc, p, yaw, beams=32, azimuths=360)
        dev.RaycastFrame(pts, T, 0.3, 3.0)
        r = esdf_step(dev, model)
        sc.step()
        if r is None:
            continue
        assert_identical(dev, model, "frame %d" % f)
        assert r[1]["voxels_changed"] == r[0]["changed"] and r[1]["generations"] == r[0]["generations"], r
        changed += r[0]["changed"]
    assert changed > 10000
