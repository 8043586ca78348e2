"""GPU test of the device depth front end: fiesta_depth_frame == restated Fiesta::DepthConversion
(Fiesta.h:319-382) followed by the serial RaycastProcess, with and without the temporal depth filter."""
import numpy as np
import pytest

from tests import scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_filter", [0, 1])
def test_depth_frame_matches_reference_pipeline(oracle_built, use_filter):
    import fiesta_b200
    origin, res, size = (-6.4, -6.4, -3.2), 0.1, (12.8, 12.8, 6.4)
    dev = fiesta_b200.ESDFMap(origin, res, size, mode="exact")
    ora = oracle_built.OracleMap(origin, res, size)
    for m in (dev, ora):
        m.SetParameters(*scenes.PARAMS_DEFAULT)
    scale = 0.25
    dp = fiesta_b200.DepthParams(scenes.FX * scale, scenes.FY * scale, scenes.CX * scale, scenes.CY * scale, use_filter, 2, 10.0, 0.1, 0.1)
    op = oracle_built.DepthParams(scenes.FX * scale, scenes.FY * scale, scenes.CX * scale, scenes.CY * scale, use_filter, 2, 10.0, 0.1, 0.1)
    sc = scenes.Scene((5.0, 5.0, 2.5), 20, 5, seed=2)
    last_img, last_T = None, None
    for f, (p, yaw) in enumerate(scenes.pose_walk(4, seed=3)):
        img, T = scenes.depth_image(sc, p, yaw, width=160, height=120, scale=scale)
        m_rel = np.linalg.inv(last_T) @ T if last_T is not None else np.eye(4)
        n = dev.DepthFrame(img, dp, T, m_rel, 0.5, 5.0)
        cloud = oracle_built.depth_conversion(img, last_img, f + 1, op, m_rel)
        assert n == len(cloud), (f, n, len(cloud))
        assert np.array_equal(dev.last_depth_cloud(), cloud), f
        if len(cloud):
            ora.RaycastFrame(cloud, T, 0.5, 5.0)
        (h1, t1), (h2, t2) = dev.export_counters(), ora.export_counters()
        assert np.array_equal(h1, h2) and np.array_equal(t1, t2), f
        assert dev.CheckUpdate() == ora.CheckUpdate()
        if dev.CheckUpdate():
            assert dev.UpdateOccupancy(True) == ora.UpdateOccupancy(True)
            dev.UpdateESDF(); ora.UpdateESDF()
            assert np.array_equal(dev.export_distance(), ora.export_distance()), f
            assert np.array_equal(dev.export_closest_obstacle(), ora.export_closest_obstacle()), f
        last_img, last_T = img, T
        sc.step()
    assert use_filter == 0 or f > 0
