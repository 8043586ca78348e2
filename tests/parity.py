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
    both = dm & (D >= 0) & (D < 10000) & (R >= 0) & (R < 10000)                 # mismatches where both sides hold a real distance
    res["dist_max_err"] = float(np.abs(D - R)[both].max()) if both.any() else 0.0
    res["reach"] = int((dm & ~both).sum())                                      # one side unreached / unknown, the other not
    cm = (C != S).any(axis=1)
    res["cobs_tie"] = int((cm & ~dm).sum())
    res["cobs_nontie"] = int((cm & dm).sum())
    res["occ"] = int((O != P).sum())
    res["finite"] = int(((R >= 0) & (R < 10000)).sum())
    if check_counters:
        (h1, t1), (h2, t2) = dev.export_counters(), ora.export_counters()
        res["counters"] = int(((h1 != h2) | (t1 != t2)).sum())
    return res


DIRS24 = np.array([(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1), (-1, -1, 0), (1, 1, 0), (0, -1, -1), (0, 1, 1),
                   (-1, 0, -1), (1, 0, 1), (-1, 1, 0), (1, -1, 0), (0, -1, 1), (0, 1, -1), (1, 0, -1), (-1, 0, 1), (-2, 0, 0), (2, 0, 0),
                   (0, -2, 0), (0, 2, 0), (0, 0, -2), (0, 0, 2)])       # parameters.h:55-68


def invariants(m, l_occ):
    """Structural checks of one map on its own (no oracle), for results that are not bit-identical to the reference: for
    every voxel with a finite distance (a) its closest obstacle is occupied, (b) its distance is the distance to that
    obstacle (ESDFMap.cpp:122-123), (c) no occupied voxel of its 24-neighbourhood is closer than that."""
    gs = m.grid_size
    D = m.export_distance().reshape(gs)
    C = m.export_closest_obstacle().reshape(gs + (3,))
    occ = m.export_occupancy().reshape(gs) > l_occ
    fin = (D >= 0) & (D < 10000)
    idx = np.stack(np.meshgrid(*[np.arange(g) for g in gs], indexing="ij"), -1)
    cf = C[fin]
    res = {}
    res["obstacle_not_occupied"] = int((~occ[cf[:, 0], cf[:, 1], cf[:, 2]]).sum())
    d = np.sqrt(((cf - idx[fin]) ** 2).sum(axis=1).astype(np.float64)) * m.resolution
    res["distance_not_to_obstacle"] = int((d != D[fin]).sum())
    bad = np.zeros(gs, bool)
    for dx, dy, dz in DIRS24:
        sh = np.zeros(gs, bool)                                  # sh[v] = occupied(v + dir), False outside the grid
        xs, ys, zs = [slice(max(0, -o), g - max(0, o)) for o, g in zip((dx, dy, dz), gs)]
        xd, yd, zd = [slice(max(0, o), g - max(0, -o)) for o, g in zip((dx, dy, dz), gs)]
        sh[xs, ys, zs] = occ[xd, yd, zd]
        bad |= fin & sh & (D > np.sqrt(float(dx * dx + dy * dy + dz * dz)) * m.resolution + 1e-9)
    res["closer_occupied_neighbour"] = int(bad.sum())
    return res


def assert_exact_distance(dev, ora, tag="", tie_frac_limit=0.35):
    r = compare(dev, ora)
    assert r["occ"] == 0, (tag, r)
    assert r["dist"] == 0, (tag, r)
    assert r["cobs_nontie"] == 0, (tag, r)
    assert r["cobs_tie"] <= tie_frac_limit * max(1, r["finite"]), (tag, r)
    return r
