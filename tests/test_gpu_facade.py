"""GPU test of the C++ drop-in facade (include/fiesta_b200/ESDFMap.h): a replay written like the reference's own
test/test_ESDF_Map.cpp is compiled against the facade and its output compared with the reference's golden values."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")))


def compile_facade(out):
    import fiesta_b200
    from fiesta_b200 import build as fb_build
    fb_build.build()
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include", "fiesta_b200"), "-I", os.path.join(ROOT, "oracle", "shims"),
           os.path.join(ROOT, "tests", "cpp", "facade_replay.cpp"), "-o", out, "-L", os.path.dirname(fiesta_b200.LIB_PATH),
           "-lfiesta_b200", "-Wl,-rpath," + os.path.dirname(fiesta_b200.LIB_PATH), "/usr/lib/x86_64-linux-gnu/libstdc++.so.6", "-lm"]
    subprocess.check_call(cmd)


def test_facade_compiles(tmp_path):
    """CPU: the facade is header-compatible with the way Fiesta.h uses ESDFMap (compiled against the Eigen/ROS shims)."""
    compile_facade(str(tmp_path / "facade_replay"))


@pytest.mark.gpu
def test_facade_pillar_replay(tmp_path):
    exe = str(tmp_path / "facade_replay")
    compile_facade(exe)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    vals = {l.split()[0]: l.split()[1:] for l in out.stdout.strip().splitlines()}
    g = GOLD["pillar_replay"]
    assert int(vals["grid_total_size_"][0]) == 262144
    assert abs(float(vals["GetDistance(30,30,10)"][0]) - g["GetDistance_30_30_10"]) < 1e-9
    t = g["trilinear_0.33_-1.27_2.51"]
    got = [float(x) for x in vals["trilinear"]]
    assert abs(got[0] - t["dist"]) < 1e-9 and all(abs(a - b) < 1e-9 for a, b in zip(got[1:], t["grad"]))
    assert vals["out_of_map"] == ["-10000", "-10000.0", "-1.0"]
    assert vals["plan_and_mirror_equal"] == ["1"]
    assert int(vals["occupied_points"][0]) == 625
    assert int(vals["slice_points"][0]) == 64 * 64
