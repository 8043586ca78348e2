"""CPU: the multiply-high division constants k_x_relax decodes voxel indices with (fiesta_b200/csrc/fb_divmagic.h) are exact for
every divisor a grid can have and every n < 2^31 (checked at both edges of every quotient step, tests/cpp/divmagic_test.cpp)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_division_constants_exact(tmp_path):
    exe = str(tmp_path / "divmagic_test")
    subprocess.check_call(["g++", "-O2", "-fopenmp", os.path.join(ROOT, "tests", "cpp", "divmagic_test.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
