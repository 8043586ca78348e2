"""CPU unit test of the exit test's bit logic (fiesta_b200/csrc/fb_exit_test.h, used by k_wavefront when built with
-DWF_EXIT_TEST=1): the row / neighbour masks and the improvement check are compared with a brute-force restatement on random
boxes (tests/cpp/exit_test_logic.cpp, compiled for the host with g++)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exit_test_bit_logic(tmp_path):
    exe = str(tmp_path / "exit_test_logic")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "fiesta_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cpp", "exit_test_logic.cpp"), "-o", exe])
    out = subprocess.run([exe, "300"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    tag, trials, checked, improving = out.stdout.split()
    assert tag == "OK" and int(checked) > 100000 and int(improving) > 1000, out.stdout
