"""The exact "%g" / integer formatter of the device-side output writer (repkiller_b200/csrc/rk_fmt.cuh, K6) is plain
integer C++: compile it for the host and compare it with printf on random bit patterns, percentages, every binade,
the neighbours of every power of ten and rounding ties (tests/native/fmt_check.cpp)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_g6_formatter_matches_printf(tmp_path):
    exe = tmp_path / "fmt_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "native", "fmt_check.cpp")])
    p = subprocess.run([str(exe), "400000"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-2000:]
    assert "0 mismatches" in p.stdout
