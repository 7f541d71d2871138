"""Host logic: the packed-arithmetic motion-compensation core of recon_inter_kernel
(arrow-h264_b200/csrc/mc_core.cuh) compiled for the CPU with the PTX primitives emulated and checked exhaustively
(every fractional position, window alignment, weighting mode, extreme sample patterns) against a per-sample
restatement of the reference's formulas (tests/mc_core_host_test.cc)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mc_core_matches_reference_formulas(tmp_path):
    out = str(tmp_path / "mc_core_test")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wno-unknown-pragmas", "-x", "c++",
                           "-I" + os.path.join(ROOT, "arrow-h264_b200", "csrc"),
                           os.path.join(ROOT, "tests", "mc_core_host_test.cc"), "-o", out])
    r = subprocess.run([out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
