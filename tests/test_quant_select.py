"""the scaling-list selection (Flat / Default / SPS / PPS lists with fall-back rules A and B) of the GPU
binding (integration/decoder_gpu.cc, Decoder::assign_quant_params) against the reference's own Transform::init
(decoder/transform.cc:173-262), both linked into one host program (tests/quant_select_test.cc, built by
integration/Makefile where /root/reference exists): all 2 x 2 x 256 x 256 present-flag patterns, plus the InvLevelScale
tables h264r_build_level_scale derives from the selected lists against set_quant (:265-302).  No GPU needed."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "integration", "_build", "quant_select_test")


@pytest.mark.skipif(not os.path.exists(BIN), reason="integration/_build/quant_select_test not built (no /root/reference)")
def test_scaling_list_selection_equals_transform_init():
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "66049 flag patterns" in r.stdout
