"""Host logic: the Decoder facade (arrow-h264_b200/csrc/decoder_facade.{h,cc}) mirrors the reference's
`vio::h264::Decoder` entry points.  tests/facade_roundtrip.cc feeds it the parser's call sequence and requires the
buffers it fills to equal the generator's picture description (inverse scans, cbp_blks, header packing)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def facade_binary(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("facade") / "facade_rt")
    csrc = os.path.join(ROOT, "arrow-h264_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + csrc,
                           os.path.join(ROOT, "tests", "facade_roundtrip.cc"), os.path.join(csrc, "decoder_facade.cc"),
                           os.path.join(csrc, "synth.cc"), os.path.join(csrc, "host_helpers.cc"), "-o", out])
    return out


@pytest.mark.parametrize("cfg,w,h,n", [(1, 22, 18, 4), (2, 9, 7, 9), (3, 12, 8, 6), (4, 7, 5, 8), (6, 12, 8, 6)])
def test_facade_reproduces_the_picture_description(facade_binary, cfg, w, h, n):
    r = subprocess.run([facade_binary, str(cfg), str(w), str(h), str(n)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
