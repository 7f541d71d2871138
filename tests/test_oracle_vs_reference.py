"""Pins the CPU restatement against the reference's own Decoder, sample for sample, where the reference could be
compiled (oracle/_ref/libh264ref.so exists: the build container).  Skipped elsewhere -- the golden digests cover it."""
import os

import pytest

import oracle_py as O
import pyapi

needs_ref = pytest.mark.skipif(not os.path.exists(O.REF_PATH), reason="oracle/_ref not built (no /root/reference)")

CASES = [(1, 2, 0, 0, 5), (2, 4, 24, 13, 9), (3, 8, 20, 11, 7), (3, 9, 20, 11, 7), (4, 2, 21, 12, 8), (5, 31, 15, 9, 7),
         (6, 0, 20, 11, 9), (6, 1, 13, 6, 12), (6, 2, 20, 11, 9)]       # field pictures (PAFF)


@needs_ref
@pytest.mark.parametrize("cfg,sidx,w,h,n", CASES)
def test_port_equals_reference(cfg, sidx, w, h, n):
    st = pyapi.SynthStream(cfg, sidx, w, h, n)
    seq = st.seq
    st.close()
    ref, port = O.CpuDecoder("ref", seq), O.CpuDecoder("port", seq)
    ref_planes, port_planes = [], []
    O.run_stream(ref, cfg, sidx, w, h, n, on_picture=lambda p, pl: ref_planes.append(pl))
    O.run_stream(port, cfg, sidx, w, h, n, on_picture=lambda p, pl: port_planes.append(pl))
    ref.close()
    port.close()
    assert len(ref_planes) == len(port_planes) == n
    for i, (a, b) in enumerate(zip(ref_planes, port_planes)):
        for name, pa, pb in zip("Y Cb Cr".split(), a, b):
            if pa != pb:
                first = next(k for k in range(len(pa)) if pa[k] != pb[k])
                pytest.fail(f"picture {i} plane {name}: first difference at sample {first} (ref {pa[first]}, port {pb[first]})")


@needs_ref
def test_mutation_is_detected():
    """The comparison must be able to fail: switching the deblocking pass off in the port changes every picture."""
    cfg, sidx, w, h, n = 3, 0, 12, 8, 4
    st = pyapi.SynthStream(cfg, sidx, w, h, n)
    seq = st.seq
    st.close()

    class NoDeblock(O.CpuDecoder):
        def reconstruct(self, dst, pic, refs):
            pic.pp.run_deblock = 0
            return super().reconstruct(dst, pic, refs)

    a = O.run_stream(O.CpuDecoder("ref", seq), cfg, sidx, w, h, n)
    b = O.run_stream(NoDeblock("port", seq), cfg, sidx, w, h, n)
    assert all(x != y for x, y in zip(a, b))
