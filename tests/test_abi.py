"""The C-ABI library must load and export every symbol include/h264recon.h (the drop-in boundary) and
include/h264recon_bench.h (measurement hooks) declare (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

import pyapi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "arrow-h264_b200", "libh264recon.so")


def declared_symbols(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(h264[rs]_[a-z0-9_]+)\s*\(", src)))


@pytest.mark.skipif(not os.path.exists(LIB), reason="libh264recon.so not built yet (run __graft_entry__.build())")
def test_recon_library_exports_every_declared_symbol():
    lib = C.CDLL(LIB)
    for header, at_least in (("h264recon.h", 20), ("h264recon_bench.h", 4)):
        names = declared_symbols(header)
        assert len(names) >= at_least
        missing = [n for n in names if not hasattr(lib, n)]
        assert not missing, f"declared in include/{header} but not exported: {missing}"
    # the measurement hooks stay out of the boundary header
    assert not [n for n in declared_symbols("h264recon.h") if "bench" in n or "replay" in n]


def test_synth_library_exports_every_declared_symbol():
    lib = pyapi.synth_lib()
    missing = [n for n in declared_symbols("h264synth.h") if not hasattr(lib, n)]
    assert not missing


def test_struct_layouts_match_the_header():
    assert C.sizeof(pyapi.Mb) == 32
    assert C.sizeof(pyapi.MbMotion) == 192
    assert C.sizeof(pyapi.Slice) == 5216
    assert pyapi.Mb.coeff_offset.offset == 16 and pyapi.Mb.coeff_count.offset == 14 and pyapi.Mb.u.offset == 20 and pyapi.Mb.cbp_blks.offset == 12
    assert pyapi.Mb.motion.offset == 28
    assert C.sizeof(pyapi.PicParams) == 8 + 4 * 32 + 8 + 4 * 32 + 32 + 4 + 4 + 32 and pyapi.PicParams.direct_8x8_inference_flag.offset == 304
    assert pyapi.PicParams.structure.offset == 308 and pyapi.PicParams.ref_structure.offset == 312
    assert C.sizeof(pyapi.PicBuffers) == 32 and pyapi.PicBuffers.picture.offset == 28


def test_pack_picture_round_trip():
    """h264r_pack_picture (what h264r_picture_fill and the parser-side facade do): the packed motion of every inter MB,
    expanded again with h264r_unpack_motion, equals the generator's sixteen entries -- on streams with one, two, four and
    sixteen distinct entries per MB (stream 2: direct_8x8_inference_flag = 0)."""
    lib = pyapi.synth_lib()
    lib.h264r_pack_picture.restype = C.c_int64
    lib.h264r_pack_picture.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
    lib.h264r_unpack_motion.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    codes = set()
    for sidx in (0, 2):
        st = pyapi.SynthStream(3, sidx, 9, 6, 6)
        for pic in st:
            n = pic.nmb
            cap = pic.info.num_levels + 48 * n
            out_mbs = (pyapi.Mb * n)()
            stream = (C.c_uint32 * cap)()
            words = lib.h264r_pack_picture(n, pic.mbs, pic.motion, pic.levels, pic.info.num_levels, out_mbs, stream, cap)
            assert pic.info.num_levels <= words <= cap
            assert bytes(stream)[:4 * pic.info.num_levels] == bytes(pic.levels)[:4 * pic.info.num_levels]
            for i in range(n):
                if pic.mbs[i].flags & 1:
                    assert out_mbs[i].motion == 0
                    continue
                codes.add(out_mbs[i].motion & 15)
                assert (out_mbs[i].motion >> 4) + 3 <= words
                back = pyapi.MbMotion()
                lib.h264r_unpack_motion(stream, out_mbs[i].motion, C.byref(back))
                assert bytes(back) == bytes(pic.motion[i]), f"stream {sidx} MB {i}"
            assert lib.h264r_pack_picture(n, pic.mbs, pic.motion, pic.levels, pic.info.num_levels, out_mbs, stream, pic.info.num_levels) < 0 or \
                all(m.flags & 1 for m in pic.mbs)
        st.close()
    assert codes == {1, 2, 3, 4, 5}


@pytest.mark.skipif(not os.path.exists(LIB), reason="libh264recon.so not built yet")
def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU the engine must refuse to create a context (there is no CPU path)."""
    lib = pyapi.recon_lib()
    if lib.h264r_device_count() > 0:
        pytest.skip("a CUDA device is present")
    st = pyapi.SynthStream(1, 0, 2, 2, 1)
    with pytest.raises(pyapi.EngineError):
        pyapi.Engine(st.seq)
    st.close()


def test_host_helpers_implicit_weights():
    lib = pyapi.synth_lib()
    w0, w1 = C.c_int(), C.c_int()
    lib.h264r_implicit_weights(4, 0, 6, 0, 0, C.byref(w0), C.byref(w1))      # tb=4, td=6: tx=2731, DistScaleFactor=171 -> w1 = 42, w0 = 22
    assert (w0.value, w1.value) == (22, 42)
    lib.h264r_implicit_weights(4, 0, 0, 0, 0, C.byref(w0), C.byref(w1))      # td == 0
    assert (w0.value, w1.value) == (32, 32)
    lib.h264r_implicit_weights(4, 0, 6, 1, 0, C.byref(w0), C.byref(w1))      # long-term
    assert (w0.value, w1.value) == (32, 32)
