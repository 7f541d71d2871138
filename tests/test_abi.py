"""The C-ABI library must load and export every symbol include/h264recon.h declares (no compute without a GPU)."""
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
    names = declared_symbols("h264recon.h")
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/h264recon.h but not exported: {missing}"


def test_synth_library_exports_every_declared_symbol():
    lib = pyapi.synth_lib()
    missing = [n for n in declared_symbols("h264synth.h") if not hasattr(lib, n)]
    assert not missing


def test_struct_layouts_match_the_header():
    assert C.sizeof(pyapi.Mb) == 32
    assert C.sizeof(pyapi.MbMotion) == 192
    assert C.sizeof(pyapi.Slice) == 5216
    assert pyapi.Mb.coeff_offset.offset == 16 and pyapi.Mb.coeff_count.offset == 14 and pyapi.Mb.u.offset == 20 and pyapi.Mb.cbp_blks.offset == 12


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
