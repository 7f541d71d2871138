"""Property test of the packed motion format (include/h264recon.h: h264r_pack_motion / h264r_unpack_motion): whatever the
sixteen per-block entries of a macroblock look like, packing and unpacking gives them back, the layout code is the smallest
one that can hold them, and the entries written are exactly the distinct blocks of that layout."""
import ctypes as C

from hypothesis import given, settings, strategies as st

import pyapi

ENTRIES = [0, 1, 2, 2, 4, 16]            # h264r_motion_entries_of_code


class Entry(C.Structure):
    _fields_ = [("mv", C.c_int16 * 2 * 2), ("ref_idx", C.c_int8 * 2), ("ref_pic", C.c_int8 * 2)]


def lib():
    L = pyapi.synth_lib()
    L.h264r_pack_motion.restype = C.c_int
    L.h264r_pack_motion.argtypes = [C.c_void_p, C.c_void_p]
    L.h264r_unpack_motion.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    return L


entry = st.tuples(st.integers(-2048, 2047), st.integers(-512, 511), st.integers(-2048, 2047), st.integers(-512, 511),
                  st.integers(-1, 3), st.integers(-1, 3))
# a small pool of entries and a layout that decides which blocks share one: uniform, halves, quadrants, free
layout = st.sampled_from(["mb", "rows", "cols", "quads", "free"])


@settings(max_examples=400, deadline=None)
@given(pool=st.lists(entry, min_size=1, max_size=16), how=layout, pick=st.lists(st.integers(0, 15), min_size=16, max_size=16))
def test_pack_unpack_round_trip_and_minimal_code(pool, how, pick):
    L = lib()
    group = {"mb": lambda b: 0, "rows": lambda b: b >> 3, "cols": lambda b: (b >> 1) & 1,
             "quads": lambda b: (b >> 3) * 2 + ((b >> 1) & 1), "free": lambda b: b}[how]
    m = pyapi.MbMotion()
    blocks = []
    for b in range(16):
        e = pool[pick[group(b)] % len(pool)]
        blocks.append(e)
        for l in range(2):
            m.mv[l][b][0], m.mv[l][b][1] = e[2 * l], e[2 * l + 1]
            m.ref_idx[l][b] = e[4 + l]
            m.ref_pic[l][b] = e[4 + l]
    out = (Entry * 16)()
    code = L.h264r_pack_motion(C.byref(m), out)
    assert 1 <= code <= 5
    # the smallest layout whose blocks are uniform inside every group
    same = lambda f: all(blocks[b] == blocks[c] for b in range(16) for c in range(16) if f(b) == f(c))
    want = 1 if same(lambda b: 0) else 2 if same(lambda b: b >> 3) else 3 if same(lambda b: (b >> 1) & 1) \
        else 4 if same(lambda b: (b >> 3) * 2 + ((b >> 1) & 1)) else 5
    assert code == want
    stream = (C.c_uint32 * (3 * 16 + 8))()
    C.memmove(C.byref(stream, 4 * 5), out, 12 * ENTRIES[code])          # the entries at word 5 of a stream
    back = pyapi.MbMotion()
    L.h264r_unpack_motion(stream, (5 << 4) | code, C.byref(back))
    assert bytes(back) == bytes(m)
