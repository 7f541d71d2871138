"""Test-side bindings of the oracles (oracle/liboracle_port.so = CPU restatement, oracle/_ref/libh264ref.so =
the reference's own Decoder).  Test infrastructure only."""
import ctypes as C
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "arrow-h264_b200"))
import pyapi  # noqa: E402

PORT_PATH = os.path.join(ROOT, "oracle", "liboracle_port.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libh264ref.so")


class CpuDecoder:
    """Uniform wrapper over port_* / ref_* (same C shape)."""

    def __init__(self, kind, seq):
        self.kind = kind
        path, pre = (PORT_PATH, "port_") if kind == "port" else (REF_PATH, "ref_")
        L = C.CDLL(path)
        self.L = L
        g = lambda n: getattr(L, pre + n)
        self._open, self._close = g("open"), g("close")
        self._alloc, self._release = g("frame_alloc"), g("frame_release")
        self._get, self._set, self._recon = g("frame_get"), g("frame_set"), g("reconstruct")
        self._open.restype = C.c_void_p
        self._open.argtypes = [C.POINTER(pyapi.SeqParams)]
        self._close.argtypes = [C.c_void_p]
        self._alloc.argtypes = [C.c_void_p]
        self._release.argtypes = [C.c_void_p, C.c_int]
        self._get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self._set.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self._recon.restype = C.c_int
        self._recon.argtypes = [C.c_void_p, C.c_int, C.POINTER(pyapi.PicParams), C.c_int, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        self.seq = seq
        self.h = self._open(C.byref(seq))
        self.ny = seq.width_mbs * 16 * seq.height_mbs * 16
        self.sec_decode = 0.0
        self.sec_deblock = 0.0

    def frame_alloc(self):
        return self._alloc(self.h)

    def frame_release(self, f):
        self._release(self.h, f)

    def reconstruct(self, dst, pic, ref_frames):
        pp = pyapi.PicParams.from_buffer_copy(pic.pp)
        for i, f in enumerate(ref_frames):
            pp.ref_frames[i] = f
        t0, t1 = C.c_double(), C.c_double()
        rc = self._recon(self.h, dst, C.byref(pp), pic.info.used_for_reference, pic.slices, pic.mbs, pic.motion,
                         pic.levels, C.byref(t0), C.byref(t1))
        self.sec_decode += t0.value
        self.sec_deblock += t1.value
        return rc

    def frame_get(self, f):
        y = (C.c_uint8 * self.ny)()
        cb = (C.c_uint8 * (self.ny // 4))()
        cr = (C.c_uint8 * (self.ny // 4))()
        self._get(self.h, f, y, cb, cr)
        return bytes(y), bytes(cb), bytes(cr)

    def close(self):
        if self.h:
            self._close(self.h)
            self.h = None


def run_stream(dec, config, stream_idx=0, width_mbs=0, height_mbs=0, num_frames=0, on_picture=None):
    """Decode one synthetic stream with a CpuDecoder; returns the list of per-picture md5 digests (Y|Cb|Cr)."""
    st = pyapi.SynthStream(config, stream_idx, width_mbs, height_mbs, num_frames)
    frames = {}
    digests = []
    for pic in st:
        dst = dec.frame_alloc()
        frames[pic.info.pic_index] = dst
        refs = [frames[pic.info.ref_pic_index[i]] for i in range(pic.info.num_refs)]
        rc = dec.reconstruct(dst, pic, refs)
        assert rc == 0, f"{dec.kind}: cbp_blks mismatch on {rc} MBs (generator/facade inconsistency)"
        planes = dec.frame_get(dst)
        digests.append(hashlib.md5(b"".join(planes)).hexdigest())
        if on_picture:
            on_picture(pic, planes)
        for i in range(pic.info.num_refs):
            if pic.info.last_use_of_ref[i]:
                dec.frame_release(frames.pop(pic.info.ref_pic_index[i]))
        if not pic.info.used_for_reference:
            dec.frame_release(frames.pop(pic.info.pic_index))
    st.close()
    return digests
