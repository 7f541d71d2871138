"""ctypes bindings of the C ABI (include/h264recon.h, include/h264recon_bench.h, include/h264synth.h).

Python is only the test/bench harness language here; the product is the C-ABI shared library
`libh264recon.so` (CUDA engine).  Nothing in this file computes a sample, and nothing here touches oracle/.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_REFS = 32
COEFFS_PER_MB = 384
P_SLICE, B_SLICE, I_SLICE = 0, 1, 2


class InterInfo(C.Structure):
    _fields_ = [("sub_mb_type", C.c_uint8 * 4), ("sub_mb_pred_mode", C.c_uint8 * 4)]


class MbUnion(C.Union):
    _fields_ = [("intra_modes", C.c_uint8 * 8), ("inter", InterInfo)]


class Mb(C.Structure):
    _fields_ = [("mb_type", C.c_uint8), ("flags", C.c_uint8), ("slice_idx", C.c_uint16),
                ("cbp_luma", C.c_uint8), ("cbp_chroma", C.c_uint8), ("qp_y", C.c_int8), ("qp_c", C.c_int8 * 2),
                ("intra16_mode", C.c_uint8), ("chroma_mode", C.c_uint8), ("reserved0", C.c_uint8),
                ("cbp_blks", C.c_uint16), ("coeff_count", C.c_uint16), ("coeff_offset", C.c_uint32),
                ("u", MbUnion), ("motion", C.c_uint32)]


class MbMotion(C.Structure):
    _fields_ = [("mv", C.c_int16 * 2 * 16 * 2), ("ref_idx", C.c_int8 * 16 * 2), ("ref_pic", C.c_int8 * 16 * 2)]


class Slice(C.Structure):
    _fields_ = [("slice_type", C.c_uint8), ("disable_deblocking_filter_idc", C.c_uint8),
                ("filter_offset_a", C.c_int8), ("filter_offset_b", C.c_int8),
                ("luma_log2_weight_denom", C.c_uint8), ("chroma_log2_weight_denom", C.c_uint8),
                ("weighted_pred_flag", C.c_uint8), ("weighted_bipred_idc", C.c_uint8),
                ("constrained_intra_pred_flag", C.c_uint8), ("direct_spatial_mv_pred_flag", C.c_uint8),
                ("num_ref", C.c_uint8 * 2),
                ("ref_pic_list", C.c_int8 * MAX_REFS * 2),
                ("wp_weight", C.c_int8 * MAX_REFS * 3 * 2), ("wp_offset", C.c_int8 * MAX_REFS * 3 * 2),
                ("implicit_w1", C.c_int16 * MAX_REFS * MAX_REFS),
                ("level_scale_4x4", C.c_uint16 * 16 * 6 * 3 * 2), ("level_scale_8x8", C.c_uint16 * 64 * 6 * 2),
                ("reserved", C.c_uint8 * 20)]


class PicParams(C.Structure):
    _fields_ = [("num_slices", C.c_int32), ("num_ref_frames", C.c_int32), ("ref_frames", C.c_int32 * MAX_REFS),
                ("run_deblock", C.c_int32), ("poc", C.c_int32), ("ref_poc", C.c_int32 * MAX_REFS),
                ("ref_long_term", C.c_uint8 * MAX_REFS), ("direct_8x8_inference_flag", C.c_int32),
                ("structure", C.c_int32), ("ref_structure", C.c_uint8 * MAX_REFS)]


class SeqParams(C.Structure):
    _fields_ = [("width_mbs", C.c_int32), ("height_mbs", C.c_int32), ("direct_8x8_inference_flag", C.c_int32),
                ("max_frames", C.c_int32), ("max_pictures_in_flight", C.c_int32),
                ("max_slices_per_picture", C.c_int32), ("max_levels_per_picture", C.c_int32)]


class PicBuffers(C.Structure):
    _fields_ = [("mbs", C.POINTER(Mb)), ("slices", C.POINTER(Slice)), ("stream", C.POINTER(C.c_uint32)),
                ("stream_capacity", C.c_uint32), ("picture", C.c_int32)]


class BenchPicture(C.Structure):
    _fields_ = [("pp", PicParams), ("dst", C.c_int32), ("stream_id", C.c_int32), ("head", C.c_void_p),
                ("stream", C.c_void_p), ("stream_words", C.c_uint32), ("pitch_y", C.c_int32), ("out", C.c_void_p)]


class PicInfo(C.Structure):
    _fields_ = [("pic_index", C.c_int32), ("pic_type", C.c_int32), ("used_for_reference", C.c_int32),
                ("poc", C.c_int32), ("num_refs", C.c_int32), ("ref_pic_index", C.c_int32 * MAX_REFS),
                ("last_use_of_ref", C.c_int32 * MAX_REFS), ("num_levels", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("pictures", C.c_uint64), ("macroblocks", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("waves", C.c_uint64)]


assert C.sizeof(Mb) == 32 and C.sizeof(MbMotion) == 192 and C.sizeof(Slice) == 5216

_synth = None
_recon = None


def synth_lib():
    global _synth
    if _synth is None:
        L = C.CDLL(os.path.join(HERE, "libh264synth.so"))
        L.h264s_open.restype = C.c_void_p
        L.h264s_open.argtypes = [C.c_int] * 5
        L.h264s_close.argtypes = [C.c_void_p]
        L.h264s_get_seq.argtypes = [C.c_void_p, C.POINTER(SeqParams), C.POINTER(C.c_int)]
        L.h264s_next.restype = C.c_int
        L.h264s_next.argtypes = [C.c_void_p, C.POINTER(PicInfo), C.POINTER(PicParams), C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_uint32]
        L.h264s_account.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        _synth = L
    return _synth


def recon_lib():
    """The CUDA engine.  Raises if it is not built: there is no CPU fallback."""
    global _recon
    if _recon is None:
        # H264R_LIB: another build of the same library (kernel-tuning experiments, scripts/tune.sh)
        path = os.environ.get("H264R_LIB") or os.path.join(HERE, "libh264recon.so")
        if not os.path.exists(path):
            raise RuntimeError("libh264recon.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                               "the engine has no CPU fallback")
        L = C.CDLL(path)
        P = C.c_void_p
        L.h264r_create.restype = C.c_int
        L.h264r_create.argtypes = [C.POINTER(P), C.c_int, C.POINTER(SeqParams)]
        L.h264r_destroy.argtypes = [P]
        L.h264r_frame_alloc.argtypes = [P, C.POINTER(C.c_int32)]
        L.h264r_frame_release.argtypes = [P, C.c_int32]
        L.h264r_picture_begin.argtypes = [P, C.c_int32, C.POINTER(PicParams), C.POINTER(PicBuffers)]
        L.h264r_picture_update.argtypes = [P, C.c_int32, C.POINTER(PicParams)]
        L.h264r_picture_fill.restype = C.c_int64
        L.h264r_picture_fill.argtypes = [P, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint32]
        L.h264r_picture_submit.argtypes = [P, C.c_int32, C.c_uint32]
        L.h264r_pack_picture.restype = C.c_int64
        L.h264r_pack_picture.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
        L.h264r_bench_kernel_name.restype = C.c_char_p
        L.h264r_bench_kernel_name.argtypes = [C.c_int]
        L.h264r_bench_feed.restype = C.c_double
        L.h264r_bench_feed.argtypes = [P, C.POINTER(BenchPicture), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.h264r_bench_copy_ceiling.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
        L.h264r_flush.argtypes = [P]
        L.h264r_wait.argtypes = [P, C.c_int32]
        L.h264r_frame_download.argtypes = [P, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.h264r_frame_upload.argtypes = [P, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.h264r_field_copy.argtypes = [P, C.c_int32, P, C.c_int32, C.c_int, C.c_int]
        L.h264r_replay_last_flush.argtypes = [P, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.h264r_frame_download_async.argtypes = [P, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.h264r_frame_download_cropped.argtypes = [P, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_int, C.c_int]
        L.h264r_host_alloc.restype = C.c_void_p
        L.h264r_host_alloc.argtypes = [C.c_size_t]
        L.h264r_host_free.argtypes = [C.c_void_p]
        L.h264r_get_stats.argtypes = [P, C.POINTER(Stats)]
        L.h264r_strerror.restype = C.c_char_p
        L.h264r_strerror.argtypes = [C.c_int]
        L.h264r_last_cuda_error.restype = C.c_char_p
        L.h264r_last_cuda_error.argtypes = [P]
        L.h264r_device_count.restype = C.c_int
        _recon = L
    return _recon


class Picture:
    """One generated picture: owns its host buffers (numpy-free, plain ctypes arrays)."""
    __slots__ = ("info", "pp", "mbs", "motion", "slices", "levels", "nmb")

    def __init__(self, nmb, scratch_levels=None):
        self.nmb = nmb
        self.info = PicInfo()
        self.pp = PicParams()
        self.mbs = (Mb * nmb)()
        self.motion = (MbMotion * nmb)()
        self.slices = (Slice * 4)()
        self.levels = scratch_levels          # replaced by a right-sized copy after generation


class SynthStream:
    def __init__(self, config, stream_idx=0, width_mbs=0, height_mbs=0, num_frames=0):
        self.L = synth_lib()
        self.h = self.L.h264s_open(config, stream_idx, width_mbs, height_mbs, num_frames)
        if not self.h:
            raise ValueError("bad synth config")
        self.seq = SeqParams()
        n = C.c_int()
        self.L.h264s_get_seq(self.h, C.byref(self.seq), C.byref(n))
        self.num_frames = n.value
        self.nmb = self.seq.width_mbs * self.seq.height_mbs
        self._scratch = None

    def next(self):
        cap = COEFFS_PER_MB * self.nmb
        if self._scratch is None:
            self._scratch = (C.c_uint32 * cap)()
        pic = Picture(self.nmb, self._scratch)
        ok = self.L.h264s_next(self.h, C.byref(pic.info), C.byref(pic.pp), pic.mbs, pic.motion, pic.slices,
                               self._scratch, cap)
        if ok <= 0:
            return None
        n = pic.info.num_levels
        pic.levels = (C.c_uint32 * max(n, 1))()
        C.memmove(pic.levels, self._scratch, 4 * n)
        return pic

    def __iter__(self):
        while True:
            p = self.next()
            if p is None:
                return
            yield p

    def close(self):
        if self.h:
            self.L.h264s_close(self.h)
            self.h = None

    def __del__(self):
        self.close()


class EngineError(RuntimeError):
    pass


class Engine:
    """Thin object wrapper over the C ABI of libh264recon.so (one context == one GPU)."""

    def __init__(self, seq, device=0, max_frames=8, max_pictures=4, max_slices=4, max_levels=0):
        self.L = recon_lib()
        sp = SeqParams.from_buffer_copy(seq)
        sp.max_frames, sp.max_pictures_in_flight, sp.max_slices_per_picture = max_frames, max_pictures, max_slices
        sp.max_levels_per_picture = max_levels
        self.seq = sp
        self.nmb = sp.width_mbs * sp.height_mbs
        self.w, self.h = sp.width_mbs * 16, sp.height_mbs * 16
        self.ctx = C.c_void_p()
        self._check(self.L.h264r_create(C.byref(self.ctx), device, C.byref(sp)), "h264r_create")

    def _check(self, rc, what):
        if rc != 0:
            detail = self.L.h264r_last_cuda_error(self.ctx).decode() if self.ctx else ""
            raise EngineError(f"{what}: {self.L.h264r_strerror(rc).decode()} ({rc}) {detail}")

    def frame_alloc(self):
        f = C.c_int32()
        self._check(self.L.h264r_frame_alloc(self.ctx, C.byref(f)), "h264r_frame_alloc")
        return f.value

    def frame_release(self, f):
        self._check(self.L.h264r_frame_release(self.ctx, f), "h264r_frame_release")

    def submit(self, pic, dst, ref_frames):
        """picture_begin + picture_fill (copy and pack a generated Picture into the pinned staging) + picture_submit."""
        pp = PicParams.from_buffer_copy(pic.pp)
        for i, f in enumerate(ref_frames):
            pp.ref_frames[i] = f
        bufs = PicBuffers()
        self._check(self.L.h264r_picture_begin(self.ctx, dst, C.byref(pp), C.byref(bufs)), "h264r_picture_begin")
        words = self.L.h264r_picture_fill(self.ctx, bufs.picture, pic.mbs, pic.motion, pic.slices, pp.num_slices,
                                          pic.levels, pic.info.num_levels)
        if words < 0:
            # give the slot back (a submit that fails frees it) and report
            self.L.h264r_picture_submit(self.ctx, bufs.picture, 0xFFFFFFFF)
            self._check(int(words), "h264r_picture_fill")
        self._check(self.L.h264r_picture_submit(self.ctx, bufs.picture, words), "h264r_picture_submit")

    def flush(self):
        self._check(self.L.h264r_flush(self.ctx), "h264r_flush")

    def wait(self, f=-1):
        self._check(self.L.h264r_wait(self.ctx, f), "h264r_wait")

    def download(self, f):
        y = (C.c_uint8 * (self.w * self.h))()
        cb = (C.c_uint8 * (self.w * self.h // 4))()
        cr = (C.c_uint8 * (self.w * self.h // 4))()
        self._check(self.L.h264r_frame_download(self.ctx, f, y, cb, cr, self.w, self.w // 2), "h264r_frame_download")
        return bytes(y), bytes(cb), bytes(cr)

    def download_cropped(self, f, left, right, top, bottom):
        """The display rectangle (frame cropping offsets in luma samples); does not wait for pictures queued behind `f`."""
        w, h = self.w - left - right, self.h - top - bottom
        y = (C.c_uint8 * (w * h))()
        cb = (C.c_uint8 * (w * h // 4))()
        cr = (C.c_uint8 * (w * h // 4))()
        self._check(self.L.h264r_frame_download_cropped(self.ctx, f, left, right, top, bottom, y, cb, cr, w, w // 2),
                    "h264r_frame_download_cropped")
        return bytes(y), bytes(cb), bytes(cr)

    def upload(self, f, y, cb, cr):
        self._check(self.L.h264r_frame_upload(self.ctx, f, y, cb, cr, self.w, self.w // 2), "h264r_frame_upload")

    REPLAY_H2D, REPLAY_TIME_KERNELS, REPLAY_ASYNC = 1, 2, 4

    def kernel_names(self):
        """Names of the kernel kinds the replay hook times (include/h264recon_bench.h)."""
        names = []
        while True:
            n = self.L.h264r_bench_kernel_name(len(names))
            if not n:
                return names
            names.append(n.decode())

    def replay(self, iterations=1, flags=0):
        """Re-runs the last flush (include/h264recon_bench.h); returns (ms[total, kind 0, kind 1, ...], launches[_, ...])."""
        ms = (C.c_float * 9)()
        n = (C.c_int * 9)()
        self._check(self.L.h264r_replay_last_flush(self.ctx, iterations, flags, ms, n), "h264r_replay_last_flush")
        k = 1 + len(self.kernel_names())
        return list(ms)[:k], list(n)[:k]

    def host_alloc(self, nbytes):
        p = self.L.h264r_host_alloc(nbytes)
        if not p:
            raise EngineError("h264r_host_alloc failed")
        return p

    def host_free(self, p):
        self.L.h264r_host_free(p)

    def download_async(self, f, y_ptr, cb_ptr, cr_ptr):
        self._check(self.L.h264r_frame_download_async(self.ctx, f, y_ptr, cb_ptr, cr_ptr, self.w, self.w // 2),
                    "h264r_frame_download_async")

    def begin(self, dst, pp):
        """picture_begin only: returns the staging buffers (and the picture handle) so a producer can write into pinned
        memory directly."""
        bufs = PicBuffers()
        self._check(self.L.h264r_picture_begin(self.ctx, dst, C.byref(pp), C.byref(bufs)), "h264r_picture_begin")
        return bufs

    def submit_filled(self, picture, stream_words):
        self._check(self.L.h264r_picture_submit(self.ctx, picture, stream_words), "h264r_picture_submit")

    def stats(self):
        s = Stats()
        self._check(self.L.h264r_get_stats(self.ctx, C.byref(s)), "h264r_get_stats")
        return s

    def close(self):
        if self.ctx:
            self.L.h264r_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def streams_of_rank(rank, world, streams_per_gpu):
    """Stream sharding: streams are independent units, no exchange step.  Weak scaling: every
    rank owns `streams_per_gpu` streams; global stream ids are contiguous per rank, so seeds differ across GPUs."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank * streams_per_gpu, (rank + 1) * streams_per_gpu))
