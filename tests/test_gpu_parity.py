"""GPU parity: the CUDA engine, called through the C ABI (libh264recon.so), against the CPU oracle on the same
seeded synthetic macroblock data -- every sample of every picture, bit-exact -- and against the golden per-frame
MD5 digests produced by the reference's own Decoder."""
import hashlib
import json
import os

import pytest

import oracle_py as O
import pyapi

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "ref_digests.json")) as f:
    GOLDEN = json.load(f)["cases"]


def run_stream_gpu(eng, config, sidx, w, h, n, flush_every=1, on_picture=None):
    """One stream through the engine; pictures are queued and flushed every `flush_every` submissions."""
    st = pyapi.SynthStream(config, sidx, w, h, n)
    frames, pending, digests, pics = {}, [], {}, []
    for pic in st:
        dst = eng.frame_alloc()
        frames[pic.info.pic_index] = dst
        refs = [frames[pic.info.ref_pic_index[i]] for i in range(pic.info.num_refs)]
        eng.submit(pic, dst, refs)
        pending.append((pic, dst))
        if len(pending) >= flush_every:
            eng.flush()
            eng.wait()
            for p, d in pending:
                planes = eng.download(d)
                digests[p.info.pic_index] = hashlib.md5(b"".join(planes)).hexdigest()
                if on_picture:
                    on_picture(p, planes)
            pending = []
    if pending:
        eng.flush()
        eng.wait()
        for p, d in pending:
            planes = eng.download(d)
            digests[p.info.pic_index] = hashlib.md5(b"".join(planes)).hexdigest()
            if on_picture:
                on_picture(p, planes)
    st.close()
    return [digests[i] for i in sorted(digests)]


def first_diff(a, b):
    return next((k for k in range(len(a)) if a[k] != b[k]), None)


@pytest.mark.parametrize("case", GOLDEN, ids=lambda c: f"cfg{c['config']}-s{c['stream']}-{c['width_mbs']}x{c['height_mbs']}")
def test_gpu_equals_oracle_and_golden(case):
    cfg, sidx, w, h, n = case["config"], case["stream"], case["width_mbs"], case["height_mbs"], case["frames"]
    st = pyapi.SynthStream(cfg, sidx, w, h, n)
    seq, nfr = st.seq, st.num_frames
    st.close()
    port = O.CpuDecoder("port", seq)
    want = []
    O.run_stream(port, cfg, sidx, w, h, n, on_picture=lambda p, pl: want.append(pl))
    port.close()
    eng = pyapi.Engine(seq, max_frames=nfr + 1, max_pictures=2)
    got = []
    digests = run_stream_gpu(eng, cfg, sidx, w, h, n, flush_every=1, on_picture=lambda p, pl: got.append(pl))
    eng.close()
    assert len(got) == len(want)
    for i, (a, b) in enumerate(zip(want, got)):
        for name, pa, pb in zip(("Y", "Cb", "Cr"), a, b):
            if pa != pb:
                k = first_diff(pa, pb)
                width = seq.width_mbs * (16 if name == "Y" else 8)
                pytest.fail(f"picture {i} plane {name}: first difference at x={k % width} y={k // width} "
                            f"(oracle {pa[k]}, gpu {pb[k]})")
    assert digests == case["md5"], "GPU output differs from the reference's golden digests"


def test_gpu_batched_multi_stream_waves():
    """Eight independent streams queued together and flushed once: the runtime must split them into dependency
    waves (I, then P, then the two Bs, ...) and every picture must still equal the oracle's."""
    cfg, w, h, n, nstreams = 5, 20, 12, 7, 8
    st = pyapi.SynthStream(cfg, 0, w, h, n)
    seq = st.seq
    st.close()
    want = {}
    for s in range(nstreams):                     # stream 2 carries direct_8x8_inference_flag = 0: every stream has its own sps
        sq = pyapi.SynthStream(cfg, s, w, h, n).seq
        port = O.CpuDecoder("port", sq)
        want[s] = O.run_stream(port, cfg, s, w, h, n)
        port.close()
    eng = pyapi.Engine(seq, max_frames=nstreams * n, max_pictures=nstreams * n)
    streams = [pyapi.SynthStream(cfg, s, w, h, n) for s in range(nstreams)]
    frames = [dict() for _ in range(nstreams)]
    order = []
    for _ in range(n):
        for s, st in enumerate(streams):
            pic = st.next()
            dst = eng.frame_alloc()
            frames[s][pic.info.pic_index] = dst
            eng.submit(pic, dst, [frames[s][pic.info.ref_pic_index[i]] for i in range(pic.info.num_refs)])
            order.append((s, pic.info.pic_index, dst))
    eng.flush()
    eng.wait()
    stats = eng.stats()
    assert stats.pictures == nstreams * n
    assert stats.waves < nstreams * n, "pictures of independent streams must share launches"
    for s, idx, dst in order:
        d = hashlib.md5(b"".join(eng.download(dst))).hexdigest()
        assert d == want[s][idx], f"stream {s} picture {idx} differs"
    # device-resident replay (the bench's kernel-only path) must reproduce the same frames
    eng.replay(2)
    eng.replay(1, pyapi.Engine.REPLAY_H2D | pyapi.Engine.REPLAY_TIME_KERNELS)
    for s, idx, dst in order:
        d = hashlib.md5(b"".join(eng.download(dst))).hexdigest()
        assert d == want[s][idx], f"after replay: stream {s} picture {idx} differs"
    eng.close()


def test_engine_rejects_bad_input():
    """Slice tables are checked on the host at submit (O(slices)); macroblock content is checked by the kernels that read it:
    an out-of-range value is clamped (no out-of-bounds access, no crash) and h264r_wait reports H264R_ERR_INVALID."""
    st = pyapi.SynthStream(2, 0, 6, 4, 3)
    eng = pyapi.Engine(st.seq, max_frames=4, max_pictures=2)
    pic = st.next()
    dst = eng.frame_alloc()
    pic.slices[0].slice_type = 4                  # SI: outside the supported subset -> refused on the host
    with pytest.raises(pyapi.EngineError, match="unsupported"):
        eng.submit(pic, dst, [])
    pic.slices[0].slice_type = pyapi.I_SLICE
    # device-side checks: slice index, QP, level offset, level position
    saved = (pic.mbs[0].slice_idx, pic.mbs[3].qp_y, pic.mbs[5].coeff_offset)
    pic.mbs[0].slice_idx = 7
    pic.mbs[3].qp_y = 99
    pic.mbs[5].coeff_offset = 0x7FFFFFF0
    eng.submit(pic, dst, [])
    eng.flush()
    with pytest.raises(pyapi.EngineError, match="invalid"):
        eng.wait()
    eng.wait()                                    # the error is reported once
    pic.mbs[0].slice_idx, pic.mbs[3].qp_y, pic.mbs[5].coeff_offset = saved
    eng.submit(pic, dst, [])
    eng.flush()
    eng.wait()
    # an inter picture whose motion names a reference slot the picture does not have
    p2 = st.next()
    assert p2.info.num_refs >= 1
    d2 = eng.frame_alloc()
    inter = next(i for i in range(p2.nmb) if not (p2.mbs[i].flags & 1))
    for b in range(16):
        p2.motion[inter].ref_pic[0][b] = 9
    eng.submit(p2, d2, [dst] * p2.info.num_refs)
    eng.flush()
    with pytest.raises(pyapi.EngineError, match="invalid"):
        eng.wait()
    with pytest.raises(pyapi.EngineError):
        eng.frame_release(99)
    eng.close()
    st.close()


def _oracle_digests(job):
    """Worker of the process pool below: the CPU restatement on one stream -> per-picture digests."""
    cfg, sidx, w, h, n = job
    st = pyapi.SynthStream(cfg, sidx, w, h, n)
    seq = st.seq
    st.close()
    port = O.CpuDecoder("port", seq)
    d = O.run_stream(port, cfg, sidx, w, h, n)
    port.close()
    return d


def oracle_digests_parallel(jobs):
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    with ProcessPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1), mp_context=mp.get_context("spawn")) as ex:
        return list(ex.map(_oracle_digests, jobs))


def test_gpu_full_size_multi_stream_workload():
    """The bench workload at full size (BASELINE configs[4]: 64 independent 1080p High-profile streams, here 7 pictures
    each = waves of 64, 64, 192 and 128 pictures, flushed at once like bench.py does).  EVERY picture of EVERY stream is
    compared with the oracle (the CPU restatement runs in a process pool, one stream per job); the same digests must
    come back when the descriptions are replayed from HBM (kernel-only path) and re-uploaded (end-to-end path)."""
    cfg, n, nstreams = 5, 7, 64
    st = pyapi.SynthStream(cfg, 0, 0, 0, n)
    seq = st.seq
    st.close()
    assert (seq.width_mbs, seq.height_mbs) == (120, 68)
    want = oracle_digests_parallel([(cfg, s, 0, 0, n) for s in range(nstreams)])
    # streams differ in direct_8x8_inference_flag (every third one is 0): the flag travels with each picture
    flags = {pyapi.SynthStream(cfg, s, 0, 0, 1).seq.direct_8x8_inference_flag for s in range(nstreams)}
    assert flags == {0, 1}, "the workload must contain both kinds of streams"
    eng = pyapi.Engine(seq, max_frames=nstreams * n, max_pictures=nstreams * n, max_slices=4, max_levels=8160 * 96)
    streams = [pyapi.SynthStream(cfg, s, 0, 0, n) for s in range(nstreams)]
    frames = [dict() for _ in range(nstreams)]
    order = []
    for _ in range(n):
        for s, st in enumerate(streams):
            pic = st.next()
            dst = eng.frame_alloc()
            frames[s][pic.info.pic_index] = dst
            eng.submit(pic, dst, [frames[s][pic.info.ref_pic_index[i]] for i in range(pic.info.num_refs)])
            order.append((s, pic.info.pic_index, dst))
    for st in streams:
        st.close()
    eng.flush()
    eng.wait()

    def check_all(what):
        for s, idx, dst in order:
            d = hashlib.md5(b"".join(eng.download(dst))).hexdigest()
            assert d == want[s][idx], f"{what}: stream {s} picture {idx} differs from the oracle"

    check_all("first flush")
    eng.replay(1, 0)
    check_all("kernel-only replay")
    eng.replay(1, pyapi.Engine.REPLAY_H2D | pyapi.Engine.REPLAY_ASYNC)
    eng.wait()
    check_all("end-to-end replay")
    eng.close()


def test_gpu_4k_wavefront_stress():
    """BASELINE configs[3] at its full size (3840x2160 = 240x135 MB): all-intra + P + B pictures with low QP and
    deblock offsets of +-6 -- the longest wavefronts (508 steps) -- against the oracle, every picture."""
    cfg, sidx, n = 4, 2, 4
    st = pyapi.SynthStream(cfg, sidx, 0, 0, n)
    seq, nfr = st.seq, st.num_frames
    st.close()
    assert (seq.width_mbs, seq.height_mbs) == (240, 135)
    port = O.CpuDecoder("port", seq)
    want = O.run_stream(port, cfg, sidx, 0, 0, n)
    port.close()
    eng = pyapi.Engine(seq, max_frames=nfr + 1, max_pictures=nfr, max_slices=4)
    got = run_stream_gpu(eng, cfg, sidx, 0, 0, n, flush_every=nfr)
    eng.close()
    assert got == want


def test_cropped_download_is_the_display_rectangle():
    """h264r_frame_download_cropped (the write_out_picture replacement, output.cc:147-187) against slices of the full
    planes, without an intervening h264r_wait: the copy must order itself behind the picture's own wave."""
    cfg, sidx, w, h, n = 3, 1, 24, 14, 4
    st = pyapi.SynthStream(cfg, sidx, w, h, n)
    eng = pyapi.Engine(st.seq, max_frames=n + 1, max_pictures=n)
    frames, order = {}, []
    for pic in st:
        dst = eng.frame_alloc()
        frames[pic.info.pic_index] = dst
        eng.submit(pic, dst, [frames[pic.info.ref_pic_index[i]] for i in range(pic.info.num_refs)])
        order.append(dst)
    st.close()
    eng.flush()
    W, H = w * 16, h * 16
    crops = [(0, 0, 0, 8), (2, 6, 4, 10), (16, 0, 0, 2), (0, 0, 0, 0)]
    got = [eng.download_cropped(f, *crops[i % len(crops)]) for i, f in enumerate(order)]     # no wait before
    eng.wait(order[1])                                       # one frame: returns once ITS wave is done
    with pytest.raises(pyapi.EngineError):
        eng.wait(10_000)
    eng.wait()
    for i, f in enumerate(order):
        l, r, t, b = crops[i % len(crops)]
        y, cb, cr = eng.download(f)
        want_y = b"".join(y[j * W + l:j * W + W - r] for j in range(t, H - b))
        want_cb = b"".join(cb[j * (W // 2) + l // 2:j * (W // 2) + (W - r) // 2] for j in range(t // 2, (H - b) // 2))
        want_cr = b"".join(cr[j * (W // 2) + l // 2:j * (W // 2) + (W - r) // 2] for j in range(t // 2, (H - b) // 2))
        assert got[i] == (want_y, want_cb, want_cr), f"picture {i} crop {crops[i % len(crops)]}"
    with pytest.raises(pyapi.EngineError):
        eng.download_cropped(order[0], 1, 0, 0, 0)          # odd offsets do not exist in 4:2:0
    eng.close()


def test_concurrent_fill_through_the_public_entry_points():
    """Pictures of eight streams are filled CONCURRENTLY by four feeder threads (h264r_picture_begin -> write the staging
    -> h264r_picture_submit, include/h264recon_bench.h h264r_bench_feed) while the calling thread flushes every five
    submitted pictures and downloads asynchronously; two passes back to back (staging slots and frames are reused while
    earlier pictures are still in flight).  Every frame must equal the oracle's."""
    import ctypes as C
    cfg, w, h, n, nstreams = 5, 9, 6, 7, 8
    lib = pyapi.recon_lib()
    streams = [pyapi.SynthStream(cfg, s, w, h, n) for s in range(nstreams)]
    seq, nmb = streams[0].seq, streams[0].nmb
    want = {}
    for s in range(nstreams):
        port = O.CpuDecoder("port", pyapi.SynthStream(cfg, s, w, h, n).seq)
        want[s] = O.run_stream(port, cfg, s, w, h, n)
        port.close()
    npics = nstreams * n
    eng = pyapi.Engine(seq, max_frames=npics, max_pictures=6, max_slices=4)     # far fewer staging slots than pictures
    fbytes = eng.w * eng.h * 3 // 2
    out_host = eng.host_alloc(fbytes * npics)
    table = (pyapi.BenchPicture * npics)()
    keep, where, frames = [], {}, [dict() for _ in range(nstreams)]
    k = 0
    for _ in range(n):
        for s, st in enumerate(streams):
            pic = st.next()
            head = C.create_string_buffer(C.sizeof(pyapi.Mb) * nmb + C.sizeof(pyapi.Slice) * pic.pp.num_slices)
            cap = pic.info.num_levels + 48 * nmb
            stream = (C.c_uint32 * cap)()
            words = lib.h264r_pack_picture(nmb, pic.mbs, pic.motion, pic.levels, pic.info.num_levels, head, stream, cap)
            assert words >= 0
            C.memmove(C.addressof(head) + C.sizeof(pyapi.Mb) * nmb, pic.slices, C.sizeof(pyapi.Slice) * pic.pp.num_slices)
            keep.append((head, stream))
            dst = eng.frame_alloc()
            frames[s][pic.info.pic_index] = dst
            e = table[k]
            C.memmove(C.byref(e.pp), C.byref(pic.pp), C.sizeof(pyapi.PicParams))
            for r in range(pic.info.num_refs):
                e.pp.ref_frames[r] = frames[s][pic.info.ref_pic_index[r]]
            e.dst, e.stream_id = dst, s
            e.head, e.stream, e.stream_words = C.addressof(head), C.addressof(stream), int(words)
            e.pitch_y, e.out = eng.w, out_host + k * fbytes
            where[(s, pic.info.pic_index)] = k
            k += 1
    for st in streams:
        st.close()
    fill, flush = C.c_double(), C.c_double()
    t = lib.h264r_bench_feed(eng.ctx, table, npics, nmb, 4, 5, 2, C.byref(fill), C.byref(flush))
    assert t > 0, lib.h264r_strerror(int(t)).decode()
    for (s, idx), kk in where.items():
        d = hashlib.md5(C.string_at(out_host + kk * fbytes, fbytes)).hexdigest()
        assert d == want[s][idx], f"stream {s} picture {idx} differs from the oracle"
    assert eng.stats().pictures == 2 * npics
    eng.host_free(out_host)
    eng.close()


def test_long_run_wraps_the_event_and_table_rings():
    """3000 pictures of a 2x1-MB stream, one h264r_flush per picture, two staging slots, three frames: the event ring (8192
    events, three per wave), the picture-table ring and the staging slots wrap many times; every picture must still equal
    the oracle's (a stale event or table entry would show up as a wrong or torn picture)."""
    cfg, w, h, n = 1, 2, 1, 3000
    st = pyapi.SynthStream(cfg, 0, w, h, n)
    seq = st.seq
    port = O.CpuDecoder("port", seq)
    want = O.run_stream(port, cfg, 0, w, h, n)
    port.close()
    eng = pyapi.Engine(seq, max_frames=3, max_pictures=2)
    frames, got = {}, []
    for pic in st:
        dst = eng.frame_alloc()
        frames[pic.info.pic_index] = dst
        eng.submit(pic, dst, [frames[pic.info.ref_pic_index[i]] for i in range(pic.info.num_refs)])
        eng.flush()
        for i in range(pic.info.num_refs):
            if pic.info.last_use_of_ref[i]:
                eng.frame_release(frames.pop(pic.info.ref_pic_index[i]))     # queued work keeps reading it: stream order
        eng.wait(dst)
        got.append(hashlib.md5(b"".join(eng.download(dst))).hexdigest())
    st.close()
    eng.wait()
    eng.close()
    bad = [i for i, (a, b) in enumerate(zip(got, want)) if a != b]
    assert not bad, f"pictures {bad[:10]} differ"


def test_resume_from_uploaded_reference_frames_and_statistics():
    """A second context takes over a stream in the middle: the reference frames it needs are downloaded from the first
    context and put into its frame pool with h264r_frame_upload (the way a decoder hands over pictures it reconstructed
    elsewhere); the remaining pictures must still equal the oracle's.  h264r_get_stats must account for exactly the
    pictures, macroblocks and copies that went through each context."""
    cfg, sidx, w, h, n, split = 2, 1, 10, 6, 9, 4
    st = pyapi.SynthStream(cfg, sidx, w, h, n)
    seq = st.seq
    port = O.CpuDecoder("port", seq)
    want = O.run_stream(port, cfg, sidx, w, h, n)
    port.close()
    a = pyapi.Engine(seq, max_frames=n + 1, max_pictures=2)
    b = pyapi.Engine(seq, max_frames=n + 1, max_pictures=2)
    frames_a, frames_b, got = {}, {}, []
    fbytes = w * h * 384
    for pic in st:
        i = pic.info.pic_index
        refs_idx = [pic.info.ref_pic_index[k] for k in range(pic.info.num_refs)]
        if len(got) < split:
            eng, frames = a, frames_a
        else:
            eng, frames = b, frames_b
            for r in refs_idx:                              # hand over the references context b does not have yet
                if r not in frames_b:
                    y, cb, cr = a.download(frames_a[r])
                    frames_b[r] = b.frame_alloc()
                    b.upload(frames_b[r], y, cb, cr)
        frames[i] = eng.frame_alloc()
        eng.submit(pic, frames[i], [frames[r] for r in refs_idx])
        eng.flush()
        eng.wait(frames[i])
        got.append(hashlib.md5(b"".join(eng.download(frames[i]))).hexdigest())
    st.close()
    assert got == want, f"first difference at picture {first_diff(got, want)}"
    sa, sb = a.stats(), b.stats()
    assert (sa.pictures, sb.pictures) == (split, n - split)
    assert (sa.macroblocks, sb.macroblocks) == (split * w * h, (n - split) * w * h)
    assert sa.waves == split and sb.waves == n - split       # one flush per picture
    assert sa.kernel_launches >= 3 * split and sb.kernel_launches >= 3 * (n - split)
    uploads = len(frames_b) - (n - split)
    assert uploads >= 1
    assert sb.h2d_bytes > uploads * fbytes                   # the uploaded planes plus the picture descriptions
    assert sa.d2h_bytes == (split + uploads) * fbytes and sb.d2h_bytes == (n - split) * fbytes
    a.close(); b.close()


def test_field_copy_between_a_frame_context_and_a_field_context():
    """h264r_field_copy (dpb_split_field / dpb_combine_field_yuv on the device): frames decoded in a frame context are split
    into fields of a field context, which must equal the de-interleaved lines of the frame; combining the two fields into a
    fresh frame must give the frame back.  The copies are enqueued right behind the flush, without waiting."""
    import ctypes as C
    cfg, sidx, w, h, n = 3, 1, 10, 6, 4
    st = pyapi.SynthStream(cfg, sidx, w, h, n)
    seq = st.seq
    fa = pyapi.Engine(seq, max_frames=2 * n, max_pictures=n)
    fseq = pyapi.SeqParams.from_buffer_copy(seq)
    fseq.height_mbs = h // 2
    fb = pyapi.Engine(fseq, max_frames=2 * n, max_pictures=n)
    L = pyapi.recon_lib()
    frames, order = {}, []
    for pic in st:
        dst = fa.frame_alloc()
        frames[pic.info.pic_index] = dst
        fa.submit(pic, dst, [frames[pic.info.ref_pic_index[i]] for i in range(pic.info.num_refs)])
        order.append(dst)
    st.close()
    # not flushed yet: the source is still to be written by a queued picture
    t0 = fb.frame_alloc()
    assert L.h264r_field_copy(fa.ctx, order[0], fb.ctx, t0, 0, 1) == -5
    fa.flush()
    fields, back = [], []
    for f in order:                                          # split every frame, then combine the fields into a new frame
        top, bot, again = (t0 if f == order[0] else fb.frame_alloc()), fb.frame_alloc(), fa.frame_alloc()
        for parity, fld in ((0, top), (1, bot)):
            assert L.h264r_field_copy(fa.ctx, f, fb.ctx, fld, parity, 1) == 0
        for parity, fld in ((0, top), (1, bot)):
            assert L.h264r_field_copy(fa.ctx, again, fb.ctx, fld, parity, 0) == 0
        fields.append((top, bot)); back.append(again)
    W, H = w * 16, h * 16
    for f, (top, bot), again in zip(order, fields, back):
        y, cb, cr = fa.download(f)
        for parity, fld in ((0, top), (1, bot)):
            fy, fcb, fcr = fb.download(fld)
            assert fy == b"".join(y[r * W:(r + 1) * W] for r in range(parity, H, 2))
            assert fcb == b"".join(cb[r * (W // 2):(r + 1) * (W // 2)] for r in range(parity, H // 2, 2))
            assert fcr == b"".join(cr[r * (W // 2):(r + 1) * (W // 2)] for r in range(parity, H // 2, 2))
        assert fa.download(again) == (y, cb, cr)
    assert L.h264r_field_copy(fa.ctx, order[0], fa.ctx, order[1], 0, 1) == -1      # one context is not a pair
    assert L.h264r_field_copy(fa.ctx, order[0], fb.ctx, fields[0][0], 2, 1) == -1
    fa.close(); fb.close()
