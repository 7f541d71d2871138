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
    for s in range(nstreams):
        port = O.CpuDecoder("port", seq)
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
    st = pyapi.SynthStream(1, 0, 4, 3, 2)
    eng = pyapi.Engine(st.seq, max_frames=3, max_pictures=2)
    pic = st.next()
    dst = eng.frame_alloc()
    pic.mbs[0].slice_idx = 7                      # out of range -> must be refused on the host, not crash the device
    with pytest.raises(pyapi.EngineError):
        eng.submit(pic, dst, [])
    pic.mbs[0].slice_idx = 0
    eng.submit(pic, dst, [])
    eng.flush()
    eng.wait()
    with pytest.raises(pyapi.EngineError):
        eng.frame_release(99)
    eng.close()
    st.close()


def test_gpu_full_size_multi_stream_workload():
    """The bench workload at full size (BASELINE configs[4]: 64 independent 1080p High-profile streams, here 7 pictures
    each = waves of 64, 64, 192 and 128 pictures, flushed at once like bench.py does).  Four sampled streams are
    compared picture by picture with the oracle; a checksum of checksums over all 448 frames must not change when
    the same descriptions are replayed from HBM (kernel-only path) and re-uploaded (end-to-end path)."""
    cfg, n, nstreams, sampled = 5, 7, 64, (0, 9, 31, 63)
    st = pyapi.SynthStream(cfg, 0, 0, 0, n)
    seq = st.seq
    st.close()
    assert (seq.width_mbs, seq.height_mbs) == (120, 68)
    want = {}
    for s in sampled:
        port = O.CpuDecoder("port", seq)
        want[s] = O.run_stream(port, cfg, s, 0, 0, n)
        port.close()
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

    def digest_all():
        per = {}
        for s, idx, dst in order:
            per[(s, idx)] = hashlib.md5(b"".join(eng.download(dst))).hexdigest()
        total = hashlib.md5("".join(per[k] for k in sorted(per)).encode()).hexdigest()
        return per, total

    per, total = digest_all()
    for s in sampled:
        for idx in range(n):
            assert per[(s, idx)] == want[s][idx], f"stream {s} picture {idx} differs from the oracle"
    eng.replay(1, 0)
    assert digest_all()[1] == total, "kernel-only replay changed the output"
    eng.replay(1, pyapi.Engine.REPLAY_H2D | pyapi.Engine.REPLAY_ASYNC)
    eng.wait()
    assert digest_all()[1] == total, "end-to-end replay changed the output"
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


@pytest.mark.parametrize("groups,policy,extra", [(2, "stream", {}), (3, "role", {}), (2, "role", {"H264R_SIDE_PRIORITY": "low", "H264R_SIDE_GATE": "1"})])
def test_stream_group_options_keep_the_output(monkeypatch, groups, policy, extra):
    """The engine's scheduling options (independent stream groups on their own CUDA streams, by stream or by role;
    side-stream priority and gating -- DESIGN.md section 3, measured and off by default) only reorder launches: every
    picture must still equal the oracle's, also after replays with cross-group dependencies."""
    monkeypatch.setenv("H264R_STREAM_GROUPS", str(groups))
    monkeypatch.setenv("H264R_GROUP_POLICY", policy)
    for k, v in extra.items():
        monkeypatch.setenv(k, v)
    cfg, w, h, n, nstreams = 2, 13, 9, 8, 6
    st = pyapi.SynthStream(cfg, 0, w, h, n)
    seq = st.seq
    st.close()
    want = {}
    for s in range(nstreams):
        port = O.CpuDecoder("port", seq)
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
    for st in streams:
        st.close()
    eng.flush()
    eng.wait()
    for rounds in range(2):
        for s, idx, dst in order:
            d = hashlib.md5(b"".join(eng.download(dst))).hexdigest()
            assert d == want[s][idx], f"groups={groups} policy={policy} round {rounds}: stream {s} picture {idx} differs"
        eng.replay(2, pyapi.Engine.REPLAY_H2D | pyapi.Engine.REPLAY_ASYNC)
        eng.wait()
    eng.close()
