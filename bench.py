#!/usr/bin/env python
"""Benchmark of the H.264 macroblock-reconstruction path (BASELINE.json metric:
"1080p MB/s reconstructed (IDCT+MC+intra+deblock) at 1/2/4/8 B200; % HBM peak").

    python bench.py --gpus N --steps K --warmup W            # our CUDA engine through the C ABI
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU Decoder on the host cores

    python bench.py --config 3 | --config 4                   # single 1080p / 4K stream (latency-bound, reported as such)

Workload (BASELINE.json configs[4], weak scaling): 64 independent synthetic 1080p High-profile I/P/B streams of 16
pictures PER GPU (stream seeds distinct across ranks).  One step = reconstruct all of them once (1024 pictures,
8.36 M macroblocks per GPU).  `value` = macroblocks/s with the picture descriptions resident in HBM (kernels only,
CUDA-event and wall time agree).  `e2e` = the same through the PUBLIC entry points with HOST buffers: feeder threads
(stand-ins for one parser thread per stream) call h264r_picture_begin, write each picture description into the pinned
staging and h264r_picture_submit it; the GPU thread calls h264r_flush (host->device copies + kernels) and
h264r_frame_download_async of every reconstructed frame into pinned host memory; one h264r_wait at the end.  The
synthetic generator stands in for the entropy decoder and is outside both timed regions.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from concurrent.futures import ProcessPoolExecutor, ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "arrow-h264_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyapi  # noqa: E402

METRIC = "1080p macroblocks/s reconstructed (IDCT+MC+intra+deblock)"
UNIT = "MB/s"          # MB = macroblocks; output bytes/s = value * 384
CONFIG_ID = 5          # --config: 5 = BASELINE configs[4] (the metric's workload), 3 = configs[2], 4 = configs[3]


def workload_name(streams, frames):
    if CONFIG_ID == 3:
        return (f"{streams} 1080p (120x68 MB) High-profile I/P/B stream(s) x {frames} pictures (BASELINE configs[2]: 8x8 transform, "
                "intra 8x8); single-stream runs are bound by the wavefront depth (254 MB steps per picture), not by throughput")
    if CONFIG_ID == 4:
        return (f"{streams} 4K 3840x2160 (240x135 MB) High-profile I/P/B stream(s) x {frames} pictures (BASELINE configs[3]: low QP, "
                "deblock offsets +-6, all-intra pictures); bound by the wavefront depth (508 MB steps per picture)")
    if CONFIG_ID == 6:
        return (f"{streams} independent 1080i streams coded in FIELD pictures (120x34 MB per field, field_pic_flag = 1, references "
                f"of both parities) x {frames} fields per GPU; not a BASELINE config (SURVEY.md 8f-4)")
    return (f"{streams} independent 1080p (120x68 MB) High-profile I/P/B streams x {frames} pictures per GPU "
            "(BASELINE configs[4]; 8x8 transform, intra 8x8, bi-pred, weighted prediction, 2 slices on odd pictures; "
            "every third stream has direct_8x8_inference_flag = 0)")


def measured_profile():
    """Per-picture DRAM bytes and warp instructions by kernel and picture type, from the newest ncu --set full capture
    committed under profiles/ (scripts/gpu_profile.sh + scripts/ncu_summary.py; bench.py never runs under a profiler)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_*_traffic.json")))
    if not files:
        return None, None
    with open(files[-1]) as f:
        return json.load(f), os.path.relpath(files[-1], ROOT)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------
# clocks

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's own Decoder (oracle/_ref, compiled unmodified from /root/reference), else the port

def _cpu_worker(args):
    kind, cfg, stream_idx, frames = args
    import oracle_py as O
    st = pyapi.SynthStream(cfg, stream_idx, 0, 0, frames)
    seq = st.seq
    st.close()
    dec = O.CpuDecoder(kind, seq)
    O.run_stream(dec, cfg, stream_idx, 0, 0, frames)
    t = dec.sec_decode + dec.sec_deblock
    dec.close()
    return frames * seq.width_mbs * seq.height_mbs, t


def cpu_reference_run(cores, streams_per_core, frames, first_stream=0):
    """Runs `cores` processes, each reconstructing `streams_per_core` streams with the CPU reference; returns
    (macroblocks, seconds) where seconds = the slowest worker's time inside Decoder::decode + deblock_filter."""
    import oracle_py as O
    kind = "ref" if os.path.exists(O.REF_PATH) else "port"
    jobs = [(kind, CONFIG_ID, first_stream + i, frames) for i in range(cores * streams_per_core)]
    per_worker = [0.0] * cores
    total_mb = 0
    import multiprocessing as mp
    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("spawn")) as ex:
        for i, (mb, t) in enumerate(ex.map(_cpu_worker, jobs, chunksize=streams_per_core)):
            total_mb += mb
            per_worker[i // streams_per_core] += t
    return kind, total_mb, max(per_worker)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames = args.frames                      # the same GOPs as the GPU arm (16 pictures: 1 I + 5 P + 10 B)
    st = pyapi.SynthStream(CONFIG_ID, 0, 0, 0, frames)
    nmb = st.nmb
    st.close()
    times, kind, mbs = [], "ref", 0
    for it in range(args.warmup + args.steps):
        kind, mbs, sec = cpu_reference_run(cores, 1, frames, first_stream=it * cores)
        if it >= args.warmup:
            times.append(sec)
    sec = sum(times) / len(times)
    value = mbs / sec
    sample = (f"each step = {cores} streams x {frames} pictures of that workload (a bounded sample of its streams, the same "
              "picture mix, fresh streams every step), one process per host core; timed: the reference Decoder's coeff_* / "
              "decode / deblock_filter calls including the harness's per-MB hand-over of the description (oracle/ref_harness.cc)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32/u8", "data": "synthetic",
            "config": {"workload": workload_name(args.streams, args.frames), "streams_per_gpu": args.streams,
                       "pictures_per_step_per_gpu": args.streams * args.frames,
                       "macroblocks_per_step_per_gpu": args.streams * args.frames * nmb, "sample": sample,
                       "same_pictures_as_gpu_arm": True},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "ref" else "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU arm

def generate_stream(args):
    cfg, stream_idx, frames = args
    st = pyapi.SynthStream(cfg, stream_idx, 0, 0, frames)
    pics = [p for p in st]
    st.close()
    return pics


class PackedPicture:
    """One generated picture in the form a parser thread leaves in the staging: [mbs | slices] and the stream words."""
    __slots__ = ("head", "stream", "words", "pp", "info", "acct")


def pack_picture(pic, nmb):
    L = pyapi.recon_lib()
    pp = pic.pp
    head = C.create_string_buffer(C.sizeof(pyapi.Mb) * nmb + C.sizeof(pyapi.Slice) * pp.num_slices)
    cap = pic.info.num_levels + 48 * nmb
    stream = (C.c_uint32 * cap)()
    words = L.h264r_pack_picture(nmb, pic.mbs, pic.motion, pic.levels, pic.info.num_levels, head, stream, cap)
    if words < 0:
        raise SystemExit(f"bench.py: h264r_pack_picture failed ({words})")
    C.memmove(C.addressof(head) + C.sizeof(pyapi.Mb) * nmb, pic.slices, C.sizeof(pyapi.Slice) * pp.num_slices)
    a = (C.c_uint64 * 8)()
    pyapi.synth_lib().h264s_account(pic.mbs, pic.slices, nmb, pp.run_deblock, a)
    out = PackedPicture()
    out.head, out.stream, out.words, out.pp, out.info, out.acct = head, stream, int(words), pp, pic.info, list(a)
    return out


def generate_and_pack(args):
    cfg, stream_idx, frames, nmb = args
    return [pack_picture(p, nmb) for p in generate_stream((cfg, stream_idx, frames))]


def oracle_digests(cfg, stream_idx, frames):
    """Checker only (never timed, never on the product path): per-picture md5 of one stream from the CPU restatement."""
    import oracle_py as O
    st = pyapi.SynthStream(cfg, stream_idx, 0, 0, frames)
    seq = st.seq
    st.close()
    dec = O.CpuDecoder("port", seq)
    d = O.run_stream(dec, cfg, stream_idx, 0, 0, frames)
    dec.close()
    return d


def run_gpu_arm(args, rank, world, local_rank, dist):
    import hashlib
    lib = pyapi.recon_lib()
    if lib.h264r_device_count() <= local_rank:
        raise SystemExit("bench.py: no CUDA device for this rank; the engine has no CPU fallback")
    streams, frames = args.streams, args.frames
    st = pyapi.SynthStream(CONFIG_ID, 0, 0, 0, frames)
    seq, nmb, frames = st.seq, st.nmb, st.num_frames
    st.close()
    npics = streams * frames
    cores = os.cpu_count() or 4
    feeders = args.feeders or max(2, min(6, (cores - 2) // max(1, min(world, 8))))   # measured: 4-6 feeders saturate the host memory next to the two DMA directions
    max_levels = args.max_levels or nmb * 96
    eng = pyapi.Engine(seq, device=local_rank, max_frames=npics, max_pictures=npics, max_slices=4, max_levels=max_levels)
    kernels = eng.kernel_names()

    # ---- generate + pack (threads; the C helpers release the GIL) ----
    t0 = time.time()
    my_streams = pyapi.streams_of_rank(rank, world, streams)
    if args.round1_mix:                               # diagnostic: only streams with direct_8x8_inference_flag = 1, as in round 1
        my_streams = [s for s in range(rank * 2 * streams, (rank + 1) * 2 * streams) if s % 3 != 2][:streams]
    with ThreadPoolExecutor(max_workers=min(32, cores)) as ex:
        all_pics = list(ex.map(generate_and_pack, [(CONFIG_ID, sid, frames, nmb) for sid in my_streams]))
    fbytes = eng.w * eng.h * 3 // 2
    out_host = eng.host_alloc(fbytes * npics)          # pinned destination of every reconstructed frame
    table = (pyapi.BenchPicture * npics)()
    acct = [0] * 8
    total_levels = 0
    pics_by_type = {"P": 0, "B": 0, "I": 0}
    frames_of = [dict() for _ in range(streams)]
    frame_of_pic = {}
    k = 0
    for i in range(frames):                             # picture i of every stream, decode order
        for s in range(streams):
            pk = all_pics[s][i]
            dst = eng.frame_alloc()
            frames_of[s][pk.info.pic_index] = dst
            frame_of_pic[(s, pk.info.pic_index)] = (dst, k)
            pics_by_type["PBI"[pk.info.pic_type]] += 1
            e = table[k]
            C.memmove(C.byref(e.pp), C.byref(pk.pp), C.sizeof(pyapi.PicParams))
            for r in range(pk.info.num_refs):
                e.pp.ref_frames[r] = frames_of[s][pk.info.ref_pic_index[r]]
            e.dst, e.stream_id = dst, s
            e.head, e.stream, e.stream_words = C.addressof(pk.head), C.addressof(pk.stream), pk.words
            e.pitch_y, e.out = eng.w, out_host + k * fbytes
            acct = [x + y for x, y in zip(acct, pk.acct)]
            total_levels += pk.info.num_levels
            k += 1
    gen_s = time.time() - t0
    total_mb = npics * nmb

    def feed(steps, flush_every):
        fill, flush = C.c_double(), C.c_double()
        t = lib.h264r_bench_feed(eng.ctx, table, npics, nmb, feeders, flush_every, steps, C.byref(fill), C.byref(flush))
        if t < 0:
            raise SystemExit(f"bench.py: h264r_bench_feed failed: {lib.h264r_strerror(int(t)).decode()} "
                             f"{lib.h264r_last_cuda_error(eng.ctx).decode()}")
        return t, fill.value, flush.value

    def barrier():
        eng.wait()
        if dist is not None:
            dist.barrier()

    def check_parity(sampled):
        """md5 of every frame of the sampled streams, as they lie in the pinned output buffer, against the oracle."""
        ok = True
        for s, want in sampled.items():
            for idx in range(frames):
                _dst, kk = frame_of_pic[(s, idx)]
                got = hashlib.md5(C.string_at(out_host + kk * fbytes, fbytes)).hexdigest()
                ok = ok and got == want[idx]
        return ok

    # the checker's answer for the sampled streams (CPU restatement; outside every timed region)
    sampled = {}
    if not args.no_parity_check:
        for s in sorted({0, 2, streams - 1} if not args.round1_mix else {0}):            # local stream 2 of every rank has direct_8x8_inference_flag = 0 (64 % 3 == 1: rank r starts at 64 r)
            if 0 <= s < streams:
                sampled[s] = oracle_digests(CONFIG_ID, my_streams[s], frames)

    # clocks and throttle reasons: sampled every 100 ms by rank 0 from here (first flush, warm-up) to the end of the
    # end-to-end region -- the GPU is busy throughout
    clk = ClockSampler(local_rank)
    if rank == 0:
        clk.start()
    # first run through the public entry points, one flush for everything (the wave structure the replays repeat);
    # afterwards the descriptions are HBM-resident
    feed(1, npics)
    parity = {"first_flush": check_parity(sampled)} if sampled else {}

    # ---- warm-up, then K timed steps on HBM-resident inputs ----
    for _ in range(args.warmup):
        eng.replay(1, 0)
    barrier()
    s0 = eng.stats()
    t_start = time.perf_counter()
    # the K steps are enqueued back to back (one call, no host round trip between steps -- a streaming decoder does not
    # stop between GOPs either) and joined once; CUDA events on the compute stream bracket exactly these K steps
    ms, _n = eng.replay(args.steps, 0)
    ev_ms = ms[0]
    eng.wait()
    t_dev = time.perf_counter() - t_start
    s1 = eng.stats()
    launches = int(s1.kernel_launches - s0.kernel_launches)
    if sampled:                                          # the frames the timed replays left in HBM
        ysz = eng.w * eng.h
        for s in sampled:
            for idx in range(frames):
                dst, kk = frame_of_pic[(s, idx)]
                base = out_host + kk * fbytes
                C.memset(base, 0, fbytes)
                eng.download_async(dst, base, base + ysz, base + ysz + ysz // 4)
        eng.wait()
        parity["after_timed_kernels"] = check_parity(sampled)
    barrier()

    # ---- per-kernel times (separate pass: the events sit between kernels) ----
    kms, kn = eng.replay(1, pyapi.Engine.REPLAY_TIME_KERNELS)

    # ---- K timed steps end to end through the public entry points ----
    flush_every = args.flush_every or min(npics, 4 * streams)      # four pictures of every stream per h264r_flush (measured best: 256)
    for _ in range(max(1, args.warmup // 2)):
        feed(1, flush_every)
    if sampled:
        for s in sampled:
            for idx in range(frames):
                C.memset(out_host + frame_of_pic[(s, idx)][1] * fbytes, 0, fbytes)
    barrier()
    s2 = eng.stats()
    t_e2e, fill_s, flush_s = feed(args.steps, flush_every)
    s3 = eng.stats()
    if sampled:
        parity["after_timed_e2e"] = check_parity(sampled)
    barrier()
    clocks = clk.stop()
    if rank == 0 and clocks["sm_mhz"] is None:
        # no sample landed in the window (very short runs): sample a dedicated untimed replay of about two seconds
        clk = ClockSampler(local_rank)
        clk.start()
        t_end = time.perf_counter() + 2.0
        while time.perf_counter() < t_end:
            eng.replay(2, 0)
        clocks = clk.stop()

    h2d_step = int((s3.h2d_bytes - s2.h2d_bytes) // args.steps)
    d2h_step = int((s3.d2h_bytes - s2.d2h_bytes) // args.steps)

    # ---- the box's own copy ceiling for the same byte counts (plain pinned cudaMemcpyAsync, both directions at once) ----
    ceiling = None
    if not args.no_ceiling:
        barrier()
        g = (C.c_double * 3)()
        rc = lib.h264r_bench_copy_ceiling(local_rank, h2d_step, d2h_step, min(fbytes, 4 << 20), 3, g)
        if rc == 0:
            ceiling = [g[0], g[1], g[2]]
        barrier()

    if args.diag:
        def timed(fn, n=3):
            eng.wait(); t0 = time.perf_counter()
            for _ in range(n): fn()
            eng.wait(); return (time.perf_counter() - t0) / n * 1e3
        E2E = pyapi.Engine.REPLAY_H2D | pyapi.Engine.REPLAY_ASYNC
        print(f"[diag] kernels {timed(lambda: eng.replay(1, 0)):.1f} ms | h2d+kernels (replay) {timed(lambda: eng.replay(1, pyapi.Engine.REPLAY_H2D)):.1f} | "
              f"pipelined h2d+kernels (replay) {timed(lambda: eng.replay(1, E2E), 4):.1f}", file=sys.stderr)
        for fe in (npics, max(streams, npics // 4)):
            for th in sorted({2, 4, 6, 8, 12, feeders}):
                saved = feeders
                feeders = th
                t, f, fl = feed(2, fe)
                feeders = saved
                print(f"[diag] feed: {th} feeders, flush every {fe}: {t / 2 * 1e3:.1f} ms/step, host fill {f / 2 * 1e3:.1f} ms (sum), flush+download calls {fl / 2 * 1e3:.1f} ms", file=sys.stderr)

    if dist is not None:
        import torch
        t = torch.tensor([t_dev, t_e2e, ceiling[2] if ceiling else 0.0], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(t[0]), float(t[1])
        if ceiling:
            ceiling[2] = float(t[2])
        ok = torch.tensor([1.0 if all(parity.values()) else 0.0], dtype=torch.float64)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        parity["all_ranks"] = bool(ok[0] > 0.5)
    if rank != 0:
        eng.host_free(out_host)
        eng.close()
        return

    value = world * total_mb * args.steps / t_dev
    e2e_value = world * total_mb * args.steps / t_e2e
    peak, peak_src = peaks()
    # short names in the engine's kernel-kind order (h264r_bench_kernel_name)
    short = {"residual_kernel": "residual", "recon_inter2_kernel": "inter", "recon_intra_kernel + recon_intra_sparse_kernel": "intra",
             "deblock_prep_kernel": "deblock_prep", "deblock_kernel": "deblock", "deblock4_kernel": "deblock", "intra_list_kernel": "intra_list"}
    names = [short[k] for k in kernels]
    # algorithmic bytes of each kernel's own pass: residual = levels in + 768 B residual plane out per coded MB;
    # inter / intra = SURVEY 8d formula restricted to their MBs; deblock_prep = headers + motion in, 64 B out;
    # deblock: 32 + 384 read, 384 written (+ 192 motion for inter MBs); intra_list: one header word per MB in, 4 B per intra MB out
    bytes_of = {"residual": 4 * total_levels + acct[6] * (32 + 768), "inter": acct[1], "intra": acct[2],
                "deblock_prep": acct[7] * (32 + 64) + acct[4] * 192, "deblock": acct[3], "intra_list": 4 * total_mb + 4 * acct[5]}
    k_bytes = [bytes_of[n] for n in names]
    nk = len(names)
    dom = max(range(nk), key=lambda i: kms[i + 1])
    dom_ms_per_launch = kms[dom + 1] / max(1, kn[dom + 1])
    dom_bytes_per_launch = k_bytes[dom] / max(1, kn[dom + 1])
    achieved = dom_bytes_per_launch / (dom_ms_per_launch * 1e-3) / 1e9 if dom_ms_per_launch > 0 else 0.0
    step_s = t_dev / args.steps
    step_gbs = acct[0] / step_s / 1e9
    # measured DRAM traffic and executed warp instructions per launch / per step: per picture and picture type from the ncu
    # capture of the same wave shapes committed under profiles/, x this workload's pictures (scaled with the picture size)
    prof, prof_src = measured_profile()
    traffic, issue = None, None
    scale = nmb / 8160.0
    if prof:
        tb = prof.get("bytes_per_picture", {}).get(names[dom])
        if tb:
            traffic = sum(tb.get(t, 0.0) * n for t, n in pics_by_type.items()) * scale / max(1, kn[dom + 1])
        ib = prof.get("warp_inst_per_picture")
        if ib and clocks["sm_mhz"]:
            inst_step = sum(sum(v.get(t, 0.0) * n for t, n in pics_by_type.items()) for v in ib.values()) * scale
            slots = 148 * 4 * clocks["sm_mhz"] * 1e6 * step_s
            issue = {"warp_inst_per_step": inst_step, "warp_inst_per_mb": inst_step / total_mb,
                     "issue_slots_per_step": slots, "frac": inst_step / slots,
                     "note": "148 SMs x 4 schedulers x sm clock x step time; instructions from the committed ncu capture"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32/u8", "data": "synthetic",
        "config": {"workload": workload_name(streams, frames),
                   "streams_per_gpu": streams, "pictures_per_step_per_gpu": npics, "macroblocks_per_step_per_gpu": total_mb,
                   "l2_policy": "inputs larger than L2 (picture descriptions + frames of one step exceed 126 MB)" if total_mb * 384 > 130e6
                                else "single stream: the working set fits the L2 (latency-bound run, reported as such)",
                   "frames_per_s": value / nmb, "output_bytes_per_s": value * 384,
                   "ms_per_picture": step_s * 1e3 / npics, "wavefront_steps_per_picture": seq.width_mbs + 2 * (seq.height_mbs - 1)},
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"]},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
                "path": "h264r_picture_begin -> memcpy into the pinned staging -> h264r_picture_submit (feeder threads) | "
                        "h264r_flush + h264r_frame_download_async (GPU thread) | h264r_wait",
                "feeder_threads": feeders, "flush_every_pictures": flush_every, "ms_per_step": t_e2e / args.steps * 1e3,
                "host_fill_ms_per_step": fill_s / args.steps * 1e3, "host_fill_ms_per_step_per_thread": fill_s / args.steps * 1e3 / feeders,
                "host_flush_ms_per_step": flush_s / args.steps * 1e3},
        "gpu_launches": launches,
        "parity_checked": bool(parity) and all(parity.values()),
        "parity": dict(parity, streams=sorted(sampled), checker="oracle/port_recon.c digests of the sampled streams, every picture"),
        "roofline": {"bound": "hbm", "kernel": kernels[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": prof_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": dom_bytes_per_launch, "ms_per_launch": dom_ms_per_launch,
                     "whole_step": {"algorithmic_bytes": acct[0], "achieved": step_gbs, "frac": step_gbs / peak,
                                    "bytes_per_mb": acct[0] / total_mb},
                     "issue_slots": issue,
                     "kernel_ms_per_step": {n: kms[i + 1] for i, n in enumerate(names)},
                     "kernel_launches_per_step": {n: kn[i + 1] for i, n in enumerate(names)},
                     "event_ms_per_step": ev_ms / args.steps},
        "mb_mix": {"inter": acct[4] / total_mb, "intra": acct[5] / total_mb, "coded": acct[6] / total_mb},
        "setup_s": gen_s,
    }
    if ceiling:
        ceil_value = world * total_mb / ceiling[2]
        line["e2e_roofline"] = {"what": "plain pinned cudaMemcpyAsync of the same bytes per step, both directions at once, no kernels "
                                        "(max over ranks when N > 1)",
                                "h2d_gbs": ceiling[0], "d2h_gbs": ceiling[1], "ms_per_step": ceiling[2] * 1e3,
                                "ceiling_value": ceil_value, "unit": UNIT, "e2e_over_ceiling": e2e_value / ceil_value}
    if world == 1 and not args.no_cpu_baseline:
        cframes = frames
        kind, mbs, sec = cpu_reference_run(cores, 1, cframes, first_stream=1000)
        line["cpu_baseline"] = {"value": mbs / sec, "unit": UNIT, "cores": cores,
                                "kind": "reference" if kind == "ref" else "port",
                                "sample": f"{cores} streams x {cframes} pictures of the same workload (same picture mix), one process "
                                          "per core; timed: the Decoder's coeff_* / decode / deblock_filter calls incl. the harness's "
                                          "per-MB hand-over"}
    print(json.dumps(line), flush=True)
    eng.host_free(out_host)
    eng.close()


def main():
    global CONFIG_ID
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=[3, 4, 5, 6],
                    help="5: 64 x 1080p streams (the metric's workload); 3: 1080p single stream; 4: 4K single stream; "
                         "6: 1080i field pictures (use --streams 64 for a throughput figure)")
    ap.add_argument("--streams", type=int, default=0, help="independent streams per GPU (default 64, or 1 for --config 3/4)")
    ap.add_argument("--frames", type=int, default=0, help="pictures per stream (default: the config's own GOP)")
    ap.add_argument("--feeders", type=int, default=0, help="feeder threads of the end-to-end path (default: (host cores - 2) / ranks, 2..6)")
    ap.add_argument("--flush-every", type=int, default=0, help="pictures per h264r_flush in the end-to-end path (default: 4 per stream)")
    ap.add_argument("--round1-mix", action="store_true", help="diagnostic: leave out the streams with direct_8x8_inference_flag = 0 (round-1 workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-ceiling", action="store_true")
    ap.add_argument("--diag", action="store_true", help="print a transfer/kernel time breakdown to stderr")
    ap.add_argument("--max-levels", type=int, default=0, help="staging capacity of one picture's level list (default 96 per MB)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    CONFIG_ID = args.config
    args.streams = args.streams or (64 if CONFIG_ID == 5 else 1)
    args.frames = args.frames or (8 if CONFIG_ID == 4 else 16)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_mod.init_process_group("gloo", rank=rank, world_size=world)
        dist = dist_mod
    run_gpu_arm(args, rank, world, local_rank, dist)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
