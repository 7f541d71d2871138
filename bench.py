#!/usr/bin/env python
"""Benchmark of the H.264 macroblock-reconstruction path (BASELINE.json metric:
"1080p MB/s reconstructed (IDCT+MC+intra+deblock) at 1/2/4/8 B200; % HBM peak").

    python bench.py --gpus N --steps K --warmup W            # our CUDA engine through the C ABI
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU Decoder on the host cores

Workload (BASELINE.json configs[4], weak scaling): 64 independent synthetic 1080p High-profile I/P/B streams of 16
pictures PER GPU (stream seeds distinct across ranks).  One step = reconstruct all of them once (1024 pictures,
8.36 M macroblocks per GPU).  `value` = macroblocks/s with the picture descriptions resident in HBM (kernels only,
CUDA-event and wall time agree); `e2e` = the same through the public C ABI with HOST buffers: per step every picture
description is copied host->device from the pinned staging and every reconstructed frame is copied back to pinned
host memory.  The synthetic generator stands in for the entropy decoder and is outside both timed regions.
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


def workload_name(streams, frames):
    return (f"{streams} independent 1080p (120x68 MB) High-profile I/P/B streams x {frames} pictures per GPU "
            "(BASELINE configs[4]; 8x8 transform, intra 8x8, bi-pred, weighted prediction, 2 slices on odd pictures)")
UNIT = "MB/s"          # MB = macroblocks; output bytes/s = value * 384
CONFIG_ID = 5


def measured_traffic():
    """DRAM bytes per picture by kernel and picture type, from the newest ncu --set full capture committed under
    profiles/ (scripts/gpu_profile.sh; bench.py never runs under a profiler itself)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), key=os.path.getmtime)
    if not files:
        return None, None
    with open(files[-1]) as f:
        return json.load(f)["bytes_per_picture"], os.path.relpath(files[-1], ROOT)


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
    kind, stream_idx, frames = args
    import oracle_py as O
    st = pyapi.SynthStream(CONFIG_ID, stream_idx, 0, 0, frames)
    seq = st.seq
    st.close()
    dec = O.CpuDecoder(kind, seq)
    O.run_stream(dec, CONFIG_ID, stream_idx, 0, 0, frames)
    t = dec.sec_decode + dec.sec_deblock
    dec.close()
    return frames * seq.width_mbs * seq.height_mbs, t


def cpu_reference_run(cores, streams_per_core, frames, first_stream=0):
    """Runs `cores` processes, each reconstructing `streams_per_core` streams with the CPU reference; returns
    (macroblocks, seconds) where seconds = the slowest worker's time inside Decoder::decode + deblock_filter."""
    import oracle_py as O
    kind = "ref" if os.path.exists(O.REF_PATH) else "port"
    jobs = [(kind, first_stream + i, frames) for i in range(cores * streams_per_core)]
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
    frames = min(args.frames, 8)
    times, kind, mbs = [], "ref", 0
    for it in range(args.warmup + args.steps):
        kind, mbs, sec = cpu_reference_run(cores, 1, frames, first_stream=it * cores)
        if it >= args.warmup:
            times.append(sec)
    sec = sum(times) / len(times)
    value = mbs / sec
    sample = (f"each step = {cores} streams x {frames} pictures of that workload (a bounded sample of its streams, fresh "
              "streams every step), one process per host core, time inside Decoder::decode + deblock_filter only")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32/u8", "data": "synthetic",
            "config": {"workload": workload_name(args.streams, args.frames), "streams_per_gpu": args.streams,
                       "pictures_per_step_per_gpu": args.streams * args.frames,
                       "macroblocks_per_step_per_gpu": args.streams * args.frames * 8160, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "ref" else "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU arm

def generate_stream(args):
    stream_idx, frames = args
    st = pyapi.SynthStream(CONFIG_ID, stream_idx, 0, 0, frames)
    pics = [p for p in st]
    st.close()
    return pics


def run_gpu_arm(args, rank, world, local_rank, dist):
    lib = pyapi.recon_lib()
    if lib.h264r_device_count() <= local_rank:
        raise SystemExit("bench.py: no CUDA device for this rank; the engine has no CPU fallback")
    streams, frames = args.streams, args.frames
    st = pyapi.SynthStream(CONFIG_ID, 0, 0, 0, frames)
    seq, nmb = st.seq, st.nmb
    st.close()
    npics = streams * frames
    eng = pyapi.Engine(seq, device=local_rank, max_frames=npics, max_pictures=npics, max_slices=4, max_levels=args.max_levels)

    # ---- generate (threads; the generator releases the GIL) and stage every picture in pinned memory ----
    t0 = time.time()
    my_streams = pyapi.streams_of_rank(rank, world, streams)
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 4)) as ex:
        all_pics = list(ex.map(generate_stream, [(sid, frames) for sid in my_streams]))
    acct = [0] * 10
    pics_by_type = {"P": 0, "B": 0, "I": 0}
    frames_of = [dict() for _ in range(streams)]
    out_frames = []
    for i in range(frames):                       # picture i of every stream, decode order
        for s in range(streams):
            pic = all_pics[s][i]
            dst = eng.frame_alloc()
            pics_by_type["PBI"[pic.info.pic_type]] += 1
            frames_of[s][pic.info.pic_index] = dst
            eng.submit(pic, dst, [frames_of[s][pic.info.ref_pic_index[k]] for k in range(pic.info.num_refs)])
            out_frames.append(dst)
            a = (C.c_uint64 * 8)()
            pyapi.synth_lib().h264s_account(pic.mbs, pic.slices, nmb, pic.pp.run_deblock, a)
            acct = [x + y for x, y in zip(acct, list(a) + [a[7] * (32 + 32) + a[4] * 192, 4 * pic.info.num_levels + a[6] * (32 + 768)])]
            all_pics[s][i] = None                 # the pinned staging now owns the data
    del all_pics
    gen_s = time.time() - t0

    # pinned destination for the reconstructed frames (e2e device->host read of every step's result)
    fbytes = eng.w * eng.h * 3 // 2
    out_host = eng.host_alloc(fbytes * npics)

    def download_all():
        ysz = eng.w * eng.h
        for k, f in enumerate(out_frames):
            base = out_host + k * fbytes
            eng.download_async(f, base, base + ysz, base + ysz + ysz // 4)

    def barrier():
        eng.wait()
        if dist is not None:
            dist.barrier()

    # clocks and throttle reasons: sampled every 100 ms by rank 0 from here (first flush, warm-up) to the end of the
    # end-to-end region -- the GPU is busy throughout, and a sampler started right at the timed region would miss short
    # runs (nvidia-smi needs several hundred ms to come up, longer with eight ranks on one box)
    clk = ClockSampler(local_rank)
    if rank == 0:
        clk.start()
    # first run = flush (uploads everything once; afterwards the descriptions are HBM-resident)
    eng.flush()
    eng.wait()
    total_mb = npics * nmb

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
    barrier()

    # ---- K timed steps end to end: H2D of every description + kernels + D2H of every frame ----
    E2E = pyapi.Engine.REPLAY_H2D | pyapi.Engine.REPLAY_ASYNC      # enqueue like h264r_flush: nothing blocks the host
    for _ in range(max(1, args.warmup // 2)):
        eng.replay(1, E2E)
        download_all()
        eng.wait()
    barrier()
    s2 = eng.stats()
    # steps are enqueued back to back like a streaming decoder would (the engine orders each wave's H2D after the
    # previous use of its staging in HBM, each frame's D2H after the wave that produced it); one join at the end
    t_start = time.perf_counter()
    for _ in range(args.steps):
        eng.replay(1, E2E)
        download_all()
    eng.wait()
    t_e2e = time.perf_counter() - t_start
    s3 = eng.stats()
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

    if args.diag:
        def timed(fn, n=3):
            eng.wait(); t0 = time.perf_counter()
            for _ in range(n): fn()
            eng.wait(); return (time.perf_counter() - t0) / n * 1e3
        def f_h2d_only(): eng.replay(1, pyapi.Engine.REPLAY_H2D)
        def f_k_d2h(): eng.replay(1, 0); download_all()
        def f_d2h_only(): download_all()
        def f_all(): eng.replay(1, E2E); download_all()
        def f_k_d2h_async(): eng.replay(1, pyapi.Engine.REPLAY_ASYNC); download_all()
        def f_h2d_k_async(): eng.replay(1, E2E)
        print(f"[diag] kernels {timed(lambda: eng.replay(1, 0)):.1f} ms | h2d+kernels {timed(f_h2d_only):.1f} | "
              f"kernels+d2h {timed(f_k_d2h):.1f} | d2h only {timed(f_d2h_only):.1f} | all {timed(f_all):.1f} | "
              f"pipelined: kernels+d2h {timed(f_k_d2h_async, 4):.1f}, h2d+kernels {timed(f_h2d_k_async, 4):.1f}, all {timed(f_all, 4):.1f}", file=sys.stderr)
        t0 = time.perf_counter(); download_all(); t_issue = (time.perf_counter() - t0) * 1e3; eng.wait()
        print(f"[diag] host time to issue {npics} async downloads: {t_issue:.1f} ms", file=sys.stderr)

    # ---- per-kernel times (separate pass: the events sit between kernels) ----
    kms, kn = eng.replay(1, pyapi.Engine.REPLAY_TIME_KERNELS)

    if dist is not None:
        import torch
        t = torch.tensor([t_dev, t_e2e], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(t[0]), float(t[1])
    if rank != 0:
        eng.host_free(out_host)
        eng.close()
        return

    value = world * total_mb * args.steps / t_dev
    e2e_value = world * total_mb * args.steps / t_e2e
    peak, peak_src = peaks()
    names = ["residual", "inter", "intra", "deblock_prep", "deblock"]
    KERNEL_NAMES = {"residual": "residual_kernel", "inter": "recon_inter2_kernel", "intra": "recon_intra_kernel + recon_intra_sparse_kernel",
                    "deblock_prep": "deblock_prep_kernel", "deblock": "deblock_kernel"}
    # algorithmic bytes of each kernel's own pass: residual = levels in + 768 B residual plane out per coded MB;
    # inter/intra = SURVEY 8d formula restricted to their MBs; deblock_prep = headers + motion in, 32 B out
    k_bytes = [acct[9], acct[1], acct[2], acct[8], acct[3]]
    dom = max(range(5), key=lambda i: kms[i + 1])
    dom_ms_per_launch = kms[dom + 1] / max(1, kn[dom + 1])
    dom_bytes_per_launch = k_bytes[dom] / max(1, kn[dom + 1])
    achieved = dom_bytes_per_launch / (dom_ms_per_launch * 1e-3) / 1e9 if dom_ms_per_launch > 0 else 0.0
    step_gbs = acct[0] / (t_dev / args.steps) / 1e9
    # measured DRAM traffic of the dominant kernel per launch: bytes per picture of each type (ncu capture of the same
    # wave shapes, committed under profiles/) x this workload's pictures / its launches; scaled with the picture size
    traffic, traffic_src = None, None
    tb, traffic_src = measured_traffic()
    tkey = {"residual": "residual", "inter": "inter", "intra": "intra", "deblock_prep": "deblock_prep", "deblock": "deblock"}[names[dom]]
    if tb and tkey in tb:
        traffic = sum(tb[tkey].get(t, 0.0) * n for t, n in pics_by_type.items()) * (nmb / 8160.0) / max(1, kn[dom + 1])
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32/u8", "data": "synthetic",
        "config": {"workload": workload_name(streams, frames),
                   "streams_per_gpu": streams, "pictures_per_step_per_gpu": npics, "macroblocks_per_step_per_gpu": total_mb,
                   "l2_policy": "inputs larger than L2 (GBs of picture descriptions + 3.2 GB of frames per step)",
                   "frames_per_s": value / nmb, "output_bytes_per_s": value * 384},
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"]},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int((s3.h2d_bytes - s2.h2d_bytes) // args.steps),
                "d2h_bytes_per_step": int((s3.d2h_bytes - s2.d2h_bytes) // args.steps)},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": KERNEL_NAMES[names[dom]], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": dom_bytes_per_launch, "ms_per_launch": dom_ms_per_launch,
                     "whole_step": {"algorithmic_bytes": acct[0], "achieved": step_gbs, "frac": step_gbs / peak,
                                    "bytes_per_mb": acct[0] / total_mb},
                     "kernel_ms_per_step": {n: kms[i + 1] for i, n in enumerate(names)},
                     "kernel_launches_per_step": {n: kn[i + 1] for i, n in enumerate(names)},
                     "event_ms_per_step": ev_ms / args.steps},
        "mb_mix": {"inter": acct[4] / total_mb, "intra": acct[5] / total_mb, "coded": acct[6] / total_mb},
        "setup_s": gen_s,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cframes = min(frames, 8)
        kind, mbs, sec = cpu_reference_run(cores, 1, cframes, first_stream=1000)
        line["cpu_baseline"] = {"value": mbs / sec, "unit": UNIT, "cores": cores,
                                "kind": "reference" if kind == "ref" else "port",
                                "sample": f"{cores} streams x {cframes} pictures of the same 1080p workload, one process per "
                                          "core, time inside Decoder::decode + deblock_filter only"}
    print(json.dumps(line), flush=True)
    eng.host_free(out_host)
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=64, help="independent streams per GPU")
    ap.add_argument("--frames", type=int, default=16, help="pictures per stream")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--diag", action="store_true", help="print a transfer/kernel time breakdown to stderr")
    ap.add_argument("--max-levels", type=int, default=8160 * 96, help="staging capacity of one picture's level list")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

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
