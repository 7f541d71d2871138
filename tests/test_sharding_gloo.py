"""N>1 path on CPU: two gloo ranks shard the streams exactly like bench.py does (no data-path collective), each
reconstructs its own streams (CPU oracle standing in for the GPU), and the results plus the max-over-ranks timing
are combined the way the benchmark combines them."""
import os
import sys

import pytest

import pyapi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, streams_per_gpu, q):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "arrow-h264_b200"))
    import time
    import torch
    import torch.distributed as dist
    import oracle_py as O
    import pyapi as P
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = P.streams_of_rank(rank, world, streams_per_gpu)
    dist.barrier()
    t0 = time.perf_counter()
    digests = {}
    for sid in mine:
        st = P.SynthStream(5, sid, 8, 5, 4)
        seq = st.seq
        st.close()
        dec = O.CpuDecoder("port", seq)
        digests[sid] = O.run_stream(dec, 5, sid, 8, 5, 4)
        dec.close()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, digests)
    if rank == 0:
        q.put((float(t[0]), gathered))
    dist.destroy_process_group()


def test_stream_sharding_is_a_partition():
    for world in (1, 2, 4, 8):
        ids = [s for r in range(world) for s in pyapi.streams_of_rank(r, world, 64)]
        assert ids == list(range(64 * world))
    with pytest.raises(ValueError):
        pyapi.streams_of_rank(2, 2, 64)


def test_two_gloo_ranks_match_a_single_process():
    import torch.multiprocessing as mp
    import oracle_py as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, per_gpu, port = 2, 2, 29517
    procs = [ctx.Process(target=_worker, args=(r, world, port, per_gpu, q)) for r in range(world)]
    for p in procs:
        p.start()
    t_max, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    merged = {}
    for d in gathered:
        assert not (set(d) & set(merged)), "a stream was reconstructed by two ranks"
        merged.update(d)
    assert sorted(merged) == list(range(world * per_gpu))
    assert t_max > 0
    for sid in merged:                                   # same result as one process doing everything
        st = pyapi.SynthStream(5, sid, 8, 5, 4)
        seq = st.seq
        st.close()
        dec = O.CpuDecoder("port", seq)
        assert O.run_stream(dec, 5, sid, 8, 5, 4) == merged[sid]
        dec.close()
