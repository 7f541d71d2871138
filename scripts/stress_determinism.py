#!/usr/bin/env python
"""Stress check for the wavefront kernels (run on a GPU box): the bench-shaped workload (64 x 1080p streams, 7 pictures)
is replayed N times from HBM and end to end; a checksum over all 448 frames must never change, and four sampled
streams must equal the oracle.  A missing synchronisation in the mailbox protocols would show up here as a changing
checksum.  Usage: stress_determinism.py [rounds]"""
import hashlib, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "arrow-h264_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyapi, oracle_py as O

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
cfg, n, nstreams, sampled = 5, 7, 64, (3, 40)
st = pyapi.SynthStream(cfg, 0, 0, 0, n); seq = st.seq; st.close()
want = {}
for s in sampled:
    port = O.CpuDecoder("port", seq); want[s] = O.run_stream(port, cfg, s, 0, 0, n); port.close()
eng = pyapi.Engine(seq, max_frames=nstreams * n, max_pictures=nstreams * n, max_slices=4, max_levels=8160 * 96)
streams = [pyapi.SynthStream(cfg, s, 0, 0, n) for s in range(nstreams)]
frames = [dict() for _ in range(nstreams)]; order = []
for _ in range(n):
    for s, st in enumerate(streams):
        pic = st.next(); dst = eng.frame_alloc(); frames[s][pic.info.pic_index] = dst
        eng.submit(pic, dst, [frames[s][pic.info.ref_pic_index[i]] for i in range(pic.info.num_refs)])
        order.append((s, pic.info.pic_index, dst))
eng.flush(); eng.wait()

def checksum():
    c = 0
    for s, idx, dst in order:
        for plane in eng.download(dst):
            c = zlib.crc32(plane, c)
    return c

ref = checksum()
for s in sampled:
    for idx in range(n):
        d = hashlib.md5(b"".join(eng.download(frames[s][idx]))).hexdigest()
        assert d == want[s][idx], (s, idx)
bad = 0
for r in range(rounds):
    eng.replay(3, 0 if r % 2 == 0 else pyapi.Engine.REPLAY_H2D | pyapi.Engine.REPLAY_ASYNC)
    eng.wait()
    c = checksum()
    if c != ref:
        bad += 1
        print(f"round {r}: checksum changed {c:08x} != {ref:08x}")
print(f"{rounds} rounds x 3 replays, {len(order)} frames: {'DETERMINISTIC, bit-exact' if not bad else str(bad) + ' MISMATCHES'}")
sys.exit(1 if bad else 0)
