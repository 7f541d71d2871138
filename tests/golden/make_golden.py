"""Generates tests/golden/ref_digests.json by running the REFERENCE's own Decoder (oracle/_ref/libh264ref.so,
compiled unmodified from /root/reference by oracle/Makefile) on the synthetic streams below.

Per-frame MD5 over Y|Cb|Cr is the reference project's own notion of parity
(script/test/model/__init__.py:119-186: digest_by_frames / compare).  Run in the build container:
    make -C oracle && python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_py as O  # noqa: E402
import pyapi  # noqa: E402

# (config, stream, width_mbs, height_mbs, frames); 0 = the config's own value
CASES = [
    (1, 0, 0, 0, 0), (1, 1, 0, 0, 8),                 # CIF Baseline, full size
    (2, 0, 20, 12, 12), (2, 1, 13, 9, 10), (2, 2, 80, 45, 4),
    (3, 0, 24, 14, 10), (3, 1, 24, 14, 10), (3, 2, 11, 7, 16), (3, 3, 120, 68, 3),
    (4, 0, 30, 17, 8), (4, 1, 17, 11, 8),
    (5, 0, 16, 10, 7), (5, 7, 16, 10, 7), (5, 63, 16, 10, 7),
    (3, 4, 1, 1, 6), (3, 5, 2, 1, 6), (3, 6, 1, 3, 6), (2, 3, 3, 2, 9),   # degenerate picture sizes
    # streams with stream % 3 == 2 carry direct_8x8_inference_flag = 0: direct sub-macroblocks / B_Skip / B_Direct_16x16
    # with motion per 4x4 block (decoder/decoder.cc:239-242); (2,2), (3,2) and (3,5) above are such streams too
    (2, 5, 20, 12, 12), (3, 8, 24, 14, 10), (4, 2, 30, 17, 8), (5, 2, 16, 10, 7), (5, 8, 16, 10, 7),
    # field pictures (field_pic_flag = 1, h264r_pic_params::structure): field scans, chroma vector offset between parities,
    # mvlimit 2 and bS 3 on horizontal MB edges; (6, 2) also has direct_8x8_inference_flag = 0; (6, 3) is a full 1080i field
    (6, 0, 24, 14, 10), (6, 1, 11, 7, 16), (6, 2, 20, 12, 12), (6, 3, 120, 34, 3),
]

if __name__ == "__main__":
    out = {"generator": "tests/golden/make_golden.py", "oracle": "reference Decoder (oracle/_ref/libh264ref.so)",
           "digest": "md5(Y|Cb|Cr) per picture in decode order", "cases": []}
    for cfg, sidx, w, h, n in CASES:
        st = pyapi.SynthStream(cfg, sidx, w, h, n)
        seq = st.seq
        st.close()
        ref = O.CpuDecoder("ref", seq)
        d = O.run_stream(ref, cfg, sidx, w, h, n)
        ref.close()
        out["cases"].append({"config": cfg, "stream": sidx, "width_mbs": w, "height_mbs": h, "frames": n, "md5": d})
        print(cfg, sidx, w, h, n, len(d))
    with open(os.path.join(HERE, "ref_digests.json"), "w") as f:
        json.dump(out, f, indent=1)
