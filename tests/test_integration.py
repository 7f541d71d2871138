"""SURVEY.md 8f items 1 + 3: the reference's OWN decoder (NAL/slice parsing, CAVLC, motion-vector prediction, DPB, YUV
writer -- compiled unmodified from /root/reference by integration/Makefile) with its reconstruction classes replaced by
the binding integration/decoder_gpu.cc + libh264recon.so, against the unmodified reference decoder, on bitstreams
written by tests/h264_writer.py.  The outputs must be byte-identical.

The binaries are built where /root/reference exists (`__graft_entry__.build()`), stay out of git and travel to the
GPU box; where they are missing the tests skip (they never read /root/reference at run time)."""
import hashlib
import os
import subprocess

import pytest

import h264_writer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "integration", "_build", "ldecod_ref")
GPU = os.path.join(ROOT, "integration", "_build", "ldecod_gpu")

CASES = [(11, 9, 8, 7), (5, 4, 10, 3), (20, 15, 6, 11), (1, 1, 4, 5), (3, 7, 7, 9)]
# (width_mbs, height_mbs, frames, seed, frame cropping offsets left/right/top/bottom in units of two luma samples)
CROP_CASES = [(8, 5, 6, 21, (0, 0, 0, 4)), (6, 6, 5, 22, (1, 3, 2, 5)), (120, 68, 3, 23, (0, 0, 0, 4))]


def decode(binary, stream, workdir, name):
    src = os.path.join(workdir, name + ".264")
    out = os.path.join(workdir, name + ".yuv")
    with open(src, "wb") as f:
        f.write(stream)
    r = subprocess.run([binary, "-i", src, "-o", out], cwd=workdir, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    with open(out, "rb") as f:
        return f.read(), r.stdout


@pytest.mark.skipif(not os.path.exists(REF), reason="integration/_build/ldecod_ref not built")
@pytest.mark.parametrize("w,h,frames,seed", CASES)
def test_writer_streams_are_decodable_by_the_reference(tmp_path, w, h, frames, seed):
    stream = h264_writer.make_stream(w, h, frames, seed)
    yuv, log = decode(REF, stream, str(tmp_path), "ref")
    assert len(yuv) == frames * w * h * 384, log[-1500:]
    assert "Error" not in log and "error" not in log, log[-1500:]


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(GPU)), reason="integration/_build binaries not built")
@pytest.mark.parametrize("w,h,frames,seed", CASES)
def test_reference_decoder_on_the_gpu_engine_is_bit_exact(tmp_path, w, h, frames, seed):
    stream = h264_writer.make_stream(w, h, frames, seed)
    want, _ = decode(REF, stream, str(tmp_path), "ref")
    got, log = decode(GPU, stream, str(tmp_path), "gpu")
    assert len(got) == len(want), log[-1500:]
    fsz = w * h * 384
    for i in range(frames):
        a, b = want[i * fsz:(i + 1) * fsz], got[i * fsz:(i + 1) * fsz]
        if a != b:
            k = next(j for j in range(fsz) if a[j] != b[j])
            plane = "Y" if k < w * h * 256 else ("Cb" if k < w * h * 320 else "Cr")
            pytest.fail(f"output frame {i}: first difference in plane {plane} at byte {k} (reference {a[k]}, gpu {b[k]})")
    assert hashlib.md5(got).hexdigest() == hashlib.md5(want).hexdigest()


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(GPU)), reason="integration/_build binaries not built")
@pytest.mark.parametrize("w,h,frames,seed,crop", CROP_CASES)
def test_cropped_output_comes_straight_from_the_device_frames(tmp_path, w, h, frames, seed, crop):
    """with the device-resident DPB the only device->host copy is the display rectangle taken when the
    DPB outputs a picture (integration/output_gpu.cc in place of framebuf/output.cc).  Streams with SPS frame cropping
    (the last one is 1920x1088 coded, 1920x1080 displayed) must come out byte-identical to the reference decoder's."""
    stream = h264_writer.make_stream(w, h, frames, seed, crop=crop)
    want, log = decode(REF, stream, str(tmp_path), "ref")
    cw, ch = w * 16 - 2 * (crop[0] + crop[1]), h * 16 - 2 * (crop[2] + crop[3])
    assert len(want) == frames * cw * ch * 3 // 2, log[-1500:]
    got, log = decode(GPU, stream, str(tmp_path), "gpu")
    assert len(got) == len(want), log[-1500:]
    assert got == want


# ---- streams with real residual data, B pictures, weighted prediction, 8x8 transform, scaling lists (tests/h264_writer_cavlc.py) ----

def _lists(seed):
    import random
    rng = random.Random(seed)
    lst = lambda n: [rng.randint(4, 60) for _ in range(n)]
    sps = [lst(16), None, "default", None, lst(16), None, lst(64), None]        # fall-back rule A for lists 1, 3, 5, 7
    pps = [None, lst(16), None, "default", None, lst(16), None, "default"]      # fall-back rule B for lists 0, 2, 4, 6
    return sps, pps


RESIDUAL_CASES = {
    "main-ip":              dict(w=6, h=5, gops=1, seed=3, b_frames=False),
    "main-ipb":             dict(w=11, h=9, gops=2, seed=4),
    "main-ipb-explicit-wp": dict(w=7, h=4, gops=2, seed=5, weighted_pred=1, weighted_bipred=1),
    "main-ipb-implicit-wp-direct4x4": dict(w=5, h=5, gops=2, seed=6, weighted_bipred=2, direct_8x8_inference=0),
    "main-constrained-intra": dict(w=5, h=4, gops=1, seed=8, constrained_intra=1, chroma_qp_offset=-3),
    "high-t8":              dict(w=6, h=5, gops=2, seed=9, profile="high", transform_8x8=True),
    "high-t8-direct4x4":    dict(w=9, h=6, gops=2, seed=14, profile="high", transform_8x8=True, direct_8x8_inference=0, weighted_bipred=1),
    "high-scaling-sps-pps": dict(w=6, h=5, gops=2, seed=10, profile="high", transform_8x8=True, scaling="both"),
    "high-scaling-sps":     dict(w=5, h=4, gops=1, seed=11, profile="high", transform_8x8=True, scaling="sps"),
    "high-scaling-pps-wp":  dict(w=5, h=4, gops=1, seed=12, profile="high", transform_8x8=True, scaling="pps", weighted_pred=1, weighted_bipred=1),
    "high-20x12":           dict(w=20, h=12, gops=2, seed=13, profile="high", transform_8x8=True, weighted_bipred=2),
}


def residual_stream(case):
    import h264_writer_cavlc
    opts = dict(RESIDUAL_CASES[case])
    w, h = opts.pop("w"), opts.pop("h")
    if "scaling" in opts:
        sps, pps = _lists(opts["seed"])
        opts["scaling"] = {"both": (sps, pps), "sps": (sps, None), "pps": (None, pps)}[opts["scaling"]]
    data, frames = h264_writer_cavlc.make_stream(w, h, **opts)
    return data, frames, w, h


@pytest.mark.skipif(not os.path.exists(REF), reason="integration/_build/ldecod_ref not built")
@pytest.mark.parametrize("case", sorted(RESIDUAL_CASES))
def test_residual_streams_are_decodable_by_the_reference(tmp_path, case):
    """The reference decoder is the judge of the writer's syntax: every picture must come out, without an error message."""
    stream, frames, w, h = residual_stream(case)
    yuv, log = decode(REF, stream, str(tmp_path), "ref")
    assert len(yuv) == frames * w * h * 384, log[-1500:]
    assert "rror" not in log, log[-1500:]


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(GPU)), reason="integration/_build binaries not built")
@pytest.mark.parametrize("case", sorted(RESIDUAL_CASES))
def test_reference_parser_on_the_gpu_engine_with_residual_b_pictures_and_scaling_lists(tmp_path, case):
    """SURVEY.md 8f-1 through the real parser: CAVLC levels reach the engine through Decoder::coeff_* (interpret_residual.cc:
    155-172, 407-431, 471-477), B pictures with spatial / temporal direct and both kinds of weighted prediction, the 8x8
    transform, SPS / PPS scaling lists with fall-back rules A and B through Decoder::assign_quant_params: byte-identical output."""
    stream, frames, w, h = residual_stream(case)
    want, _ = decode(REF, stream, str(tmp_path), "ref")
    got, log = decode(GPU, stream, str(tmp_path), "gpu")
    assert len(want) == frames * w * h * 384
    assert len(got) == len(want), log[-1500:]
    fsz = w * h * 384
    for i in range(frames):
        a, b = want[i * fsz:(i + 1) * fsz], got[i * fsz:(i + 1) * fsz]
        if a != b:
            k = next(j for j in range(fsz) if a[j] != b[j])
            plane = "Y" if k < w * h * 256 else ("Cb" if k < w * h * 320 else "Cr")
            pytest.fail(f"{case}: output frame {i}: first difference in plane {plane} at byte {k} (reference {a[k]}, gpu {b[k]})")


# ---- streams coded in FIELD pictures (field_pic_flag = 1 on every picture; SURVEY.md 8f-4, the PAFF part) ----

FIELD_CASES = {
    "field-main-ip":       dict(w=6, h=6, gops=1, seed=21, b_frames=False),
    "field-main-ipb":      dict(w=11, h=10, gops=2, seed=22),
    "field-main-explicit-wp": dict(w=7, h=4, gops=2, seed=23, weighted_pred=1, weighted_bipred=1),
    "field-main-implicit-wp": dict(w=5, h=6, gops=2, seed=24, weighted_bipred=2, chroma_qp_offset=-2),
    "field-high-t8-scaling": dict(w=6, h=8, gops=2, seed=25, profile="high", transform_8x8=True, scaling="both", constrained_intra=1),
    "field-high-20x12":    dict(w=20, h=12, gops=1, seed=26, profile="high", transform_8x8=True, weighted_bipred=1),
    # the stream ends with one field of a frame: written with an empty other field (write_unpaired_field, output.cc:228-267)
    "field-unpaired-top":  dict(w=6, h=6, gops=1, seed=27, b_frames=False, unpaired_tail="top"),
    "field-unpaired-bottom": dict(w=5, h=4, gops=1, seed=28, unpaired_tail="bottom"),
}


def field_stream(case):
    import h264_writer_cavlc
    opts = dict(FIELD_CASES[case])
    w, h = opts.pop("w"), opts.pop("h")
    if "scaling" in opts:
        sps, pps = _lists(opts["seed"])
        opts["scaling"] = {"both": (sps, pps)}[opts["scaling"]]
    data, frames = h264_writer_cavlc.make_field_stream(w, h, **opts)
    return data, frames, w, h


@pytest.mark.skipif(not os.path.exists(REF), reason="integration/_build/ldecod_ref not built")
@pytest.mark.parametrize("case", sorted(FIELD_CASES))
def test_field_streams_are_decodable_by_the_reference(tmp_path, case):
    stream, frames, w, h = field_stream(case)
    yuv, log = decode(REF, stream, str(tmp_path), "ref")
    assert len(yuv) == frames * w * h * 384, log[-1500:]
    assert "rror" not in log, log[-1500:]


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(GPU)), reason="integration/_build binaries not built")
@pytest.mark.parametrize("case", sorted(FIELD_CASES))
def test_reference_parser_on_the_gpu_engine_with_field_pictures(tmp_path, case):
    """Field pictures through the real parser: the reference's slice / MB parsing and DPB (field pairs, field reference lists),
    field scans in the coefficient hand-off (transform.cc:339-386), references of both parities (chroma vector offset,
    inter_prediction.cc:352-354), the field deblocking rules (deblock.cc:86-108), both fields interleaved on output:
    byte-identical frames."""
    stream, frames, w, h = field_stream(case)
    want, _ = decode(REF, stream, str(tmp_path), "ref")
    got, log = decode(GPU, stream, str(tmp_path), "gpu")
    assert len(want) == frames * w * h * 384
    assert len(got) == len(want), log[-1500:]
    fsz = w * h * 384
    for i in range(frames):
        a, b = want[i * fsz:(i + 1) * fsz], got[i * fsz:(i + 1) * fsz]
        if a != b:
            k = next(j for j in range(fsz) if a[j] != b[j])
            plane = "Y" if k < w * h * 256 else ("Cb" if k < w * h * 320 else "Cr")
            pytest.fail(f"{case}: output frame {i}: first difference in plane {plane} at byte {k} (reference {a[k]}, gpu {b[k]})")


# ---- PAFF: frames and field pairs in one stream; references change form (dpb_split_field / dpb_combine_field on the device) ----

PAFF_CASES = {
    "paff-main-ipb":        dict(w=6, h=6, gops=2, seed=31),
    "paff-main-ip":         dict(w=11, h=10, gops=2, seed=32, b_frames=False, pattern=(0, 1, 0, 0, 1, 1, 0)),
    "paff-main-wp":         dict(w=7, h=4, gops=2, seed=33, weighted_pred=1, weighted_bipred=1, pattern=(1, 1, 0, 1, 0)),
    "paff-high-t8-scaling": dict(w=6, h=8, gops=2, seed=35, profile="high", transform_8x8=True, scaling="both", weighted_bipred=2),
    "paff-high-20x12":      dict(w=20, h=12, gops=1, seed=36, profile="high", transform_8x8=True, pattern=(0, 0, 1, 0, 1, 1)),
}


def paff_stream(case):
    import h264_writer_cavlc
    opts = dict(PAFF_CASES[case])
    w, h = opts.pop("w"), opts.pop("h")
    if "scaling" in opts:
        sps, pps = _lists(opts["seed"])
        opts["scaling"] = {"both": (sps, pps)}[opts["scaling"]]
    data, frames = h264_writer_cavlc.make_paff_stream(w, h, **opts)
    return data, frames, w, h


@pytest.mark.skipif(not os.path.exists(REF), reason="integration/_build/ldecod_ref not built")
@pytest.mark.parametrize("case", sorted(PAFF_CASES))
def test_paff_streams_are_decodable_by_the_reference(tmp_path, case):
    stream, frames, w, h = paff_stream(case)
    yuv, log = decode(REF, stream, str(tmp_path), "ref")
    assert len(yuv) == frames * w * h * 384, log[-1500:]
    assert "rror" not in log, log[-1500:]


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(GPU)), reason="integration/_build binaries not built")
@pytest.mark.parametrize("case", sorted(PAFF_CASES))
def test_reference_parser_on_the_gpu_engine_with_paff_streams(tmp_path, case):
    """Picture-adaptive frame / field coding through the real parser: frame pictures live in one engine context, field
    pictures in a second one of half the height; a reference that exists in the other form only is converted on the device
    (h264r_field_copy = dpb_split_field / dpb_combine_field_yuv, framebuf/picture.cc:408-660).  Byte-identical frames."""
    stream, frames, w, h = paff_stream(case)
    want, _ = decode(REF, stream, str(tmp_path), "ref")
    got, log = decode(GPU, stream, str(tmp_path), "gpu")
    assert len(want) == frames * w * h * 384
    assert len(got) == len(want), log[-1500:]
    fsz = w * h * 384
    for i in range(frames):
        a, b = want[i * fsz:(i + 1) * fsz], got[i * fsz:(i + 1) * fsz]
        if a != b:
            k = next(j for j in range(fsz) if a[j] != b[j])
            plane = "Y" if k < w * h * 256 else ("Cb" if k < w * h * 320 else "Cr")
            pytest.fail(f"{case}: output frame {i}: first difference in plane {plane} at byte {k} (reference {a[k]}, gpu {b[k]})")


@pytest.mark.skipif(not os.path.exists(GPU), reason="integration/_build/ldecod_gpu not built")
def test_mbaff_stream_stops_with_unsupported_and_no_cpu_fallback(tmp_path):
    """A stream whose SPS announces MBAFF is outside the supported subset: the GPU decoder must stop at the first slice with the
    engine's own 'unsupported' status (no crash, no output, no silent CPU reconstruction).  Needs no GPU: the check precedes
    every CUDA call."""
    import h264_writer_cavlc
    s = h264_writer_cavlc.Stream(6, 6, seed=41, field="adaptive", claim_mbaff=True)
    s.picture("idr", 0, field=False)
    src, out = str(tmp_path / "mbaff.264"), str(tmp_path / "mbaff.yuv")
    with open(src, "wb") as f:
        f.write(s.data())
    r = subprocess.run([GPU, "-i", src, "-o", out], cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "h264recon" in (r.stdout + r.stderr) and "nsupported" in (r.stdout + r.stderr), (r.stdout + r.stderr)[-800:]
    assert not os.path.exists(out) or os.path.getsize(out) == 0
