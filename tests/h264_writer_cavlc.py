"""Test-only H.264 Annex-B writer with REAL residual data: CAVLC Main / High profile I / P / B streams (SURVEY.md 8f items 1
and 3, second stage).

tests/h264_writer.py writes residual-free Baseline streams; this one exercises what is left of the boundary when the
reference's own parser drives the GPU engine: the coefficient hand-over `Decoder::coeff_luma_dc / coeff_luma_ac /
coeff_chroma_dc / coeff_chroma_ac` + `transform_luma_dc / transform_chroma_dc` (parser/interpret_residual.cc:155-172,
407-431, 471-477), `cbp_blks`, B slices with spatial / temporal direct prediction (parser/interpret_mv.cc:116-434), explicit and
implicit weighted prediction (parser/interpret_rbsp.cc:832-907), the 8x8 transform with Intra 8x8 and SPS / PPS scaling
lists including the fall-back rules (decoder/transform.cc:173-262), direct_8x8_inference_flag 0 and 1.

What is written is legal syntax with random content: macroblock types, prediction modes, motion-vector differences and
coefficient levels are drawn at random (levels small, so that every intermediate stays far inside 16 bits); the decoder's
own prediction turns them into vectors and samples.  Parity does not depend on WHAT the stream says, only on both decoders
reading the same thing -- the reference decoder is the judge of legality (it must decode the stream without an error), the
GPU build must then produce the same bytes.

CAVLC tables: ITU-T H.264 Tables 9-4 (coded_block_pattern), 9-5 (coeff_token), 9-7 / 9-8 / 9-9 (total_zeros), 9-10
(run_before), written as (length, code) arrays and checked for prefix-freeness when the module is imported.
"""
import random

from h264_writer import BitWriter, nal, BLK_XY, BLK_IDX

# ---- Table 9-5 coeff_token: [nC class][trailing ones][total coeff] -> (length, code) ------------------------------
_CT_LEN = [
    [[1, 6, 8, 9, 10, 11, 13, 13, 13, 14, 14, 15, 15, 16, 16, 16, 16],
     [0, 2, 6, 8, 9, 10, 11, 13, 13, 14, 14, 15, 15, 15, 16, 16, 16],
     [0, 0, 3, 7, 8, 9, 10, 11, 13, 13, 14, 14, 15, 15, 16, 16, 16],
     [0, 0, 0, 5, 6, 7, 8, 9, 10, 11, 13, 14, 14, 15, 15, 16, 16]],
    [[2, 6, 6, 7, 8, 8, 9, 11, 11, 12, 12, 12, 13, 13, 13, 14, 14],
     [0, 2, 5, 6, 6, 7, 8, 9, 11, 11, 12, 12, 13, 13, 14, 14, 14],
     [0, 0, 3, 6, 6, 7, 8, 9, 11, 11, 12, 12, 13, 13, 13, 14, 14],
     [0, 0, 0, 4, 4, 5, 6, 6, 7, 9, 11, 11, 12, 13, 13, 13, 14]],
    [[4, 6, 6, 6, 7, 7, 7, 7, 8, 8, 9, 9, 9, 10, 10, 10, 10],
     [0, 4, 5, 5, 5, 5, 6, 6, 7, 8, 8, 9, 9, 9, 10, 10, 10],
     [0, 0, 4, 5, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 10],
     [0, 0, 0, 4, 4, 4, 4, 4, 5, 6, 7, 8, 8, 9, 10, 10, 10]],
]
_CT_CODE = [
    [[1, 5, 7, 7, 7, 7, 15, 11, 8, 15, 11, 15, 11, 15, 11, 7, 4],
     [0, 1, 4, 6, 6, 6, 6, 14, 10, 14, 10, 14, 10, 1, 14, 10, 6],
     [0, 0, 1, 5, 5, 5, 5, 5, 13, 9, 13, 9, 13, 9, 13, 9, 5],
     [0, 0, 0, 3, 3, 4, 4, 4, 4, 4, 12, 12, 8, 12, 8, 12, 8]],
    [[3, 11, 7, 7, 7, 4, 7, 15, 11, 15, 11, 8, 15, 11, 7, 9, 7],
     [0, 2, 7, 10, 6, 6, 6, 6, 14, 10, 14, 10, 14, 10, 11, 8, 6],
     [0, 0, 3, 9, 5, 5, 5, 5, 13, 9, 13, 9, 13, 9, 6, 10, 5],
     [0, 0, 0, 5, 4, 6, 8, 4, 4, 4, 12, 8, 12, 12, 8, 1, 4]],
    [[15, 15, 11, 8, 15, 11, 9, 8, 15, 11, 15, 11, 8, 13, 9, 5, 1],
     [0, 14, 15, 12, 10, 8, 14, 10, 14, 14, 10, 14, 10, 7, 12, 8, 4],
     [0, 0, 13, 14, 11, 9, 13, 9, 13, 10, 13, 9, 13, 9, 11, 7, 3],
     [0, 0, 0, 12, 11, 10, 9, 8, 13, 12, 12, 12, 8, 12, 10, 6, 2]],
]
_CT_DC_LEN = [[2, 6, 6, 6, 6], [0, 1, 6, 7, 8], [0, 0, 3, 7, 8], [0, 0, 0, 6, 7]]
_CT_DC_CODE = [[1, 7, 4, 3, 2], [0, 1, 6, 3, 3], [0, 0, 1, 2, 2], [0, 0, 0, 5, 0]]
# ---- Tables 9-7 / 9-8: total_zeros for 4x4 blocks, index total_coeff - 1 ----
_TZ_LEN = [
    [1, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 9], [3, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 6, 6, 6, 6],
    [4, 3, 3, 3, 4, 4, 3, 3, 4, 5, 5, 6, 5, 6], [5, 3, 4, 4, 3, 3, 3, 4, 3, 4, 5, 5, 5],
    [4, 4, 4, 3, 3, 3, 3, 3, 4, 5, 4, 5], [6, 5, 3, 3, 3, 3, 3, 3, 4, 3, 6], [6, 5, 3, 3, 3, 2, 3, 4, 3, 6],
    [6, 4, 5, 3, 2, 2, 3, 3, 6], [6, 6, 4, 2, 2, 3, 2, 5], [5, 5, 3, 2, 2, 2, 4], [4, 4, 3, 3, 1, 3], [4, 4, 2, 1, 3],
    [3, 3, 1, 2], [2, 2, 1], [1, 1]]
_TZ_CODE = [
    [1, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 1], [7, 6, 5, 4, 3, 5, 4, 3, 2, 3, 2, 3, 2, 1, 0],
    [5, 7, 6, 5, 4, 3, 4, 3, 2, 3, 2, 1, 1, 0], [3, 7, 5, 4, 6, 5, 4, 3, 3, 2, 2, 1, 0],
    [5, 4, 3, 7, 6, 5, 4, 3, 2, 1, 1, 0], [1, 1, 7, 6, 5, 4, 3, 2, 1, 1, 0], [1, 1, 5, 4, 3, 3, 2, 1, 1, 0],
    [1, 1, 1, 3, 3, 2, 2, 1, 0], [1, 0, 1, 3, 2, 1, 1, 1], [1, 0, 1, 3, 2, 1, 1], [0, 1, 1, 2, 1, 3], [0, 1, 1, 1, 1],
    [0, 1, 1, 1], [0, 1, 1], [0, 1]]
# ---- Table 9-9 (a): total_zeros for chroma DC 2x2 ----
_TZ_DC_LEN = [[1, 2, 3, 3], [1, 2, 2], [1, 1]]
_TZ_DC_CODE = [[1, 1, 1, 0], [1, 1, 0], [1, 0]]
# ---- Table 9-10: run_before, index min(zerosLeft, 7) - 1 ----
_RB_LEN = [[1, 1], [1, 2, 2], [2, 2, 2, 2], [2, 2, 2, 3, 3], [2, 2, 3, 3, 3, 3], [2, 3, 3, 3, 3, 3, 3],
           [3, 3, 3, 3, 3, 3, 3, 4, 5, 6, 7, 8, 9, 10, 11]]
_RB_CODE = [[1, 0], [1, 1, 0], [3, 2, 1, 0], [3, 2, 1, 1, 0], [3, 2, 3, 2, 1, 0], [3, 0, 1, 3, 2, 5, 4],
            [7, 6, 5, 4, 3, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1]]
# ---- Table 9-4: coded_block_pattern -> codeNum (ChromaArrayType 1), as the inverse of the codeNum -> cbp columns ----
_CBP_INTRA = [47, 31, 15, 0, 23, 27, 29, 30, 7, 11, 13, 14, 39, 43, 45, 46, 16, 3, 5, 10, 12, 19, 21, 26, 28, 35, 37, 42, 44, 1, 2, 4, 8,
              17, 18, 20, 24, 6, 9, 22, 25, 32, 33, 34, 36, 40, 38, 41]
_CBP_INTER = [0, 16, 1, 2, 4, 8, 32, 3, 5, 10, 12, 15, 47, 7, 11, 13, 14, 6, 9, 31, 35, 37, 42, 44, 33, 34, 36, 40, 39, 43, 45, 46, 17,
              18, 20, 24, 19, 21, 26, 28, 23, 27, 29, 30, 22, 25, 38, 41]
CBP_CODENUM_INTRA = {cbp: k for k, cbp in enumerate(_CBP_INTRA)}
CBP_CODENUM_INTER = {cbp: k for k, cbp in enumerate(_CBP_INTER)}

ZIGZAG4 = [(0, 0), (1, 0), (0, 1), (0, 2), (1, 1), (2, 0), (3, 0), (2, 1), (1, 2), (0, 3), (1, 3), (2, 2), (3, 1), (3, 2), (2, 3), (3, 3)]


def _check_prefix_free(pairs, what):
    codes = sorted({(l, c) for l, c in pairs if l})
    strs = [format(c, "0%db" % l) for l, c in codes]
    assert len(set(strs)) == len(strs), what + ": duplicate code"
    for a in strs:
        for b in strs:
            assert a is b or not b.startswith(a), f"{what}: {a} is a prefix of {b}"


for _k in range(3):
    _check_prefix_free([(_CT_LEN[_k][t][n], _CT_CODE[_k][t][n]) for t in range(4) for n in range(17) if n >= t], f"coeff_token {_k}")
_check_prefix_free([(_CT_DC_LEN[t][n], _CT_DC_CODE[t][n]) for t in range(4) for n in range(5) if n >= t], "coeff_token chroma DC")
for _k in range(15):
    _check_prefix_free(zip(_TZ_LEN[_k], _TZ_CODE[_k]), f"total_zeros {_k + 1}")
for _k in range(3):
    _check_prefix_free(zip(_TZ_DC_LEN[_k], _TZ_DC_CODE[_k]), f"total_zeros chroma DC {_k + 1}")
for _k in range(7):
    _check_prefix_free(zip(_RB_LEN[_k], _RB_CODE[_k]), f"run_before {_k + 1}")
assert sorted(_CBP_INTRA) == list(range(48)) and sorted(_CBP_INTER) == list(range(48))


def write_residual_block(w, coeffs, nc, max_coeff):
    """residual_block_cavlc (7.3.5.3.2 / 9.2): coeffs = the block's levels in scan order (len == max_coeff).
    nc: -1 for chroma DC.  Returns total_coeff."""
    nz = [(i, c) for i, c in enumerate(coeffs) if c]
    total = len(nz)
    t1 = 0
    for _, c in reversed(nz):                   # trailing ones: up to three +-1 at the high-frequency end
        if abs(c) == 1 and t1 < 3:
            t1 += 1
        else:
            break
    if nc == -1:
        w.u(_CT_DC_LEN[t1][total], _CT_DC_CODE[t1][total])
    elif nc >= 8:
        w.u(6, 0b000011 if total == 0 else ((total - 1) << 2) | t1)
    else:
        k = 0 if nc < 2 else (1 if nc < 4 else 2)
        w.u(_CT_LEN[k][t1][total], _CT_CODE[k][t1][total])
    if total == 0:
        return 0
    levels = [c for _, c in reversed(nz)]       # highest frequency first
    for c in levels[:t1]:
        w.u(1, 1 if c < 0 else 0)
    suffix_len = 1 if total > 10 and t1 < 3 else 0
    for i in range(t1, total):
        c = levels[i]
        code = 2 * c - 2 if c > 0 else -2 * c - 1
        if i == t1 and t1 < 3:
            code -= 2
        if suffix_len == 0:
            if code < 14:
                w.u(code + 1, 1)
            elif code < 30:
                w.u(15, 1); w.u(4, code - 14)
            else:
                assert code - 30 < 4096
                w.u(16, 1); w.u(12, code - 30)
        else:
            if code < (15 << suffix_len):
                w.u((code >> suffix_len) + 1, 1); w.u(suffix_len, code & ((1 << suffix_len) - 1))
            else:
                assert code - (15 << suffix_len) < 4096
                w.u(16, 1); w.u(12, code - (15 << suffix_len))
        if suffix_len == 0:
            suffix_len = 1
        if abs(c) > (3 << (suffix_len - 1)) and suffix_len < 6:
            suffix_len += 1
    if total < max_coeff:
        last = nz[-1][0]
        total_zeros = last + 1 - total
        if nc == -1:
            w.u(_TZ_DC_LEN[total - 1][total_zeros], _TZ_DC_CODE[total - 1][total_zeros])
        else:
            w.u(_TZ_LEN[total - 1][total_zeros], _TZ_CODE[total - 1][total_zeros])
        zeros_left = total_zeros
        pos = [i for i, _ in reversed(nz)]
        for k in range(total - 1):
            if zeros_left <= 0:
                break
            run = pos[k] - pos[k + 1] - 1
            t = min(zeros_left, 7) - 1
            w.u(_RB_LEN[t][run], _RB_CODE[t][run])
            zeros_left -= run
    return total


# B-slice macroblock types 1..21 (Table 7-14): (partition shape 0 = 16x16, 1 = 16x8, 2 = 8x16; prediction of each partition)
L0, L1, BI = 0, 1, 2
B_MB_TYPES = {1: (0, [L0]), 2: (0, [L1]), 3: (0, [BI]),
              4: (1, [L0, L0]), 5: (2, [L0, L0]), 6: (1, [L1, L1]), 7: (2, [L1, L1]), 8: (1, [L0, L1]), 9: (2, [L0, L1]),
              10: (1, [L1, L0]), 11: (2, [L1, L0]), 12: (1, [L0, BI]), 13: (2, [L0, BI]), 14: (1, [L1, BI]), 15: (2, [L1, BI]),
              16: (1, [BI, L0]), 17: (2, [BI, L0]), 18: (1, [BI, L1]), 19: (2, [BI, L1]), 20: (1, [BI, BI]), 21: (2, [BI, BI])}
# B sub-macroblock types 1..12 (Table 7-18): (prediction, number of sub-partitions)
B_SUB_TYPES = {1: (L0, 1), 2: (L1, 1), 3: (BI, 1), 4: (L0, 2), 5: (L0, 2), 6: (L1, 2), 7: (L1, 2), 8: (BI, 2), 9: (BI, 2),
               10: (L0, 4), 11: (L1, 4), 12: (BI, 4)}


class Stream:
    LOG2_MAX_FRAME_NUM = 4
    LOG2_MAX_POC_LSB = 6

    def __init__(self, width_mbs, height_mbs, seed=1, profile="main", num_refs=2, weighted_pred=0, weighted_bipred=0,
                 direct_8x8_inference=1, transform_8x8=False, scaling=None, constrained_intra=0, chroma_qp_offset=0, field=False,
                 claim_mbaff=False):
        """profile: "main" (4x4 transform) or "high" (transform_8x8 / scaling allowed).  scaling = None, or a pair
        (sps_lists, pps_lists) where each is None (matrix not present) or a list of eight entries: None (list not present:
        fall-back rule A / B), "default" (useDefaultScalingMatrixFlag) or a list of 16 / 64 values in raster order."""
        # field=True: every picture is a field (frame_mbs_only_flag = 0, field_pic_flag = 1); field="adaptive": PAFF, picture()
        # says per picture whether it is a frame or a field.  height_mbs is the height of the FRAME and must be even then.
        # claim_mbaff: the SPS says mb_adaptive_frame_field_flag = 1 (the macroblock layer is NOT written as MBAFF: only good
        # for checking that a decoder without MBAFF support stops at the slice header)
        self.claim_mbaff = claim_mbaff
        self.interlaced = bool(field)
        self.field = field is True                    # form of the picture being written
        assert not field or height_mbs % 2 == 0
        self.Hframe, self.Hfield = height_mbs, height_mbs // 2
        self.W, self.H = width_mbs, self.Hfield if self.field else self.Hframe
        self.rng = random.Random(seed)
        self.profile = profile
        self.num_refs = num_refs
        self.weighted_pred, self.weighted_bipred = weighted_pred, weighted_bipred
        self.direct8x8 = 1 if field else direct_8x8_inference      # frame_mbs_only_flag = 0 requires direct_8x8_inference_flag = 1
        self.ref_frames = 0                           # complete reference frames in the DPB
        self.first_field_is_ref = False               # the first field of the frame being written is a reference field
        self.t8 = transform_8x8 and profile == "high"
        self.scaling = scaling if profile == "high" else None
        self.constrained_intra = constrained_intra
        self.chroma_qp_offset = chroma_qp_offset
        self.out = bytearray()
        self.frame_num = 0
        self.idr_id = 0
        self.refs_available = 0
        self._sps_pps()

    # ---- parameter sets ----
    @staticmethod
    def _scaling_list(w, entry, size):
        """scaling_list(): delta_scale in zig-zag order; entry = "default" -> first delta makes nextScale 0."""
        if entry == "default":
            w.se(-8)                                  # nextScale = (8 - 8) = 0 at j == 0: useDefaultScalingMatrixFlag
            return
        order = ZIGZAG4 if size == 16 else Stream._zigzag8()
        n = 4 if size == 16 else 8
        last = 8
        for x, y in order:
            v = entry[y * n + x]
            d = v - last
            if d > 127: d -= 256
            if d < -128: d += 256
            w.se(d)
            last = v

    @staticmethod
    def _zigzag8():
        out, x, y, up = [], 0, 0, True
        for _ in range(64):
            out.append((x, y))
            if up:
                if x == 7: y += 1; up = False
                elif y == 0: x += 1; up = False
                else: x += 1; y -= 1
            else:
                if y == 7: x += 1; up = True
                elif x == 0: y += 1; up = True
                else: x -= 1; y += 1
        return out

    def _scaling_matrix(self, w, lists, n8):
        for i in range(6 + n8):
            e = lists[i]
            w.u(1, 0 if e is None else 1)             # *_scaling_list_present_flag[i]
            if e is not None:
                self._scaling_list(w, e, 16 if i < 6 else 64)

    def _sps_pps(self):
        w = BitWriter()
        high = self.profile == "high"
        w.u(8, 100 if high else 77)
        w.u(8, 0)                                     # constraint flags
        w.u(8, 40)
        w.ue(0)
        if high:
            w.ue(1)                                   # chroma_format_idc 4:2:0
            w.ue(0); w.ue(0)                          # bit depths
            w.u(1, 0)                                 # qpprime_y_zero_transform_bypass_flag
            sps_lists = self.scaling[0] if self.scaling else None
            w.u(1, 1 if sps_lists else 0)             # seq_scaling_matrix_present_flag
            if sps_lists:
                self._scaling_matrix(w, sps_lists, 2)
        w.ue(self.LOG2_MAX_FRAME_NUM - 4)
        w.ue(0)                                       # pic_order_cnt_type
        w.ue(self.LOG2_MAX_POC_LSB - 4)
        w.ue(self.num_refs + 1)                   # max_num_ref_frames: one more than a slice lists, so that the picture a
        w.u(1, 0)                                 # co-located block of a temporal-direct MB refers to is still in the DPB
        w.ue(self.W - 1); w.ue((self.Hfield if self.interlaced else self.Hframe) - 1)   # pic_height_in_map_units: the field height when frame_mbs_only_flag = 0
        w.u(1, 0 if self.interlaced else 1)           # frame_mbs_only_flag
        if self.interlaced:
            w.u(1, 1 if self.claim_mbaff else 0)      # mb_adaptive_frame_field_flag
        w.u(1, self.direct8x8)
        w.u(1, 0)                                     # frame_cropping_flag
        w.u(1, 0)                                     # vui_parameters_present_flag
        w.trailing()
        self.out += nal(3, 7, w.payload())
        w = BitWriter()
        w.ue(0); w.ue(0)
        w.u(1, 0)                                     # CAVLC
        w.u(1, 0)
        w.ue(0)
        w.ue(self.num_refs - 1)                       # num_ref_idx_l0_default_active_minus1
        w.ue(0)                                       # num_ref_idx_l1_default_active_minus1
        w.u(1, self.weighted_pred); w.u(2, self.weighted_bipred)
        w.se(0); w.se(0); w.se(self.chroma_qp_offset)
        w.u(1, 1)                                     # deblocking_filter_control_present_flag
        w.u(1, self.constrained_intra)
        w.u(1, 0)
        if high:
            w.u(1, 1 if self.t8 else 0)               # transform_8x8_mode_flag
            pps_lists = self.scaling[1] if self.scaling else None
            w.u(1, 1 if pps_lists else 0)             # pic_scaling_matrix_present_flag
            if pps_lists:
                self._scaling_matrix(w, pps_lists, 2 if self.t8 else 0)
            w.se(self.chroma_qp_offset - 1 if self.chroma_qp_offset else 2)      # second_chroma_qp_index_offset
        w.trailing()
        self.out += nal(3, 8, w.payload())

    # ---- picture state ----
    def _new_picture(self):
        n = self.W * self.H
        self.slice_of = [-1] * n
        self.kind = ["none"] * n                      # "i4", "i8", "i16", "pcm", "inter", "skip"
        self.i4modes = [[2] * 16 for _ in range(n)]   # Intra4x4PredMode per 4x4 block (Intra8x8: replicated), luma4x4BlkIdx order
        self.tc_luma = [[0] * 16 for _ in range(n)]   # total_coeff per 4x4 block, raster (by * 4 + bx)
        self.tc_chroma = [[[0] * 4 for _ in range(2)] for _ in range(n)]

    def _mb_avail(self, addr, cur):
        return 0 <= addr < cur and self.slice_of[addr] == self.slice_of[cur]

    def _neighbours(self, cur):
        x, y = cur % self.W, cur // self.W
        a = cur - 1 if x > 0 else -1
        b = cur - self.W if y > 0 else -1
        c = cur - self.W + 1 if y > 0 and x + 1 < self.W else -1
        d = cur - self.W - 1 if y > 0 and x > 0 else -1
        return [m if m >= 0 and self._mb_avail(m, cur) else -1 for m in (a, b, c, d)]

    def _intra_ok(self, nb, cur_is_intra=True):
        """neighbour usable for intra prediction (constrained_intra_pred: inter neighbours are not)"""
        if nb < 0:
            return False
        if self.constrained_intra and self.kind[nb] in ("inter", "skip"):
            return False
        return True

    def _nc(self, cur, bx, by, plane=None):
        """nC of a 4x4 block (9.2.1): luma (plane None, bx/by in 0..3) or chroma AC (plane 0/1, bx/by in 0..1)."""
        a, b, _, _ = self._neighbours(cur)
        size = 4 if plane is None else 2

        def tc(mb, x, y):
            return self.tc_luma[mb][y * 4 + x] if plane is None else self.tc_chroma[mb][plane][y * 2 + x]
        na = tc(cur, bx - 1, by) if bx > 0 else (tc(a, size - 1, by) if a >= 0 else None)
        nb = tc(cur, bx, by - 1) if by > 0 else (tc(b, bx, size - 1) if b >= 0 else None)
        if na is not None and nb is not None:
            return (na + nb + 1) >> 1
        return na if na is not None else (nb if nb is not None else 0)

    # ---- residual ----
    def _level(self):
        r = self.rng.random()
        m = 1 if r < 0.55 else (2 if r < 0.75 else (3 if r < 0.85 else (self.rng.randint(4, 9) if r < 0.97 else self.rng.randint(10, 60))))
        return -m if self.rng.random() < 0.5 else m

    def _random_block(self, n, start=0, density=None):
        """levels of one block in scan order: mostly low frequencies"""
        c = [0] * n
        k = self.rng.choice([0, 1, 1, 2, 3, 5, 8]) if density is None else density
        for _ in range(k):
            pos = min(n - 1, start + int(abs(self.rng.gauss(0, n / 5.0))))
            c[pos] = self._level()
        if self.rng.random() < 0.03:                  # now and then a dense block (total_coeff > 10, all three VLC tables)
            for i in range(start, n):
                if self.rng.random() < 0.8:
                    c[i] = self._level()
        return c

    def _write_residual(self, w, cur, i16, cbp_luma, cbp_chroma, t8):
        if i16:
            dc = self._random_block(16, 0, self.rng.choice([0, 1, 2, 4]))
            write_residual_block(w, dc, self._nc(cur, 0, 0), 16)
        for i8 in range(4):
            coded = (cbp_luma >> i8) & 1
            blocks4 = None
            if coded and t8:                          # one 8x8 block = 64 levels, interleaved into four 4x4 blocks (CAVLC)
                lv = self._random_block(64)
                blocks4 = [[lv[4 * i + k] for i in range(16)] for k in range(4)]
            for i4 in range(4):
                bx, by = BLK_XY[i8 * 4 + i4]
                if not coded:
                    self.tc_luma[cur][by * 4 + bx] = 0
                    continue
                nc = self._nc(cur, bx, by)
                if i16:
                    tcn = write_residual_block(w, self._random_block(15), nc, 15)
                elif t8:
                    tcn = write_residual_block(w, blocks4[i4], nc, 16)
                else:
                    tcn = write_residual_block(w, self._random_block(16), nc, 16)
                self.tc_luma[cur][by * 4 + bx] = tcn
        if cbp_chroma:
            for pl in range(2):
                write_residual_block(w, self._random_block(4, 0, self.rng.choice([0, 1, 2])), -1, 4)
        for pl in range(2):
            for blk in range(4):
                if cbp_chroma == 2:
                    nc = self._nc(cur, blk & 1, blk >> 1, pl)
                    self.tc_chroma[cur][pl][blk] = write_residual_block(w, self._random_block(15), nc, 15)
                else:
                    self.tc_chroma[cur][pl][blk] = 0

    def _qp_delta(self, w):
        d = self.rng.choice([d for d in (0, 0, 0, 1, -1, 2, -2, 3) if 14 <= self.qp_running + d <= 44])
        self.qp_running += d
        w.se(d)

    # ---- intra macroblocks ----
    def _chroma_mode(self, a, b, d):
        legal = [0]
        if self._intra_ok(a): legal.append(1)
        if self._intra_ok(b): legal.append(2)
        if self._intra_ok(a) and self._intra_ok(b) and self._intra_ok(d): legal.append(3)
        return self.rng.choice(legal)

    def _mb_i16(self, w, cur, base):
        a, b, _, d = self._neighbours(cur)
        legal = [2]
        if self._intra_ok(b): legal.append(0)
        if self._intra_ok(a): legal.append(1)
        if self._intra_ok(a) and self._intra_ok(b) and self._intra_ok(d): legal.append(3)
        mode = self.rng.choice(legal)
        cbp_luma = self.rng.choice([0, 15])
        cbp_chroma = self.rng.randrange(3)
        w.ue(base + 1 + mode + 4 * cbp_chroma + (12 if cbp_luma else 0))
        w.ue(self._chroma_mode(a, b, d))
        self._qp_delta(w)
        self.kind[cur] = "i16"
        self._write_residual(w, cur, True, cbp_luma, cbp_chroma, False)

    def _pred_mode_of(self, cur, nbx, nby, nb_mb):
        """Intra4x4/8x8PredMode of the 4x4 block at (nbx, nby) relative to MB cur, for the predIntraNxNPredMode rule (8.3.1.1 /
        8.3.2.1): None = not available (dcPredModePredictedFlag), 2 for neighbours that are not Intra NxN."""
        if 0 <= nbx < 4 and 0 <= nby < 4:
            return self.i4modes[cur][BLK_IDX[(nbx, nby)]]
        if nb_mb < 0 or not self._intra_ok(nb_mb):
            return None
        if self.kind[nb_mb] not in ("i4", "i8"):
            return 2
        return self.i4modes[nb_mb][BLK_IDX[(nbx % 4, nby % 4)]]

    def _mb_inxn(self, w, cur, base, use8):
        a, b, c, d = self._neighbours(cur)
        w.ue(base)                                    # I_NxN
        if self.t8:
            w.u(1, 1 if use8 else 0)                  # transform_size_8x8_flag
        self.kind[cur] = "i8" if use8 else "i4"
        step = 2 if use8 else 1
        order = [0, 4, 8, 12] if use8 else range(16)  # luma4x4BlkIdx of the top-left 4x4 block of every 8x8 / 4x4 block
        for k in order:
            bx, by = BLK_XY[k]
            av_a = bx > 0 or self._intra_ok(a)
            av_b = by > 0 or self._intra_ok(b)
            av_d = (bx > 0 and by > 0) or (bx > 0 and by == 0 and self._intra_ok(b)) or (bx == 0 and by > 0 and self._intra_ok(a)) or \
                   (bx == 0 and by == 0 and self._intra_ok(d))
            legal = [2]
            if av_b: legal += [0, 3, 7]
            if av_a: legal += [1, 8]
            if av_a and av_b and av_d: legal += [4, 5, 6]
            mode = self.rng.choice(legal)
            # the block to the left / above (for 8x8 blocks: the one beside the top-left 4x4 block)
            ma = self._pred_mode_of(cur, bx - 1, by, a)
            mb_ = self._pred_mode_of(cur, bx, by - 1, b)
            pred = 2 if ma is None or mb_ is None else min(ma, mb_)
            if mode == pred:
                w.u(1, 1)
            else:
                w.u(1, 0)
                w.u(3, mode if mode < pred else mode - 1)
            for dy in range(step):
                for dx in range(step):
                    self.i4modes[cur][BLK_IDX[(bx + dx, by + dy)]] = mode
        w.ue(self._chroma_mode(a, b, d))
        cbp_luma = self.rng.choice([0, 15, self.rng.randrange(16), self.rng.randrange(16)])
        cbp_chroma = self.rng.randrange(3)
        w.ue(CBP_CODENUM_INTRA[cbp_luma | cbp_chroma << 4])
        if cbp_luma or cbp_chroma:
            self._qp_delta(w)
            self._write_residual(w, cur, False, cbp_luma, cbp_chroma, use8)

    def _mb_pcm(self, w, cur, base):
        w.ue(base + 25)
        w.align_zero()
        w.bytes_raw(bytes(self.rng.getrandbits(8) for _ in range(384)))
        self.kind[cur] = "pcm"
        self.tc_luma[cur] = [16] * 16
        self.tc_chroma[cur] = [[16] * 4, [16] * 4]

    def _mb_intra(self, w, cur, base):
        r = self.rng.random()
        if r < 0.04:
            self._mb_pcm(w, cur, base)
        elif r < 0.4:
            self._mb_i16(w, cur, base)
        else:
            self._mb_inxn(w, cur, base, self.t8 and self.rng.random() < 0.5)

    # ---- inter macroblocks ----
    def _mvd(self, w):
        big = self.rng.random() < 0.05
        for _ in range(2):
            w.se(self.rng.randint(-40, 40) if big else self.rng.randint(-9, 9))

    def _ref(self, w, n):
        if n > 1:
            w.te(self.rng.randrange(n), n - 1)

    def _inter_tail(self, w, cur, all_8x8_or_larger):
        """coded_block_pattern, transform_size_8x8_flag, mb_qp_delta, residual of an inter MB"""
        cbp_luma = self.rng.choice([0, 0, 15, self.rng.randrange(16), self.rng.randrange(16)])
        cbp_chroma = self.rng.randrange(3)
        w.ue(CBP_CODENUM_INTER[cbp_luma | cbp_chroma << 4])
        use8 = False
        if cbp_luma and self.t8 and all_8x8_or_larger:
            use8 = self.rng.random() < 0.5
            w.u(1, 1 if use8 else 0)
        self.kind[cur] = "inter"
        if cbp_luma or cbp_chroma:
            self._qp_delta(w)
            self._write_residual(w, cur, False, cbp_luma, cbp_chroma, use8)

    def _mb_inter_p(self, w, cur, n0):
        t = self.rng.choice([0, 0, 1, 2, 3, 3])
        w.ue(t)
        big = True
        if t == 0:
            self._ref(w, n0); self._mvd(w)
        elif t in (1, 2):
            self._ref(w, n0); self._ref(w, n0)
            self._mvd(w); self._mvd(w)
        else:
            sub = [self.rng.randrange(4) for _ in range(4)]
            for s in sub: w.ue(s)
            for _ in range(4): self._ref(w, n0)
            for s in sub:
                for _ in range((1, 2, 2, 4)[s]): self._mvd(w)
            big = all(s == 0 for s in sub)
        self._inter_tail(w, cur, big)

    def _mb_inter_b(self, w, cur, n0, n1):
        r = self.rng.random()
        if r < 0.15:                                  # B_Direct_16x16 (with residual; the skipped form is B_Skip)
            w.ue(0)
            self._inter_tail(w, cur, bool(self.direct8x8))
            return
        if r < 0.7:
            t = self.rng.randint(1, 21)
            w.ue(t)
            _, preds = B_MB_TYPES[t]
            for p in preds:
                if p != L1: self._ref(w, n0)
            for p in preds:
                if p != L0: self._ref(w, n1)
            for p in preds:
                if p != L1: self._mvd(w)
            for p in preds:
                if p != L0: self._mvd(w)
            self._inter_tail(w, cur, True)
            return
        w.ue(22)                                      # B_8x8
        sub = [self.rng.choice([0, 0, 1, 2, 3, self.rng.randint(4, 12)]) for _ in range(4)]
        for s in sub: w.ue(s)
        for s in sub:
            if s and B_SUB_TYPES[s][0] != L1: self._ref(w, n0)
        for s in sub:
            if s and B_SUB_TYPES[s][0] != L0: self._ref(w, n1)
        for s in sub:
            if s and B_SUB_TYPES[s][0] != L1:
                for _ in range(B_SUB_TYPES[s][1]): self._mvd(w)
        for s in sub:
            if s and B_SUB_TYPES[s][0] != L0:
                for _ in range(B_SUB_TYPES[s][1]): self._mvd(w)
        big = all((s == 0 and self.direct8x8) or (s and B_SUB_TYPES[s][1] == 1) for s in sub)
        self._inter_tail(w, cur, big)

    # ---- slices / pictures ----
    def _pred_weight_table(self, w, n0, n1, is_b):
        ld, cd = self.rng.randint(0, 6), self.rng.randint(0, 6)
        w.ue(ld); w.ue(cd)
        for n in ([n0, n1] if is_b else [n0]):
            for _ in range(n):
                f = self.rng.random() < 0.75
                w.u(1, 1 if f else 0)
                if f:
                    w.se(self.rng.randint(-(1 << ld), (1 << ld) + (1 << ld) // 2 + 1)); w.se(self.rng.randint(-12, 12))
                f = self.rng.random() < 0.75
                w.u(1, 1 if f else 0)
                if f:
                    for _ in range(2):
                        w.se(self.rng.randint(-(1 << cd), (1 << cd) + (1 << cd) // 2 + 1)); w.se(self.rng.randint(-12, 12))

    def _slice_header(self, w, first_mb, kind, is_ref, poc, qp, idc, off_a, off_b, bottom=False):
        """kind: "idr", "i", "p", "b".  Returns (n0, n1) = active reference counts."""
        w.ue(first_mb)
        w.ue({"idr": 2, "i": 2, "p": 0, "b": 1}[kind])
        w.ue(0)
        w.u(self.LOG2_MAX_FRAME_NUM, self.frame_num % (1 << self.LOG2_MAX_FRAME_NUM))
        if self.interlaced:
            w.u(1, 1 if self.field else 0)            # field_pic_flag
            if self.field:
                w.u(1, 1 if bottom else 0)            # bottom_field_flag
        if kind == "idr":
            w.ue(self.idr_id)
        w.u(self.LOG2_MAX_POC_LSB, poc % (1 << self.LOG2_MAX_POC_LSB))
        n0 = n1 = 0
        if kind == "b":
            w.u(1, self.direct_spatial)               # direct_spatial_mv_pred_flag
        if kind in ("p", "b"):
            avail = 2 * self.ref_frames + (1 if self.first_field_is_ref else 0) if self.field else self.ref_frames
            n0 = min(self.num_refs, avail)
            n1 = 1
            override = n0 != self.num_refs
            w.u(1, 1 if override else 0)
            if override:
                w.ue(n0 - 1)
                if kind == "b": w.ue(n1 - 1)
            w.u(1, 0)                                 # ref_pic_list_modification_flag_l0
            if kind == "b": w.u(1, 0)
            if (kind == "p" and self.weighted_pred) or (kind == "b" and self.weighted_bipred == 1):
                self._pred_weight_table(w, n0, n1, kind == "b")
        if is_ref:
            if kind == "idr":
                w.u(1, 0); w.u(1, 0)
            else:
                w.u(1, 0)                             # adaptive_ref_pic_marking_mode_flag
        w.se(qp - 26)
        w.ue(idc)
        if idc != 1:
            w.se(off_a); w.se(off_b)
        return n0, n1

    def picture(self, kind, poc, qp=30, idc=0, off_a=0, off_b=0, slices=2, intra_share=0.12, skip_share=0.2, bottom=False,
                second_field=False, field=None):
        """One picture; with field=True one FIELD: `bottom` its parity, `second_field` = it completes the frame the previous
        call started (same frame_num; frame_num moves on after the second field of a reference frame).  The second field of
        an IDR frame is a non-IDR P field (complementary reference field pair, H.264 3.30).  field (PAFF streams only): this
        picture is a field (True) or a frame (False)."""
        is_ref = kind != "b"
        if field is not None:
            assert self.interlaced
            self.field = field
        self.H = self.Hfield if self.field else self.Hframe
        if not second_field:
            self.first_field_is_ref = False
        if kind == "idr":
            self.frame_num = 0
            self.ref_frames = 0
        self._new_picture()
        # field / PAFF streams: spatial direct only (the co-located field of a temporal-direct MB may refer to a field that the
        # two entries of this writer's list 0 do not hold)
        self.direct_spatial = 1 if self.interlaced else self.rng.randrange(2)
        n = self.W * self.H
        cuts = sorted(set([0] + ([self.rng.randrange(1, n)] if slices > 1 and n > 1 else [])))
        for si, first in enumerate(cuts):
            last = cuts[si + 1] if si + 1 < len(cuts) else n
            w = BitWriter()
            n0, n1 = self._slice_header(w, first, kind, is_ref, poc, qp, idc, off_a, off_b, bottom)
            self.qp_running = qp
            skip_run = 0
            base = {"idr": 0, "i": 0, "p": 5, "b": 23}[kind]
            for cur in range(first, last):
                self.slice_of[cur] = si
                if kind in ("idr", "i"):
                    self._mb_intra(w, cur, 0)
                    continue
                r = self.rng.random()
                if r < skip_share:
                    skip_run += 1
                    self.kind[cur] = "skip"
                    continue
                w.ue(skip_run)
                skip_run = 0
                if r < skip_share + intra_share:
                    self._mb_intra(w, cur, base)
                elif kind == "p":
                    self._mb_inter_p(w, cur, n0)
                else:
                    self._mb_inter_b(w, cur, n0, n1)
            if kind in ("p", "b") and skip_run:
                w.ue(skip_run)
            w.trailing()
            self.out += nal(3 if kind == "idr" else (2 if is_ref else 0), 5 if kind == "idr" else 1, w.payload())
        if kind == "idr":
            self.idr_id += 1
        if is_ref and self.field and not second_field:
            self.first_field_is_ref = True
        if is_ref and (not self.field or second_field):
            self.ref_frames = min(self.num_refs + 1, self.ref_frames + 1)
            self.frame_num += 1

    def data(self):
        return bytes(self.out)


def make_field_stream(width_mbs=11, height_mbs=10, gops=2, seed=7, b_frames=True, unpaired_tail=None, **opts):
    """The same GOP structure with every frame coded as two fields, top first: IDR frame = I field + P field, anchor frames
    = P + P fields, B frames = B + B fields (not used for reference).  The fields of a frame share frame_num; POC = 2 x
    display index of the frame (+ 1 for the bottom field).  Returns (bytes, number of FRAMES)."""
    s = Stream(width_mbs, height_mbs, seed, field=True, **opts)
    frames = 0
    for g in range(gops):
        disp = 0
        s.picture("idr", 0, qp=28 + 2 * g, slices=1 + (g & 1))
        s.picture("p", 1, qp=29, slices=2, bottom=True, second_field=True)
        frames += 1
        for k in range(3):
            nb = 2 if b_frames else 0
            disp_anchor = disp + nb + 1
            for bottom in (False, True):
                s.picture("p", 2 * disp_anchor + bottom, qp=(26, 32, 38)[k % 3] + bottom, idc=(0, 2, 0)[k % 3], off_a=(0, 2, -4)[k % 3],
                          off_b=(2, 0, -2)[k % 3], slices=1 + (frames + bottom) % 2, bottom=bottom, second_field=bottom)
            frames += 1
            for b in range(nb):
                for bottom in (False, True):
                    s.picture("b", 2 * (disp + 1 + b) + bottom, qp=(30, 34)[b], idc=(0, 1)[(b + k + bottom) % 2], off_a=2 * b, off_b=-2 * b,
                              slices=1 + b, bottom=bottom, second_field=bottom)
                frames += 1
            disp = disp_anchor
    if unpaired_tail is not None:                    # the stream ends with one field of a frame: "top" or "bottom"
        s.picture("p", 2 * (disp + 1) + (unpaired_tail == "bottom"), qp=31, slices=2, bottom=unpaired_tail == "bottom")
        frames += 1
    return s.data(), frames


def make_paff_stream(width_mbs=11, height_mbs=10, gops=2, seed=7, b_frames=True, pattern=(1, 0, 1, 1, 0, 0, 1, 0, 0, 1), **opts):
    """Picture-adaptive frame / field coding: the GOP structure of make_stream, every frame coded as a frame (0) or as two
    fields (1) following `pattern` in decode order, so that frames reference fields and fields reference frames.  Returns
    (bytes, number of FRAMES)."""
    s = Stream(width_mbs, height_mbs, seed, field="adaptive", **opts)
    frames = 0

    def frame(kind, disp, **kw):
        nonlocal frames
        as_field = bool(pattern[frames % len(pattern)])
        if as_field:
            s.picture(kind, 2 * disp, field=True, **kw)
            kw2 = dict(kw); kw2["qp"] = kw.get("qp", 30) + 1
            s.picture("p" if kind == "idr" else kind, 2 * disp + 1, field=True, bottom=True, second_field=True, **kw2)
        else:
            s.picture(kind, 2 * disp, field=False, **kw)
        frames += 1

    for g in range(gops):
        disp = 0
        frame("idr", 0, qp=28 + 2 * g, slices=1 + (g & 1))
        for k in range(3):
            nb = 2 if b_frames else 0
            disp_anchor = disp + nb + 1
            frame("p", disp_anchor, qp=(26, 32, 38)[k % 3], idc=(0, 2, 0)[k % 3], off_a=(0, 2, -4)[k % 3], off_b=(2, 0, -2)[k % 3], slices=1 + frames % 2)
            for b in range(nb):
                frame("b", disp + 1 + b, qp=(30, 34)[b], idc=(0, 1)[(b + k) % 2], off_a=2 * b, off_b=-2 * b, slices=1 + b)
            disp = disp_anchor
    return s.data(), frames


def make_stream(width_mbs=11, height_mbs=9, gops=2, seed=7, b_frames=True, **opts):
    """IDR P [B B] P [B B] ... in decode order (display: I B B P B B P); a second IDR starts every further GOP."""
    s = Stream(width_mbs, height_mbs, seed, **opts)
    count = 0
    for g in range(gops):
        anchors = 3
        disp = 0
        s.picture("idr", 0, qp=28 + 2 * g, idc=0, slices=1 + (g & 1))
        count += 1
        for k in range(anchors):
            nb = 2 if b_frames else 0
            disp_anchor = disp + nb + 1
            i = count
            s.picture("p", 2 * disp_anchor, qp=(26, 32, 38)[k % 3], idc=(0, 2, 0)[k % 3], off_a=(0, 2, -4)[k % 3], off_b=(2, 0, -2)[k % 3],
                      slices=2 if i % 2 else 1)
            count += 1
            for b in range(nb):
                s.picture("b", 2 * (disp + 1 + b), qp=(30, 34)[b], idc=(0, 1)[(b + k) % 2], off_a=2 * b, off_b=-2 * b, slices=1 + b)
                count += 1
            disp = disp_anchor
    return s.data(), count


if __name__ == "__main__":
    import sys
    data, n = make_stream()
    open(sys.argv[1], "wb").write(data)
    print(n, "pictures")
