"""Test-only H.264 Annex-B bitstream writer.

There is no encoder and no conformance stream in the build container, so the end-to-end check of the reference's own
parser driving the GPU engine (tests/test_integration_gpu.py) needs streams made here.  The writer emits Baseline CAVLC
streams with ZERO residual, which needs none of the CAVLC coefficient tables beyond the "no coefficients" coeff_token:

  * IDR pictures: I_PCM (raw random samples), Intra16x16 (all four modes) and Intra4x4 (all nine modes) macroblocks;
  * P pictures  : P_Skip, P_L0_16x16 / 16x8 / 8x16, P_8x8 with every sub-macroblock type, two reference pictures,
                  random motion-vector differences (the decoder's own prediction turns them into vectors, many of
                  them pointing across the picture border), plus Intra16x16 / Intra4x4 macroblocks;
  * two slices per picture, disable_deblocking_filter_idc 0 / 1 / 2 and non-zero filter offsets, varying QP.

It exercises NAL/slice parsing, motion-vector prediction, reference lists and the DPB of the reference decoder and,
behind the Decoder boundary, motion compensation, intra prediction, I_PCM and the deblocking filter (the transforms are
covered by the synthetic-data parity tests).  Legality rules for the intra modes follow SURVEY.md 8a quirk 11.
"""
import random


class BitWriter:
    def __init__(self):
        self.bits = []

    def u(self, n, v):
        for i in range(n - 1, -1, -1):
            self.bits.append((v >> i) & 1)

    def ue(self, v):
        v += 1
        n = v.bit_length()
        self.u(n - 1, 0)
        self.u(n, v)

    def se(self, v):
        self.ue(2 * v - 1 if v > 0 else -2 * v)

    def te(self, v, cmax):
        if cmax > 1:
            self.ue(v)
        else:
            self.u(1, 0 if v else 1)

    def aligned(self):
        return len(self.bits) % 8 == 0

    def align_zero(self):
        while not self.aligned():
            self.bits.append(0)

    def bytes_raw(self, data):
        assert self.aligned()
        for b in data:
            self.u(8, b)

    def trailing(self):
        self.bits.append(1)
        self.align_zero()

    def payload(self):
        assert self.aligned()
        out = bytearray()
        for i in range(0, len(self.bits), 8):
            b = 0
            for k in range(8):
                b = (b << 1) | self.bits[i + k]
            out.append(b)
        return bytes(out)


def nal(ref_idc, unit_type, rbsp):
    out = bytearray(b"\x00\x00\x00\x01")
    out.append((ref_idc << 5) | unit_type)
    zeros = 0
    for b in rbsp:
        if zeros >= 2 and b <= 3:
            out.append(3)
            zeros = 0
        out.append(b)
        zeros = zeros + 1 if b == 0 else 0
    return bytes(out)


# luma4x4BlkIdx -> (bx, by) in 4x4 units
BLK_XY = [((k >> 2 & 1) * 2 + (k & 1), (k >> 3) * 2 + (k >> 1 & 1)) for k in range(16)]
BLK_IDX = {xy: k for k, xy in enumerate(BLK_XY)}


class Stream:
    LOG2_MAX_FRAME_NUM = 4
    LOG2_MAX_POC_LSB = 6

    def __init__(self, width_mbs, height_mbs, seed=1, num_refs=2, crop=None):
        self.W, self.H = width_mbs, height_mbs
        self.crop = crop                 # (left, right, top, bottom) in frame_crop_*_offset units (2 luma samples), or None
        self.rng = random.Random(seed)
        self.num_refs = num_refs
        self.out = bytearray()
        self.frame_num = 0
        self.poc = 0
        self.idr_id = 0
        self.refs_available = 0
        self._sps_pps()

    # ---- parameter sets ----
    def _sps_pps(self):
        w = BitWriter()
        w.u(8, 66)                       # profile_idc: Baseline
        w.u(8, 0b11000000)               # constraint_set0/1 flags, reserved zero bits
        w.u(8, 40)                       # level_idc
        w.ue(0)                          # seq_parameter_set_id
        w.ue(self.LOG2_MAX_FRAME_NUM - 4)
        w.ue(0)                          # pic_order_cnt_type
        w.ue(self.LOG2_MAX_POC_LSB - 4)
        w.ue(self.num_refs)              # max_num_ref_frames
        w.u(1, 0)                        # gaps_in_frame_num_value_allowed_flag
        w.ue(self.W - 1)
        w.ue(self.H - 1)
        w.u(1, 1)                        # frame_mbs_only_flag
        w.u(1, 1)                        # direct_8x8_inference_flag
        if self.crop:
            w.u(1, 1)                    # frame_cropping_flag
            for v in self.crop:
                w.ue(v)                  # frame_crop_left/right/top/bottom_offset
        else:
            w.u(1, 0)                    # frame_cropping_flag
        w.u(1, 0)                        # vui_parameters_present_flag
        w.trailing()
        self.out += nal(3, 7, w.payload())
        w = BitWriter()
        w.ue(0); w.ue(0)                 # pic_parameter_set_id, seq_parameter_set_id
        w.u(1, 0)                        # entropy_coding_mode_flag: CAVLC
        w.u(1, 0)                        # bottom_field_pic_order_in_frame_present_flag
        w.ue(0)                          # num_slice_groups_minus1
        w.ue(self.num_refs - 1)          # num_ref_idx_l0_default_active_minus1
        w.ue(0)                          # num_ref_idx_l1_default_active_minus1
        w.u(1, 0); w.u(2, 0)             # weighted_pred_flag, weighted_bipred_idc
        w.se(0); w.se(0); w.se(0)        # pic_init_qp/qs_minus26, chroma_qp_index_offset
        w.u(1, 1)                        # deblocking_filter_control_present_flag
        w.u(1, 0)                        # constrained_intra_pred_flag
        w.u(1, 0)                        # redundant_pic_cnt_present_flag
        w.trailing()
        self.out += nal(3, 8, w.payload())

    # ---- picture state ----
    def _new_picture(self):
        n = self.W * self.H
        self.slice_of = [-1] * n                         # slice index of every coded MB
        self.is_pcm = [False] * n
        self.kind = ["none"] * n                         # "i4", "i16", "pcm", "inter", "skip"
        self.i4modes = [[2] * 16 for _ in range(n)]

    def _mb_avail(self, addr, cur):
        """MB `addr` is available for MB `cur`: inside the picture, already coded, same slice."""
        return 0 <= addr < cur and self.slice_of[addr] == self.slice_of[cur]

    def _neighbours(self, cur):
        x, y = cur % self.W, cur // self.W
        a = cur - 1 if x > 0 else -1
        b = cur - self.W if y > 0 else -1
        c = cur - self.W + 1 if y > 0 and x + 1 < self.W else -1
        d = cur - self.W - 1 if y > 0 and x > 0 else -1
        return [m if m >= 0 and self._mb_avail(m, cur) else -1 for m in (a, b, c, d)]

    def _nc_dc(self, cur):
        """nC of the Intra16x16 DC block: total_coeff of the neighbouring 4x4 blocks is 16 inside I_PCM MBs, else 0."""
        a, b, _, _ = self._neighbours(cur)
        na = 16 if a >= 0 and self.is_pcm[a] else 0
        nb = 16 if b >= 0 and self.is_pcm[b] else 0
        if a >= 0 and b >= 0:
            return (na + nb + 1) >> 1
        return na if a >= 0 else (nb if b >= 0 else 0)

    @staticmethod
    def _coeff_token_zero(w, nc):
        if nc < 2:
            w.u(1, 1)
        elif nc < 4:
            w.u(2, 0b11)
        elif nc < 8:
            w.u(4, 0b1111)
        else:
            w.u(6, 0b000011)

    # ---- macroblocks ----
    def _mb_i16(self, w, cur, mb_type_base):
        a, b, _, d = self._neighbours(cur)
        legal = [2]
        if b >= 0: legal.append(0)
        if a >= 0: legal.append(1)
        if a >= 0 and b >= 0 and d >= 0: legal.append(3)
        mode = self.rng.choice(legal)
        w.ue(mb_type_base + 1 + mode)                   # cbp luma 0, cbp chroma 0
        w.ue(self._chroma_mode(a, b, d))
        delta = self.rng.choice([d for d in (0, 0, 1, -1, 2, -2) if 12 <= self.qp_running + d <= 45])
        self.qp_running += delta
        w.se(delta)                                      # mb_qp_delta
        self._coeff_token_zero(w, self._nc_dc(cur))      # Intra16x16DCLevel: no coefficients
        self.kind[cur] = "i16"

    def _chroma_mode(self, a, b, d):
        legal = [0]                                      # DC
        if a >= 0: legal.append(1)                       # horizontal
        if b >= 0: legal.append(2)                       # vertical
        if a >= 0 and b >= 0 and d >= 0: legal.append(3)
        return self.rng.choice(legal)

    def _i4_pred_mode(self, cur, k, a, b):
        """predIntra4x4PredMode of block k (8.3.1.1): min of the modes of the left / top blocks, DC when missing."""
        bx, by = BLK_XY[k]

        def mode_at(nbx, nby, nb_mb):
            if 0 <= nbx < 4 and 0 <= nby < 4:
                return self.i4modes[cur][BLK_IDX[(nbx, nby)]]
            if nb_mb < 0:
                return None                              # not available -> dcPredModePredictedFlag
            if self.kind[nb_mb] != "i4":
                return 2
            return self.i4modes[nb_mb][BLK_IDX[(nbx % 4, nby % 4)]]
        ma = mode_at(bx - 1, by, a)
        mb_ = mode_at(bx, by - 1, b)
        if ma is None or mb_ is None:
            return 2
        return min(ma, mb_)

    def _mb_i4(self, w, cur, mb_type_base):
        a, b, c, d = self._neighbours(cur)
        w.ue(mb_type_base)                              # I_NxN
        self.kind[cur] = "i4"
        for k in range(16):
            bx, by = BLK_XY[k]
            av_a = bx > 0 or a >= 0
            av_b = by > 0 or b >= 0
            av_d = (bx > 0 and by > 0) or (bx > 0 and by == 0 and b >= 0) or (bx == 0 and by > 0 and a >= 0) or \
                   (bx == 0 and by == 0 and d >= 0)
            legal = [2]
            if av_b: legal += [0, 3, 7]
            if av_a: legal += [1, 8]
            if av_a and av_b and av_d: legal += [4, 5, 6]
            mode = self.rng.choice(legal)
            pred = self._i4_pred_mode(cur, k, a, b)
            if mode == pred:
                w.u(1, 1)
            else:
                w.u(1, 0)
                w.u(3, mode if mode < pred else mode - 1)
            self.i4modes[cur][k] = mode
        w.ue(self._chroma_mode(a, b, d))
        w.ue(3)                                          # coded_block_pattern: Intra 4x4, cbp 0 (Table 9-4 codeNum 3)

    def _mb_pcm(self, w, cur, mb_type_base):
        w.ue(mb_type_base + 25)
        w.align_zero()
        w.bytes_raw(bytes(self.rng.getrandbits(8) for _ in range(384)))
        self.is_pcm[cur] = True
        self.kind[cur] = "pcm"

    def _mvd(self, w):
        big = self.rng.random() < 0.05
        for _ in range(2):
            w.se(self.rng.randint(-40, 40) if big else self.rng.randint(-9, 9))

    def _mb_inter(self, w, cur, nref):
        t = self.rng.choice([0, 0, 1, 2, 3, 3])
        w.ue(t)
        self.kind[cur] = "inter"
        if t == 0:
            if nref > 1: w.te(self.rng.randrange(nref), nref - 1)
            self._mvd(w)
        elif t in (1, 2):
            if nref > 1:
                for _ in range(2): w.te(self.rng.randrange(nref), nref - 1)
            for _ in range(2): self._mvd(w)
        else:
            sub = [self.rng.randrange(4) for _ in range(4)]
            for s in sub: w.ue(s)
            if nref > 1:
                for _ in range(4): w.te(self.rng.randrange(nref), nref - 1)
            for s in sub:
                for _ in range((1, 2, 2, 4)[s]): self._mvd(w)
        w.ue(0)                                          # coded_block_pattern: Inter, cbp 0 (codeNum 0)

    # ---- slices / pictures ----
    def _slice_header(self, w, first_mb, is_idr, is_p, qp, idc, off_a, off_b):
        w.ue(first_mb)
        w.ue(0 if is_p else 2)                           # slice_type P / I
        w.ue(0)                                          # pic_parameter_set_id
        w.u(self.LOG2_MAX_FRAME_NUM, self.frame_num % (1 << self.LOG2_MAX_FRAME_NUM))
        if is_idr:
            w.ue(self.idr_id)
        w.u(self.LOG2_MAX_POC_LSB, self.poc % (1 << self.LOG2_MAX_POC_LSB))
        nref = 0
        if is_p:
            nref = min(self.num_refs, self.refs_available)
            override = nref != self.num_refs
            w.u(1, 1 if override else 0)                 # num_ref_idx_active_override_flag
            if override:
                w.ue(nref - 1)
            w.u(1, 0)                                    # ref_pic_list_modification_flag_l0
        if is_idr:
            w.u(1, 0); w.u(1, 0)                         # no_output_of_prior_pics_flag, long_term_reference_flag
        else:
            w.u(1, 0)                                    # adaptive_ref_pic_marking_mode_flag
        w.se(qp - 26)                                    # slice_qp_delta
        w.ue(idc)
        if idc != 1:
            w.se(off_a); w.se(off_b)
        return nref

    def picture(self, is_idr, qp=32, idc=0, off_a=0, off_b=0, slices=2, intra_share=0.15, skip_share=0.25):
        if is_idr:
            self.frame_num = 0
            self.poc = 0
        self._new_picture()
        n = self.W * self.H
        cuts = sorted(set([0] + ([self.rng.randrange(1, n)] if slices > 1 and n > 1 else [])))
        for si, first in enumerate(cuts):
            last = cuts[si + 1] if si + 1 < len(cuts) else n
            w = BitWriter()
            nref = self._slice_header(w, first, is_idr, not is_idr, qp, idc, off_a, off_b)
            self.qp_running = qp
            skip_run = 0
            for cur in range(first, last):
                self.slice_of[cur] = si
                if is_idr:
                    r = self.rng.random()
                    if r < 0.2: self._mb_pcm(w, cur, 0)
                    elif r < 0.6: self._mb_i16(w, cur, 0)
                    else: self._mb_i4(w, cur, 0)
                    continue
                r = self.rng.random()
                if r < skip_share:
                    skip_run += 1
                    self.kind[cur] = "skip"
                    continue
                w.ue(skip_run)
                skip_run = 0
                if r < skip_share + intra_share:
                    if self.rng.random() < 0.5: self._mb_i16(w, cur, 5)
                    else: self._mb_i4(w, cur, 5)
                else:
                    self._mb_inter(w, cur, nref)
            if not is_idr and skip_run:
                w.ue(skip_run)
            w.trailing()
            self.out += nal(3 if is_idr else 2, 5 if is_idr else 1, w.payload())
        if is_idr:
            self.idr_id += 1
            self.refs_available = 1
        else:
            self.refs_available = min(self.num_refs, self.refs_available + 1)
        self.frame_num += 1
        self.poc += 2

    def data(self):
        return bytes(self.out)


def make_stream(width_mbs=11, height_mbs=9, frames=8, seed=7, crop=None):
    """IDR + P pictures with every deblocking mode and a second IDR in the middle; `crop` = SPS frame cropping offsets."""
    s = Stream(width_mbs, height_mbs, seed, crop=crop)
    for i in range(frames):
        idr = i == 0 or i == frames // 2 + 1
        idc = (0, 0, 2, 1)[i % 4]
        s.picture(idr, qp=(30, 36, 42, 26)[i % 4], idc=idc, off_a=(0, 2, -2, 4)[i % 4], off_b=(0, -2, 2, 0)[i % 4],
                  slices=2 if i % 3 else 1)
    return s.data()


if __name__ == "__main__":
    import sys
    open(sys.argv[1], "wb").write(make_stream())
