// Arithmetic core of the motion-compensation kernel: quarter-pel luma (get_block_luma, reference
// decoder/inter_prediction.cc:158-340), eighth-pel chroma (get_block_chroma, :342-406) and weighted sample
// prediction + reconstruction (mc_prediction / bi_prediction :53-156, Transform::construction transform.cc:913-984)
// for the 4x4 luma / 2x2 chroma patch one lane owns.  Written on packed data:
//   * horizontal 6-tap   : IDP.4A (u8 x s8 dot product) on 4-byte-aligned row words, taps as shifted constants
//   * vertical 6-tap     : 16x2 SIMD-in-register (IADD3 / IMAD on biased, non-negative lanes; VIADDMNMX / VIMNMX clamps)
//   * centre sample j    : int32 vertical 6-tap over the unrounded horizontal sums
//   * clip + pack        : cvt.pack.sat.u8.s32 (I2IP), PRMT
// Every quarter-pel sample is G, b, h, j or the rounded average of two of them (spec 8.4.2.2.1 == the reference's
// branches), so the 16 fractional cases are stages predicated per lane instead of 16 code paths.
//
// The file compiles for the device (kernels.cu) and, with H264R_HOST_EMUL defined, for the host with the PTX
// primitives emulated: tests/mc_core_host_test.cc checks it exhaustively against a plain restatement on the CPU.
#ifndef H264R_MC_CORE_CUH_
#define H264R_MC_CORE_CUH_

#include <stdint.h>

namespace h264r {

#ifdef H264R_HOST_EMUL
#define MC_FN static inline
// ---- host emulation of the PTX primitives ----
MC_FN uint32_t mc_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    const uint64_t src = (uint64_t)a | ((uint64_t)b << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t s = (sel >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(src >> (8 * (s & 7))) & 0xFF;
        if (s & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
MC_FN uint32_t mc_shf_r(uint32_t lo, uint32_t hi, uint32_t sh)          // funnel shift right, sh in [0, 31]
{
    const uint64_t v = (uint64_t)lo | ((uint64_t)hi << 32);
    return (uint32_t)(v >> (sh & 31));
}
MC_FN int mc_dp4a_us(uint32_t a, uint32_t b, int c)                      // u8 x s8
{
    for (int i = 0; i < 4; ++i) c += (int)((a >> (8 * i)) & 0xFF) * (int)(int8_t)((b >> (8 * i)) & 0xFF);
    return c;
}
MC_FN uint32_t mc_dp4a_uu(uint32_t a, uint32_t b, uint32_t c)            // u8 x u8
{
    for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xFF) * ((b >> (8 * i)) & 0xFF);
    return c;
}
MC_FN int mc_dp2a_lo_su(uint32_t a, uint32_t b, int c)                  // s16 lanes of a x u8 bytes 0, 1 of b
{
    return c + (int)(int16_t)(a & 0xFFFF) * (int)(b & 0xFF) + (int)(int16_t)(a >> 16) * (int)((b >> 8) & 0xFF);
}
MC_FN uint32_t mc_pack_sat_u8(int a, int b, uint32_t c)                 // (c << 16) | sat_u8(a) << 8 | sat_u8(b)
{
    const uint32_t sa = (uint32_t)(a < 0 ? 0 : (a > 255 ? 255 : a)), sb = (uint32_t)(b < 0 ? 0 : (b > 255 ? 255 : b));
    return (c << 16) | (sa << 8) | sb;
}
MC_FN uint32_t mc_viaddmax_s16x2_relu(uint32_t a, uint32_t b, uint32_t c)   // per halfword: max(max(a + b, c), 0)
{
    uint32_t r = 0;
    for (int i = 0; i < 2; ++i) {
        int x = (int)(int16_t)(uint16_t)((int16_t)(a >> (16 * i)) + (int16_t)(b >> (16 * i)));
        const int cc = (int16_t)(c >> (16 * i));
        if (x < cc) x = cc;
        if (x < 0) x = 0;
        r |= (uint32_t)(uint16_t)x << (16 * i);
    }
    return r;
}
MC_FN uint32_t mc_vmins2(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 0; i < 2; ++i) {
        const int x = (int16_t)(a >> (16 * i)), y = (int16_t)(b >> (16 * i));
        r |= (uint32_t)(uint16_t)(x < y ? x : y) << (16 * i);
    }
    return r;
}
#else
#define MC_FN __device__ __forceinline__
MC_FN uint32_t mc_prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
MC_FN uint32_t mc_shf_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
MC_FN int mc_dp4a_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
MC_FN uint32_t mc_dp4a_uu(uint32_t a, uint32_t b, uint32_t c) { return __dp4a(a, b, c); }
MC_FN int mc_dp2a_lo_su(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
MC_FN uint32_t mc_pack_sat_u8(int a, int b, uint32_t c)
{
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
MC_FN uint32_t mc_viaddmax_s16x2_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2_relu(a, b, c); }
MC_FN uint32_t mc_vmins2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
#endif

// (a + b + 1) >> 1 on four packed bytes
MC_FN uint32_t mc_avg_u8x4(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) >> 1) & 0x7F7F7F7Fu); }
// four int32 -> four saturated bytes, v0 in byte 0
MC_FN uint32_t mc_pack4_sat(int v0, int v1, int v2, int v3) { return mc_pack_sat_u8(v1, v0, mc_pack_sat_u8(v3, v2, 0u)); }

// Horizontal 6-tap (1,-5,20,20,-5,1) at four consecutive positions.  a0|a1|a2 = twelve consecutive row bytes r0..r11
// (r0 in byte 0 of a0); out[x] = r[x] - 5 r[x+1] + 20 r[x+2] + 20 r[x+3] - 5 r[x+4] + r[x+5] + 16.
MC_FN void mc_htap4(uint32_t a0, uint32_t a1, uint32_t a2, int out[4])
{
    out[0] = mc_dp4a_us(a1, 0x000001FBu, mc_dp4a_us(a0, 0x1414FB01u, 16));
    out[1] = mc_dp4a_us(a1, 0x0001FB14u, mc_dp4a_us(a0, 0x14FB0100u, 16));
    out[2] = mc_dp4a_us(a1, 0x01FB1414u, mc_dp4a_us(a0, 0xFB010000u, 16));
    out[3] = mc_dp4a_us(a2, 0x00000001u, mc_dp4a_us(a1, 0xFB1414FBu, mc_dp4a_us(a0, 0x01000000u, 16)));
}

// Luma 4x2 patch.  win = the lane's window (4-byte aligned, row pitch kLumaPitchWords words); off = byte offset, inside
// the window, of the integer sample (0, 0) of the patch; the window holds the samples 2 to the left / above and 3 (+1
// for the patch width/height) to the right / below.  xf, yf = quarter-sample fractions.  Returns the two rows packed.
//
// One pass over the seven window rows (sample rows -2..4).  Per row, if the lane needs it:
//   horizontal stage: four unrounded 6-tap sums (+16) -> (rows 0..2 of the patch) clipped half sample b, and the
//                     running vertical 6-tap sums of the centre sample j (int32; the +16s add up to the +512);
//   raw stage       : the four samples at column offset dx -> integer samples G, and the running vertical 6-tap of
//                     the half sample h in 16x2 lanes (E = columns 0,2; O = columns 1,3).  The lanes start at
//                     2560 + 16 (= (80 << 5) + 16) and never go negative, so packed IMADs are exact per lane.
// The stages a lane needs are row masks (mc_luma_masks); the caller ORs them over the warp and passes the WARP masks:
// every lane then runs the same stages (lanes that do not need one compute values they never select), the branches
// are warp-uniform and cost no divergence bookkeeping.  Bit 0 of the horizontal mask is set only by lanes that need j.
template <int R>                                               // R = rows of the patch: 2 (4x2) or 4 (4x4)
MC_FN void mc_luma_masks_r(int xf, int yf, unsigned& hmask, unsigned& cmask)
{
    const bool hasB = xf != 0 && yf != 2, hasH = yf != 0 && xf != 2;
    const bool hasJ = (xf == 2 && yf != 0) || (yf == 2 && xf != 0);
    const bool hasG = (xf == 0 && yf != 2) || (yf == 0 && xf != 2);
    const int dy = yf == 3;
    constexpr unsigned all = (1u << (R + 5)) - 1, rows = (1u << R) - 1;
    hmask = hasJ ? all : (hasB ? rows << (2 + dy) : 0u);                 // rows that get the horizontal 6-tap
    cmask = hasH ? all : ((hasG && yf == 0) ? rows << 2 : 0u);           // rows whose raw samples are needed
}
MC_FN void mc_luma_masks(int xf, int yf, unsigned& hmask, unsigned& cmask) { mc_luma_masks_r<2>(xf, yf, hmask, cmask); }

constexpr int kLumaPitchWords = 4;
// 4 x R luma patch: one pass over the R + 5 window rows (sample rows -2 .. R + 2); out[r] = row r, four packed samples.
template <int R>
MC_FN void mc_luma_patch(const uint32_t* win, int off, int xf, int yf, unsigned hmask, unsigned cmask, uint32_t (&out)[R],
                         const int pitch_words = kLumaPitchWords)
{
    const bool hasB = xf != 0 && yf != 2;                       // clipped horizontal half sample b (row + dy)
    const bool hasH = yf != 0 && xf != 2;                       // clipped vertical half sample h (column + dx)
    const bool hasJ = (xf == 2 && yf != 0) || (yf == 2 && xf != 0);
    const bool hasG = (xf == 0 && yf != 2) || (yf == 0 && xf != 2);
    const int dx = xf == 3, dy = yf == 3;
    const bool any_j = hmask & 1u;

    const int o2 = off - 2;
    const uint32_t* wa = win + (o2 >> 2);
    const uint32_t sha = (uint32_t)(o2 & 3) * 8;
    const int oc = off + dx;
    const uint32_t* wc = win + (oc >> 2);
    const uint32_t shc = (uint32_t)(oc & 3) * 8;

    int jacc[R][4];
    uint32_t tE[R], tO[R];
    uint32_t Brow[R + 1];                                       // window rows 2 .. R + 2 (the integer samples G of those rows
                                                                // are re-read from the window at the end: five registers less)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        tE[r] = tO[r] = 0x0A100A10u;
#pragma unroll
        for (int x = 0; x < 4; ++x) jacc[r][x] = 0;
    }
#pragma unroll
    for (int r = 0; r <= R; ++r) Brow[r] = 0;
    constexpr int tap[6] = { 1, -5, 20, 20, -5, 1 };
#pragma unroll
    for (int k = 0; k < R + 5; ++k) {
        if ((hmask >> k) & 1) {
            const uint32_t w0 = wa[k * pitch_words], w1 = wa[k * pitch_words + 1], w2 = wa[k * pitch_words + 2];
            int b1[4];
            mc_htap4(mc_shf_r(w0, w1, sha), mc_shf_r(w1, w2, sha), w2 >> sha, b1);
            if (k >= 2 && k <= R + 2) Brow[k - 2] = mc_pack4_sat(b1[0] >> 5, b1[1] >> 5, b1[2] >> 5, b1[3] >> 5);
            if (any_j) {
#pragma unroll
                for (int r = 0; r < R; ++r)                     // output row r takes window rows r .. r + 5
                    if (k >= r && k <= r + 5) {
#pragma unroll
                        for (int x = 0; x < 4; ++x) jacc[r][x] += tap[k - r] * b1[x];
                    }
            }
        }
        if ((cmask >> k) & 1) {
            const uint32_t c = mc_shf_r(wc[k * pitch_words], wc[k * pitch_words + 1], shc);
            const uint32_t e = mc_prmt(c, 0u, 0x4240u), o = mc_prmt(c, 0u, 0x4341u);
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (k >= r && k <= r + 5) { tE[r] += (uint32_t)tap[k - r] * e; tO[r] += (uint32_t)tap[k - r] * o; }
        }
    }

    const bool down = xf == 0 && dy;                             // integer samples at (0, dy)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        uint32_t P, Q;                                           // the (up to) two samples averaged per position
        const uint32_t Bv = dy ? Brow[r + 1] : Brow[r];
        // first: G, else b, else h, else j
        if (hasG) { const uint32_t* g = wc + (r + 2 + (down ? 1 : 0)) * pitch_words; P = mc_shf_r(g[0], g[1], shc); }
        else if (hasB) P = Bv;
        else P = 0;
        uint32_t Hv = 0, Jv = 0;
        if (hasH) {
            const uint32_t re = mc_vmins2(mc_viaddmax_s16x2_relu((tE[r] >> 5) & 0x07FF07FFu, 0xFFB0FFB0u, 0u), 0x00FF00FFu);
            const uint32_t ro = mc_vmins2(mc_viaddmax_s16x2_relu((tO[r] >> 5) & 0x07FF07FFu, 0xFFB0FFB0u, 0u), 0x00FF00FFu);
            Hv = mc_prmt(re, ro, 0x6240u);
        }
        if (hasJ) Jv = mc_pack4_sat(jacc[r][0] >> 10, jacc[r][1] >> 10, jacc[r][2] >> 10, jacc[r][3] >> 10);
        if (!hasG && !hasB) P = hasH ? Hv : Jv;
        // second: j, else h, else b, else G (== first when only one kind is present)
        if (hasJ) Q = Jv;
        else if (hasH) Q = Hv;
        else if (hasB) Q = Bv;
        else Q = P;
        out[r] = mc_avg_u8x4(P, Q);
    }
}
MC_FN void mc_luma_patch_4x2(const uint32_t* win, int off, int xf, int yf, unsigned hmask, unsigned cmask, uint32_t& out0, uint32_t& out1)
{
    uint32_t o[2];
    mc_luma_patch<2>(win, off, xf, yf, hmask, cmask, o);
    out0 = o[0]; out1 = o[1];
}

// Chroma 2x2 patch of one plane.  win: row pitch `pitch` words (2 = 8 bytes; the TMA build stages uniform quadrants with 4);
// off = byte offset of sample (0, 0), at most 5; xf, yf = eighth-sample fractions.  Returns the four samples packed (row 0
// in bytes 0-1, row 1 in bytes 2-3).
MC_FN uint32_t mc_chroma_patch_2x2(const uint32_t* win, int off, int xf, int yf, int pitch = 2)
{
    const uint32_t* w = win + (off >> 2);
    const uint32_t sh = (uint32_t)(off & 3) * 8;
    uint32_t R[3];
#pragma unroll
    for (int y = 0; y < 3; ++y) R[y] = mc_shf_r(w[pitch * y], w[pitch * y + 1], sh);
    const uint32_t wgt = (uint32_t)((8 - xf) * (8 - yf)) | (uint32_t)(xf * (8 - yf)) << 8 | (uint32_t)((8 - xf) * yf) << 16 | (uint32_t)(xf * yf) << 24;
    uint32_t v[4];
#pragma unroll
    for (int y = 0; y < 2; ++y) {
        v[2 * y]     = mc_dp4a_uu(mc_prmt(R[y], R[y + 1], 0x5410u), wgt, 32u) >> 6;
        v[2 * y + 1] = mc_dp4a_uu(mc_prmt(R[y], R[y + 1], 0x6521u), wgt, 32u) >> 6;
    }
    return v[0] | v[1] << 8 | v[2] << 16 | v[3] << 24;
}

// Weighted sample prediction of four packed samples (mc_prediction / bi_prediction, inter_prediction.cc:53-156):
//   mode 0: p0                         mode 1: clip1(((w0 p0 + 2^(d-1)) >> d) + o)      [d == 0: no rounding shift]
//   mode 2: (p0 + p1 + 1) >> 1         mode 3: clip1(((w0 p0 + w1 p1 + 2^d) >> (d + 1)) + o)
// `o` is the final offset (mode 3: (o0 + o1 + 1) >> 1).  Returns the four prediction samples packed.
MC_FN uint32_t mc_weight4(int mode, uint32_t p0, uint32_t p1, int w0, int w1, int d, int o)
{
    if (mode & 1) {
        // one dot product per sample: (w0, w1) as 16-bit lanes (implicit weights reach 128) x (p0_i, p1_i) as bytes
        int v[4];
        const int sh = mode == 1 ? d : d + 1;
        const int rnd = sh > 0 ? 1 << (sh - 1) : 0;
        const uint32_t w = (uint32_t)(uint16_t)w0 | (mode == 3 ? (uint32_t)(uint16_t)w1 << 16 : 0u);
        v[0] = (mc_dp2a_lo_su(w, mc_prmt(p0, p1, 0x0040u), rnd) >> sh) + o;
        v[1] = (mc_dp2a_lo_su(w, mc_prmt(p0, p1, 0x0051u), rnd) >> sh) + o;
        v[2] = (mc_dp2a_lo_su(w, mc_prmt(p0, p1, 0x0062u), rnd) >> sh) + o;
        v[3] = (mc_dp2a_lo_su(w, mc_prmt(p0, p1, 0x0073u), rnd) >> sh) + o;
        return mc_pack4_sat(v[0], v[1], v[2], v[3]);              // clip1
    }
    return mode == 2 ? mc_avg_u8x4(p0, p1) : p0;
}
// Reconstruction clip1(pred + residual) (transform.cc:913-984) of four packed samples; res01 / res23 = residuals as
// int16 pairs, each within [-32768 + 255, 32767 - 255].
MC_FN uint32_t mc_recon4(uint32_t pred, uint32_t res01, uint32_t res23)
{
    uint32_t lo = mc_prmt(pred, 0u, 0x4140u), hi = mc_prmt(pred, 0u, 0x4342u);
    lo = mc_vmins2(mc_viaddmax_s16x2_relu(lo, res01, 0u), 0x00FF00FFu);
    hi = mc_vmins2(mc_viaddmax_s16x2_relu(hi, res23, 0u), 0x00FF00FFu);
    return mc_prmt(lo, hi, 0x6420u);
}
MC_FN uint32_t mc_weight_recon4(int mode, uint32_t p0, uint32_t p1, int w0, int w1, int d, int o, uint32_t res01, uint32_t res23)
{
    return mc_recon4(mc_weight4(mode, p0, p1, w0, w1, d, o), res01, res23);
}

} // namespace h264r
#endif
