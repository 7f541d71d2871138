// deblock4_kernel: the in-loop deblocking filter (deblock.cc:327-552) with TWO sample lines per lane in packed fp16x2
// arithmetic: eight lanes filter one picture, a warp filters the same macroblock row of FOUR pictures.  Same row
// wavefront, same mailbox protocol and same descriptors as deblock_kernel (kernel_deblock.cuh), half the warps.
// A build option (-DH264R_DEBLOCK_PACKED=1, kernels.cu): bit-exact on every GPU parity test, 22-30 % fewer warp instructions
// than deblock_kernel, but 8.7 ms per step against 7.9 -- the wavefront is bound by the latency of its dependent chain,
// which two lines per lane make longer (profiles/r2_deblock4_variant.json).
//
// Why fp16x2 is exact here.  A sample x is held as the half 1024 + x (bit pattern 0x6400 | x: a byte permute builds it
// and takes it apart, no conversion instruction).  The filters are written on DIFFERENCES of samples (the bias
// cancels) relative to the sample they update: every intermediate is an integer of magnitude < 2048 or such an
// integer times 1/2, 1/4, 1/8 plus a rounding constant of 1/4, 1/8, 1/16, all exactly representable in the ranges that
// can reach an output (proof per formula below).  floor(n / 2^k) is taken as rint(n / 2^k - (2^k - 1) / 2^(k+1)), rint(y)
// as (y + 1536) - 1536 (ulp 1 in [1024, 2048), round-to-nearest-even never meets a tie).
#ifndef H264R_KERNEL_DEBLOCK4_CUH_
#define H264R_KERNEL_DEBLOCK4_CUH_

#include <cuda_fp16.h>
#include "kernel_deblock.cuh"

namespace h264r {

#ifndef H264R_DEBLOCK4_CTAS
#define H264R_DEBLOCK4_CTAS (16 / H264R_WARPS_PER_CTA)
#endif

typedef __half2 h2;
__device__ __forceinline__ unsigned h2b(h2 v) { return *reinterpret_cast<unsigned*>(&v); }
__device__ __forceinline__ h2 b2h(unsigned u) { return *reinterpret_cast<h2*>(&u); }
__device__ __forceinline__ h2 h2c(float v) { return __float2half2_rn(v); }                  // compile-time constant
__device__ __forceinline__ h2 h2sel(unsigned m, h2 a, h2 b) { return b2h((h2b(a) & m) | (h2b(b) & ~m)); }
__device__ __forceinline__ h2 h2round(h2 y) { return __hadd2(__hadd2(y, h2c(1536.f)), h2c(-1536.f)); }
__device__ __forceinline__ h2 h2clamp(h2 v, h2 lim) { return __hmax2(__hmin2(v, lim), __hneg2(lim)); }   // to [-lim, lim]
__device__ __forceinline__ unsigned h2lt(h2 a, h2 b) { return __hlt2_mask(a, b); }          // 0xFFFF per half where a < b
constexpr unsigned kH2Magic = 0x64646464u;                     // byte 0x64 = the high byte of the halves 1024 .. 1279
// small unsigned integer (< 256, in the low byte of `v`) in both halves
__device__ __forceinline__ h2 h2_bcast_byte(unsigned v) { return __hadd2(b2h(__byte_perm(v, kH2Magic, 0x4040)), h2c(-1024.f)); }
// low bytes of a (low half) and b (high half)
__device__ __forceinline__ h2 h2_pair_bytes(unsigned a, unsigned b)
{
    return __hadd2(b2h((__byte_perm(a, b, 0x0400) & 0x00FF00FFu) | 0x64006400u), h2c(-1024.f));
}

// One side of the bS = 4 luma filter (deblock.cc:327-371), on biased samples S0..S3 of the side being updated.
// d1 = s1 - s0, d2 = s2 - s0, e = o0 - s0 (the other side's first sample), o1 = o1 - o0; strong = mask of the lines
// that take the three-tap branch.  With dS = differences:
//   s0' = s0 + floor((d2 + 2 d1 + 3 e + o1 + 4) / 8)         [(s2 + 2 s1 + 2 s0 + 2 o0 + o1 + 4) >> 3]
//   s1' = s1 + floor((d2 - 3 d1 + e + 2) / 4)                [(s2 + s1 + s0 + o0 + 2) >> 2]
//   s2' = s2 + floor((2 (s3 - s2) + d1 - 3 d2 + e + 4) / 8)  [(2 s3 + 3 s2 + s1 + s0 + o0 + 4) >> 3]
//   weak: s0' = s0 + floor((2 d1 + e + o1 + 2) / 4)          [(2 s1 + s0 + o1 + 2) >> 2]
// Ranges on lines that are filtered at all: |d1|, |o1| < beta <= 18, |e| < alpha <= 255; on strong lines also |d2| < 18
// and |e| < 66, |s3 - s2| <= 255: every numerator stays below 1100 in magnitude, every quotient + constant below 128
// where halves resolve 1/16.
__device__ __forceinline__ void strong_side(h2& S0, h2& S1, h2& S2, h2 S3, h2 d1, h2 d2, h2 e, h2 o1, unsigned strong)
{
    const h2 eo = __hadd2(e, o1);
    const h2 weak = h2round(__hfma2(__hfma2(d1, h2c(2.f), eo), h2c(0.25f), h2c(0.125f)));
    const h2 n0 = __hfma2(d1, h2c(2.f), __hfma2(e, h2c(2.f), __hadd2(d2, eo)));                // d2 + 2 d1 + 3 e + o1
    const h2 a0 = h2round(__hfma2(n0, h2c(0.125f), h2c(0.0625f)));
    const h2 n1 = __hfma2(d1, h2c(-3.f), __hadd2(d2, e));
    const h2 a1 = h2round(__hfma2(n1, h2c(0.25f), h2c(0.125f)));
    const h2 n2 = __hfma2(__hsub2(S3, S2), h2c(2.f), __hfma2(d2, h2c(-3.f), __hadd2(d1, e)));
    const h2 a2 = h2round(__hfma2(n2, h2c(0.125f), h2c(0.0625f)));
    S2 = h2sel(strong, __hadd2(S2, a2), S2);
    S1 = h2sel(strong, __hadd2(S1, a1), S1);
    S0 = __hadd2(S0, h2sel(strong, a0, weak));
}

// filter_strong / filter_normal (deblock.cc:327-415) across one edge for the two lines of a lane; P[0] = p0 ... P[3] = p3
// as halves 1024 + sample.  alpha, beta, tc0: per half (the two lines of a luma lane share them; a chroma lane holds Cb
// in the low and Cr in the high half).  bS is the same for both lines.
template <bool kChroma>
__device__ __forceinline__ void filter_edge2(int bS, h2 alpha, h2 beta, h2 tc0, h2 alpha4, h2 (&P)[4], h2 (&Q)[4])
{
    const h2 e = __hsub2(Q[0], P[0]);                       // q0 - p0
    const h2 dp1 = __hsub2(P[1], P[0]), dq1 = __hsub2(Q[1], Q[0]);
    const unsigned m = h2lt(__habs2(e), alpha) & h2lt(__habs2(dp1), beta) & h2lt(__habs2(dq1), beta);
    if (m == 0) return;
    if (kChroma) {
        h2 np, nq;
        if (bS == 4) {                                      // p0 = (2 p1 + p0 + q1 + 2) >> 2 = p0 + floor((2 dp1 + e + dq1 + 2) / 4)
            np = __hadd2(P[0], h2round(__hfma2(__hfma2(dp1, h2c(2.f), __hadd2(e, dq1)), h2c(0.25f), h2c(0.125f))));
            nq = __hadd2(Q[0], h2round(__hfma2(__hfma2(dq1, h2c(2.f), __hsub2(dp1, e)), h2c(0.25f), h2c(0.125f))));
        } else {
            const h2 tc = __hadd2(tc0, h2c(1.f));
            // delta = clip(((q0 - p0) * 4 + (p1 - q1) + 4) >> 3); p1 - q1 = dp1 - dq1 - e.  A quotient beyond +-128 loses
            // its fraction but is clipped to tc <= 26 anyway.
            const h2 t = __hfma2(e, h2c(3.f), __hsub2(dp1, dq1));
            const h2 d = h2clamp(h2round(__hfma2(t, h2c(0.125f), h2c(0.0625f))), tc);
            np = __hmin2(__hmax2(__hadd2(P[0], d), h2c(1024.f)), h2c(1279.f));
            nq = __hmin2(__hmax2(__hsub2(Q[0], d), h2c(1024.f)), h2c(1279.f));
        }
        P[0] = h2sel(m, np, P[0]); Q[0] = h2sel(m, nq, Q[0]);
        return;
    }
    const h2 dp2 = __hsub2(P[2], P[0]), dq2 = __hsub2(Q[2], Q[0]);
    const unsigned ap = h2lt(__habs2(dp2), beta), aq = h2lt(__habs2(dq2), beta);
    if (bS == 4) {
        const unsigned small = h2lt(__habs2(e), alpha4);    // |p0 - q0| < (alpha >> 2) + 2
        h2 p0 = P[0], p1 = P[1], p2 = P[2], q0 = Q[0], q1 = Q[1], q2 = Q[2];
        strong_side(p0, p1, p2, P[3], dp1, dp2, e, dq1, ap & small);
        strong_side(q0, q1, q2, Q[3], dq1, dq2, __hneg2(e), dp1, aq & small);
        P[0] = h2sel(m, p0, P[0]); P[1] = h2sel(m, p1, P[1]); P[2] = h2sel(m, p2, P[2]);
        Q[0] = h2sel(m, q0, Q[0]); Q[1] = h2sel(m, q1, Q[1]); Q[2] = h2sel(m, q2, Q[2]);
        return;
    }
    const h2 one = h2c(1.f);
    const h2 tc = __hadd2(__hadd2(tc0, b2h(ap & h2b(one))), b2h(aq & h2b(one)));
    const h2 t = __hfma2(e, h2c(3.f), __hsub2(dp1, dq1));
    const h2 d = h2clamp(h2round(__hfma2(t, h2c(0.125f), h2c(0.0625f))), tc);
    const h2 np0 = __hmin2(__hmax2(__hadd2(P[0], d), h2c(1024.f)), h2c(1279.f));
    const h2 nq0 = __hmin2(__hmax2(__hsub2(Q[0], d), h2c(1024.f)), h2c(1279.f));
    // p1 += clip((p2 + ((p0 + q0 + 1) >> 1) - 2 p1) >> 1): with avg - p0 = floor((e + 1) / 2) = rint(e / 2 + 1/4),
    // p2 + avg - 2 p1 = dp2 + (avg - p0) - 2 dp1; q side: q2 + avg - 2 q1 = dq2 + (avg - p0 - e) - 2 dq1
    const h2 av = h2round(__hfma2(e, h2c(0.5f), h2c(0.25f)));
    const h2 up = __hfma2(dp1, h2c(-2.f), __hadd2(dp2, av));
    const h2 uq = __hfma2(dq1, h2c(-2.f), __hadd2(dq2, __hsub2(av, e)));
    const h2 np1 = __hadd2(P[1], h2clamp(h2round(__hfma2(up, h2c(0.5f), h2c(-0.25f))), tc0));
    const h2 nq1 = __hadd2(Q[1], h2clamp(h2round(__hfma2(uq, h2c(0.5f), h2c(-0.25f))), tc0));
    P[1] = h2sel(m & ap, np1, P[1]); Q[1] = h2sel(m & aq, nq1, Q[1]);
    P[0] = h2sel(m, np0, P[0]); Q[0] = h2sel(m, nq0, Q[0]);
}

// thresholds of one luma edge out of the descriptor word (alpha | beta << 8 | tc0[bS 1..3] << 13, 18, 23)
__device__ __forceinline__ void edge_params1(unsigned par, int bS, h2& alpha, h2& beta, h2& tc0, h2& alpha4)
{
    alpha = h2_bcast_byte(par);
    beta = h2_bcast_byte((par >> 8) & 31);
    tc0 = h2_bcast_byte((par >> (8 + 5 * (bS & 3))) & 31);  // (bS 4 does not use it)
    alpha4 = h2_bcast_byte(((par & 0xFF) >> 2) + 2);
}
// ... of one chroma edge: Cb in the low half, Cr in the high half
__device__ __forceinline__ void edge_params2(unsigned par_cb, unsigned par_cr, int bS, h2& alpha, h2& beta, h2& tc0)
{
    const int sh = 8 + 5 * (bS & 3);
    alpha = h2_pair_bytes(par_cb, par_cr);
    beta = h2_pair_bytes((par_cb >> 8) & 31, (par_cr >> 8) & 31);
    tc0 = h2_pair_bytes((par_cb >> sh) & 31, (par_cr >> sh) & 31);
}

// four sample pairs out of two words: byte i of `a` in the low half, byte i of `b` in the high half
__device__ __forceinline__ void unpack4x2(unsigned a, unsigned b, h2* v)
{
    const unsigned lo = __byte_perm(a, b, 0x5140), hi = __byte_perm(a, b, 0x7362);   // a0 b0 a1 b1 | a2 b2 a3 b3
    v[0] = b2h(__byte_perm(lo, kH2Magic, 0x4140)); v[1] = b2h(__byte_perm(lo, kH2Magic, 0x4342));
    v[2] = b2h(__byte_perm(hi, kH2Magic, 0x4140)); v[3] = b2h(__byte_perm(hi, kH2Magic, 0x4342));
}
// ... and back: the low halves of v[0..3] into `a`, the high halves into `b`
__device__ __forceinline__ void pack4x2(const h2* v, unsigned& a, unsigned& b)
{
    const unsigned x = __byte_perm(h2b(v[0]), h2b(v[1]), 0x6240), y = __byte_perm(h2b(v[2]), h2b(v[3]), 0x6240);   // a0 a1 b0 b1 | a2 a3 b2 b3
    a = __byte_perm(x, y, 0x5410); b = __byte_perm(x, y, 0x7632);
}
// two pairs, interleaved: a0 b0 a1 b1
__device__ __forceinline__ unsigned pack2x2_interleaved(h2 v0, h2 v1) { return __byte_perm(h2b(v0), h2b(v1), 0x6420); }
// the pair held in the two low bytes of `w` (byte 0 -> low half)
__device__ __forceinline__ h2 unpack_u16(unsigned w) { return b2h(__byte_perm(w, kH2Magic, 0x4140)); }
__device__ __forceinline__ unsigned short pack_u16(h2 v) { return (unsigned short)__byte_perm(h2b(v), 0u, 0x4420); }

// Shared memory of one picture (a quarter warp): luma tile 16 rows x 16 B, rows 8..15 shifted by 16 B so that the 128-bit row
// stores of the eight lanes (rows 2j, 2j+1) hit eight distinct bank groups; chroma tile 8 rows x 16 B with Cb and Cr
// interleaved (one 16-bit load = the pair of a chroma lane); the mailbox words of the MB above (4 luma rows x 16 B, then
// [plane][row 6, 7][8 B]).  528 = 16 (mod 128): the 16-bit column reads of the four quarters fall on disjoint banks.
struct __align__(16) Deblock4Tile {
    uint8_t y[16 * 16 + 16];
    uint8_t c[8 * 16];
    uint8_t top_y[4 * 16];
    uint8_t top_c[2 * 2 * 8];
    uint8_t pad[528 - (16 * 16 + 16) - 8 * 16 - 4 * 16 - 32];
};
static_assert(sizeof(Deblock4Tile) == 528, "quarter tiles are skewed by 16 bytes");
__device__ __forceinline__ int tile_row(int r) { return r * 16 + (r & 8) * 2; }

__global__ void __launch_bounds__(kWarpsPerCta * 32, H264R_DEBLOCK4_CTAS)
deblock4_kernel(const DevPicture* __restrict__ pics, int num_pics, int* tickets, FrameGeom g, uint32_t epoch)
{
    __shared__ __align__(16) Deblock4Tile smem_all[kWarpsPerCta][4];
    __shared__ int s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(&tickets[1], 1);
    __syncthreads();
    const int W = g.width_mbs, H = g.height_mbs;
    const int groups = (H + kWarpsPerCta - 1) / kWarpsPerCta;
    const int nquads = (num_pics + 3) >> 2;
    const int rg = s_ticket / nquads, quad = s_ticket - rg * nquads;           // row-group-major, see recon_intra_kernel
    if (rg >= groups) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mby = rg * kWarpsPerCta + warp;
    if (mby >= H) return;
    const int qd = lane >> 3, j = lane & 7, qbase = lane & 24;
    const int pic_i = quad * 4 + qd;
    const DevPicture& pic = pics[min(pic_i, num_pics - 1)];
    const bool enabled = pic_i < num_pics && pic.run_deblock;                  // this quarter has a picture to filter
    if (!__any_sync(0xFFFFFFFFu, enabled)) return;
    Deblock4Tile& sm = smem_all[warp][qd];
    uint8_t* const TYa = sm.y + tile_row(2 * j);           // this lane's luma rows 2j, 2j + 1 (vertical pass)
    uint8_t* const TYb = sm.y + tile_row(2 * j + 1);
    uint8_t* const TCr = sm.c + j * 16;                    // chroma row j, Cb / Cr interleaved
    uint8_t* const dY = pic.dst;
    uint8_t* const dCb = pic.dst + g.off_cb;
    uint8_t* const dCr = pic.dst + g.off_cr;
    const int pitch_y = g.pitch_y, pitch_c = g.pitch_c;
    const uint4* const desc = reinterpret_cast<const uint4*>(pic.desc + (size_t)mby * W);
    const int py = mby * 16, cy = mby * 8;
    const int gsh = (j >> 1) * 4;                          // nibble of the 4-sample group of this lane's lines
    const bool has_above = mby > 0, has_below = mby + 1 < H;                   // warp-uniform
    const bool own_ya = enabled && (2 * j <= 12 || !has_below);                // frame rows this warp stores itself
    const bool own_yb = enabled && (2 * j + 1 <= 12 || !has_below);
    const bool own_c = enabled && (j <= 6 || !has_below);
    uint64_t* const box_out = pic.mbox + (size_t)mby * W * kMboxWords;
    const uint64_t* const box_in = pic.mbox + (size_t)(has_above ? mby - 1 : 0) * W * kMboxWords;
    // Mailbox words this lane posts: luma words 2j, 2j + 1 = row 12 + (j >> 1), samples 8 (j & 1) .. + 7; chroma word j =
    // plane j >> 2, row 6 + ((j >> 1) & 1), samples 4 (j & 1) .. + 3.  The last word of a row is final only after the next
    // MB's left edge: it comes from the lane that owns that row in the vertical pass.
    const uint8_t* const boxsrc_y = sm.y + tile_row(12 + (j >> 1)) + 8 * (j & 1);
    const uint8_t* const boxsrc_c = sm.c + (6 + ((j >> 1) & 1)) * 16 + 8 * (j & 1);
    const int boxlane_y = qbase + 6 + (j >> 2), boxlane_c = qbase + 6 + ((j >> 1) & 1);

    uint4 n_bs = make_uint4(0, 0, 0, 0), n_pa = n_bs, n_pb = n_bs, n_ya = n_bs, n_yb = n_bs;
    uint2 n_cb = make_uint2(0, 0), n_cr = n_cb; uint32_t n_pz = 0;
    if (enabled) {
        n_bs = __ldg(desc); n_pa = __ldg(desc + 1); n_pb = __ldg(desc + 2); n_pz = __ldg(reinterpret_cast<const unsigned int*>(desc) + 12);
        n_ya = __ldcg(reinterpret_cast<const uint4*>(dY + (uint32_t)((py + 2 * j) * pitch_y)));
        n_yb = __ldcg(reinterpret_cast<const uint4*>(dY + (uint32_t)((py + 2 * j + 1) * pitch_y)));
        n_cb = __ldcg(reinterpret_cast<const uint2*>(dCb + (uint32_t)((cy + j) * pitch_c)));
        n_cr = __ldcg(reinterpret_cast<const uint2*>(dCr + (uint32_t)((cy + j) * pitch_c)));
    }
    uint32_t boxY0 = 0, boxY1 = 0, boxC = 0;               // this lane's mailbox words of the previous MB (after its horizontal pass)

    for (int mbx = 0; mbx < W; ++mbx) {
        const uint4 bs = n_bs, ya = n_ya, yb = n_yb; const uint2 ocb = n_cb, ocr = n_cr;
        const uint32_t parY0 = n_pa.x, parY1 = n_pa.y, parY2 = n_pa.z;                                     // left MB edge, internal, top MB edge
        const uint32_t parB0 = n_pa.w, parB1 = n_pb.x, parB2 = n_pb.y, parR0 = n_pb.z, parR1 = n_pb.w, parR2 = n_pz;
        const int px = mbx * 16, cx = mbx * 8;

        // mailbox of the MB above: issued now, looked at after the vertical pass
        uint64_t t0 = 0, t1 = 0, t2 = 0;
        if (has_above && enabled) {
            t0 = ld_mbox(box_in + mbx * kMboxWords + 2 * j);
            t1 = ld_mbox(box_in + mbx * kMboxWords + 2 * j + 1);
            t2 = ld_mbox(box_in + mbx * kMboxWords + 16 + j);
        }
        // previous MB's last four samples of this lane's lines (final but for this MB's left edge)
        const uint32_t carryA = *reinterpret_cast<const uint32_t*>(TYa + 12), carryB = *reinterpret_cast<const uint32_t*>(TYb + 12);
        const uint2 carryC = *reinterpret_cast<const uint2*>(TCr + 8);

        // prefetch the next MB: independent of every other MB of this kernel
        if (mbx + 1 < W && enabled) {
            n_bs = __ldg(desc + (mbx + 1) * 4); n_pa = __ldg(desc + (mbx + 1) * 4 + 1); n_pb = __ldg(desc + (mbx + 1) * 4 + 2);
            n_pz = __ldg(reinterpret_cast<const unsigned int*>(desc + (mbx + 1) * 4) + 12);
            n_ya = __ldcg(reinterpret_cast<const uint4*>(dY + (uint32_t)((py + 2 * j) * pitch_y + px + 16)));
            n_yb = __ldcg(reinterpret_cast<const uint4*>(dY + (uint32_t)((py + 2 * j + 1) * pitch_y + px + 16)));
            n_cb = __ldcg(reinterpret_cast<const uint2*>(dCb + (uint32_t)((cy + j) * pitch_c + cx + 8)));
            n_cr = __ldcg(reinterpret_cast<const uint2*>(dCr + (uint32_t)((cy + j) * pitch_c + cx + 8)));
        }

        // ---- vertical edges, in registers ----
        uint32_t leftA, leftB, leftCb, leftCr;             // the previous MB's last four samples after this MB's left edge
        {
            h2 v[20];
            unpack4x2(carryA, carryB, v); unpack4x2(ya.x, yb.x, v + 4); unpack4x2(ya.y, yb.y, v + 8);
            unpack4x2(ya.z, yb.z, v + 12); unpack4x2(ya.w, yb.w, v + 16);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int s = enabled ? ((e < 2 ? bs.x : bs.y) >> ((e & 1) * 16 + gsh)) & 7 : 0;
                if (s) {
                    h2 alpha, beta, tc0, alpha4;
                    edge_params1(e ? parY1 : parY0, s, alpha, beta, tc0, alpha4);
                    h2 p[4] = { v[4 * e + 3], v[4 * e + 2], v[4 * e + 1], v[4 * e] }, q[4] = { v[4 * e + 4], v[4 * e + 5], v[4 * e + 6], v[4 * e + 7] };
                    filter_edge2<false>(s, alpha, beta, tc0, alpha4, p, q);
                    v[4 * e + 3] = p[0]; v[4 * e + 2] = p[1]; v[4 * e + 1] = p[2];
                    v[4 * e + 4] = q[0]; v[4 * e + 5] = q[1]; v[4 * e + 6] = q[2];
                }
            }
            uint4 ra, rb;
            pack4x2(v + 4, ra.x, rb.x); pack4x2(v + 8, ra.y, rb.y); pack4x2(v + 12, ra.z, rb.z); pack4x2(v + 16, ra.w, rb.w);
            *reinterpret_cast<uint4*>(TYa) = ra; *reinterpret_cast<uint4*>(TYb) = rb;
            pack4x2(v, leftA, leftB);
            if ((bs.x & 0xFFFF) && mbx > 0) {                 // columns 13..15 of the left MB
                if (own_ya) *reinterpret_cast<uint32_t*>(dY + (uint32_t)((py + 2 * j) * pitch_y + px - 4)) = leftA;
                if (own_yb) *reinterpret_cast<uint32_t*>(dY + (uint32_t)((py + 2 * j + 1) * pitch_y + px - 4)) = leftB;
            }
        }
        {
            h2 v[12];                                      // (Cb, Cr) pairs: the left MB's columns 4..7, this MB's columns 0..7
            v[0] = unpack_u16(carryC.x); v[1] = unpack_u16(carryC.x >> 16); v[2] = unpack_u16(carryC.y); v[3] = unpack_u16(carryC.y >> 16);
            unpack4x2(ocb.x, ocr.x, v + 4); unpack4x2(ocb.y, ocr.y, v + 8);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int s = enabled ? ((e ? bs.y : bs.x) >> gsh) & 7 : 0;           // chroma edge e <- luma edge 2e
                if (s) {
                    h2 alpha, beta, tc0;
                    edge_params2(e ? parB1 : parB0, e ? parR1 : parR0, s, alpha, beta, tc0);
                    h2 p[4] = { v[4 * e + 3], v[4 * e + 2], v[4 * e + 2], v[4 * e + 2] }, q[4] = { v[4 * e + 4], v[4 * e + 5], v[4 * e + 5], v[4 * e + 5] };
                    filter_edge2<true>(s, alpha, beta, tc0, tc0, p, q);
                    v[4 * e + 3] = p[0]; v[4 * e + 4] = q[0];
                }
            }
            *reinterpret_cast<uint4*>(TCr) = make_uint4(pack2x2_interleaved(v[4], v[5]), pack2x2_interleaved(v[6], v[7]),
                                                        pack2x2_interleaved(v[8], v[9]), pack2x2_interleaved(v[10], v[11]));
            pack4x2(v, leftCb, leftCr);
            if (own_c && (bs.x & 0xFFFF) && mbx > 0) {
                *reinterpret_cast<uint32_t*>(dCb + (uint32_t)((cy + j) * pitch_c + cx - 4)) = leftCb;
                *reinterpret_cast<uint32_t*>(dCr + (uint32_t)((cy + j) * pitch_c + cx - 4)) = leftCr;
            }
        }

        // ---- post the mailbox of the previous MB: its bottom rows are final now ----
        {
            const uint32_t fa = __shfl_sync(0xFFFFFFFFu, leftA, boxlane_y), fb = __shfl_sync(0xFFFFFFFFu, leftB, boxlane_y);
            const uint32_t fcb = __shfl_sync(0xFFFFFFFFu, leftCb, boxlane_c), fcr = __shfl_sync(0xFFFFFFFFu, leftCr, boxlane_c);
            if (has_below && enabled && mbx > 0) {
                // luma word 2j + 1 is the last word of its row when j is odd: row 12 + (j >> 1) = line (j >> 1) & 1 of lane 6 + (j >> 2)
                st_mbox(box_out + (mbx - 1) * kMboxWords + 2 * j, boxY0, epoch);
                st_mbox(box_out + (mbx - 1) * kMboxWords + 2 * j + 1, (j & 1) ? ((j & 2) ? fb : fa) : boxY1, epoch);
                st_mbox(box_out + (mbx - 1) * kMboxWords + 16 + j, (j & 1) ? ((j & 4) ? fcr : fcb) : boxC, epoch);
            }
        }

        // ---- mailbox of the MB above: normally there already ----
        if (has_above) {
            bool waiting = enabled && ((uint32_t)(t0 >> 32) != epoch || (uint32_t)(t1 >> 32) != epoch || (uint32_t)(t2 >> 32) != epoch);
            unsigned ns = H264R_POLL_NS0;
            while (__any_sync(0xFFFFFFFFu, waiting)) {
                if (waiting) {
                    __nanosleep(ns); if (ns < H264R_POLL_NS1) ns *= 2;
                    t0 = ld_mbox(box_in + mbx * kMboxWords + 2 * j);
                    t1 = ld_mbox(box_in + mbx * kMboxWords + 2 * j + 1);
                    t2 = ld_mbox(box_in + mbx * kMboxWords + 16 + j);
                    waiting = (uint32_t)(t0 >> 32) != epoch || (uint32_t)(t1 >> 32) != epoch || (uint32_t)(t2 >> 32) != epoch;
                }
            }
            reinterpret_cast<uint2*>(sm.top_y)[j] = make_uint2((uint32_t)t0, (uint32_t)t1);    // row 12 + (j >> 1), samples 8 (j & 1) ..
            reinterpret_cast<uint32_t*>(sm.top_c)[j] = (uint32_t)t2;
        }
        __syncwarp();                                      // tile rows (vertical pass) and the rows above visible to the column owners

        // ---- horizontal edges: lane j = luma columns 2j, 2j + 1 and chroma column j of both planes ----
        uint32_t up1 = 0, up2 = 0, up3 = 0, upC = 0;       // samples of the MB above after the top edge
        {
            h2 v[20];
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = has_above ? unpack_u16(*reinterpret_cast<const unsigned short*>(sm.top_y + r * 16 + 2 * j)) : h2c(1024.f);
#pragma unroll
            for (int r = 0; r < 16; ++r) v[4 + r] = unpack_u16(*reinterpret_cast<const unsigned short*>(sm.y + tile_row(r) + 2 * j));
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int s = enabled ? ((e < 2 ? bs.z : bs.w) >> ((e & 1) * 16 + gsh)) & 7 : 0;
                if (s) {
                    h2 alpha, beta, tc0, alpha4;
                    edge_params1(e ? parY1 : parY2, s, alpha, beta, tc0, alpha4);
                    h2 p[4] = { v[4 * e + 3], v[4 * e + 2], v[4 * e + 1], v[4 * e] }, q[4] = { v[4 * e + 4], v[4 * e + 5], v[4 * e + 6], v[4 * e + 7] };
                    filter_edge2<false>(s, alpha, beta, tc0, alpha4, p, q);
                    v[4 * e + 3] = p[0]; v[4 * e + 2] = p[1]; v[4 * e + 1] = p[2];
                    v[4 * e + 4] = q[0]; v[4 * e + 5] = q[1]; v[4 * e + 6] = q[2];
                }
            }
#pragma unroll
            for (int r = 0; r < 15; ++r) *reinterpret_cast<unsigned short*>(sm.y + tile_row(r) + 2 * j) = pack_u16(v[4 + r]);
            up1 = pack_u16(v[1]); up2 = pack_u16(v[2]); up3 = pack_u16(v[3]);        // rows 13..15 of the MB above
        }
        {
            h2 v[10];                                      // rows -2, -1, 0..7 of (Cb, Cr) column j
            v[0] = v[1] = h2c(1024.f);
            if (has_above) {
                v[0] = b2h((__byte_perm(sm.top_c[j], sm.top_c[16 + j], 0x0400) & 0x00FF00FFu) | 0x64006400u);
                v[1] = b2h((__byte_perm(sm.top_c[8 + j], sm.top_c[24 + j], 0x0400) & 0x00FF00FFu) | 0x64006400u);
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) v[2 + r] = unpack_u16(*reinterpret_cast<const unsigned short*>(sm.c + r * 16 + 2 * j));
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int s = enabled ? ((e ? bs.w : bs.z) >> gsh) & 7 : 0;
                if (s) {
                    h2 alpha, beta, tc0;
                    edge_params2(e ? parB1 : parB2, e ? parR1 : parR2, s, alpha, beta, tc0);
                    h2 p[4] = { v[4 * e + 1], v[4 * e], v[4 * e], v[4 * e] }, q[4] = { v[4 * e + 2], v[4 * e + 3], v[4 * e + 3], v[4 * e + 3] };
                    filter_edge2<true>(s, alpha, beta, tc0, tc0, p, q);
                    v[4 * e + 1] = p[0]; v[4 * e + 2] = q[0];
                }
            }
            *reinterpret_cast<unsigned short*>(sm.c + 0 * 16 + 2 * j) = pack_u16(v[2]);
            *reinterpret_cast<unsigned short*>(sm.c + 3 * 16 + 2 * j) = pack_u16(v[5]);
            *reinterpret_cast<unsigned short*>(sm.c + 4 * 16 + 2 * j) = pack_u16(v[6]);
            upC = pack_u16(v[1]);                          // row 7 of the MB above: Cb | Cr << 8
        }
        __syncwarp();                                      // the tile holds the MB after both passes

        // ---- write back: rows 13..15 / row 7 of the MB above (this warp is their only writer), then the MB's own rows ----
        if (has_above && enabled) {
            uint8_t* ty = dY + (uint32_t)((py - 3) * pitch_y + px + 2 * j);
            *reinterpret_cast<unsigned short*>(ty) = (unsigned short)up1;
            *reinterpret_cast<unsigned short*>(ty + pitch_y) = (unsigned short)up2;
            *reinterpret_cast<unsigned short*>(ty + 2 * pitch_y) = (unsigned short)up3;
            dCb[(uint32_t)((cy - 1) * pitch_c + cx + j)] = (uint8_t)upC;
            dCr[(uint32_t)((cy - 1) * pitch_c + cx + j)] = (uint8_t)(upC >> 8);
        }
        if (own_ya) *reinterpret_cast<uint4*>(dY + (uint32_t)((py + 2 * j) * pitch_y + px)) = *reinterpret_cast<const uint4*>(TYa);
        if (own_yb) *reinterpret_cast<uint4*>(dY + (uint32_t)((py + 2 * j + 1) * pitch_y + px)) = *reinterpret_cast<const uint4*>(TYb);
        if (own_c) {
            const uint4 w = *reinterpret_cast<const uint4*>(TCr);                    // Cb / Cr interleaved -> planar
            *reinterpret_cast<uint2*>(dCb + (uint32_t)((cy + j) * pitch_c + cx)) = make_uint2(__byte_perm(w.x, w.y, 0x6420), __byte_perm(w.z, w.w, 0x6420));
            *reinterpret_cast<uint2*>(dCr + (uint32_t)((cy + j) * pitch_c + cx)) = make_uint2(__byte_perm(w.x, w.y, 0x7531), __byte_perm(w.z, w.w, 0x7531));
        }
        // this lane's mailbox words of the MB (their last samples are replaced after the next MB's left edge)
        {
            const uint2 wy = *reinterpret_cast<const uint2*>(boxsrc_y), wc = *reinterpret_cast<const uint2*>(boxsrc_c);
            boxY0 = wy.x; boxY1 = wy.y;
            boxC = __byte_perm(wc.x, wc.y, (j & 4) ? 0x7531 : 0x6420);
        }
        __syncwarp();                                      // before the next vertical pass overwrites the tile rows
    }
    // the last MB of the row has no right neighbour: its bottom rows are final as they stand
    if (has_below && enabled) {
        st_mbox(box_out + (W - 1) * kMboxWords + 2 * j, boxY0, epoch);
        st_mbox(box_out + (W - 1) * kMboxWords + 2 * j + 1, boxY1, epoch);
        st_mbox(box_out + (W - 1) * kMboxWords + 16 + j, boxC, epoch);
    }
}

} // namespace h264r
#endif
