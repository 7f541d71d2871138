// Shared device code of the reconstruction kernels: macroblock header access + domain checks, packed motion access,
// the residual primitives (dequantisation, DC Hadamards, 4x4 / 8x8 inverse transforms: transform.cc:394-456, 460-554,
// 597-733, 825-910) and the deblock-descriptor primitives (boundary strengths and thresholds: deblock.cc:35-289,
// 469-474, tables :294-324).
//
// Residual and deblock descriptors keep their own kernels (residual_kernel, deblock_prep_kernel).  Folding both into the
// reconstruction kernels was built and measured in round 2 (git 43dfbcb, profiles/r2_fused_*): bit-exact, 47 % less DRAM
// traffic, but 8 % MORE instructions (a half warp per MB does the per-MB residual work at 16 lanes, like the residual
// kernel did) and 64 KB of straight-line inter-kernel code that stalls on instruction fetch (no_instruction 5-7 cycles
// per issue): 2.4x slower.  The path is bound by instruction issue, not by HBM; the split keeps each kernel near 32 KB.
#ifndef H264R_KERNELS_COMMON_CUH_
#define H264R_KERNELS_COMMON_CUH_

#include "device_types.h"
#include "mc_core.cuh"

#include <stdint.h>
#include <stddef.h>

namespace h264r {

// Rows of one CTA of the wavefront kernels (row intra, deblock); the CTA takes one ticket for them.  Measured on the
// 64-stream workload, deblock ms per step: 1 row (= a ticket per warp) 10.6, 2 rows 7.73, 4 rows 7.89, 8 rows 8.06.
#ifndef H264R_WARPS_PER_CTA
#define H264R_WARPS_PER_CTA 2
#endif
constexpr int kWarpsPerCta = H264R_WARPS_PER_CTA;

// ---------------------------------------------------------------------------------------------------
// small helpers

__device__ __forceinline__ int clip3i(int lo, int hi, int v) { return min(max(v, lo), hi); }
__device__ __forceinline__ int clip255(int v) { return min(max(v, 0), 255); }

__device__ __forceinline__ uint32_t ldcg_u32(const void* p) { return __ldcg(reinterpret_cast<const unsigned int*>(p)); }
__device__ __forceinline__ uint8_t  ldcg_u8(const uint8_t* p) { return __ldcg(p); }
__device__ __forceinline__ uint32_t ldg_u32(const uint8_t* p) { return __ldg(reinterpret_cast<const unsigned int*>(p)); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

__device__ __forceinline__ int ld_acquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// Domain errors: a kernel that meets a value outside its domain clamps it and ORs a bit into the context's error word
// (host-mapped; only ever touched on the error path).  h264r_wait reports H264R_ERR_INVALID.
enum { ERR_HEADER = 1, ERR_MOTION = 2, ERR_LEVEL = 4 };
__device__ __noinline__ void report_error(uint32_t* err, uint32_t bit) { atomicOr(err, bit); }

// ---------------------------------------------------------------------------------------------------
// macroblock header

struct MbHdr {
    int mb_type, flags, slice_idx, cbp_luma, cbp_chroma, qp_y, qp_c[2], i16mode, cmode, cbp_blks;
    uint32_t coeff_offset, u0, u1;        // u0/u1: the 8-byte union (intra modes | sub_mb_type, sub_mb_pred_mode)
    uint32_t packed;                      // inter MBs: word index of the first motion entry << 4 | layout code
    int coeff_count;
    __device__ __forceinline__ bool intra() const { return flags & H264R_MB_FLAG_INTRA; }
    __device__ __forceinline__ bool t8() const { return flags & H264R_MB_FLAG_T8x8; }
    // the MB has a residual to transform (I_PCM carries samples, not levels)
    __device__ __forceinline__ bool has_resid() const { return coeff_count > 0 && mb_type != H264R_MB_IPCM; }
};

__device__ __forceinline__ MbHdr load_hdr(const h264r_mb* mbs, int addr)
{
    const uint4* p = reinterpret_cast<const uint4*>(mbs + addr);
    uint4 a = __ldg(p), b = __ldg(p + 1);
    MbHdr h;
    h.mb_type = a.x & 0xFF; h.flags = (a.x >> 8) & 0xFF; h.slice_idx = a.x >> 16;
    h.cbp_luma = a.y & 0xFF; h.cbp_chroma = (a.y >> 8) & 0xFF;
    h.qp_y = (int)(int8_t)(a.y >> 16); h.qp_c[0] = (int)(int8_t)(a.y >> 24);
    h.qp_c[1] = (int)(int8_t)(a.z & 0xFF); h.i16mode = (a.z >> 8) & 0xFF; h.cmode = (a.z >> 16) & 0xFF;
    h.cbp_blks = a.w & 0xFFFF; h.coeff_count = a.w >> 16;
    h.coeff_offset = b.x; h.u0 = b.y; h.u1 = b.z; h.packed = b.w;
    return h;
}
// first word only: mb_type | flags << 8 | slice_idx << 16
__device__ __forceinline__ uint32_t load_hdr_word0(const h264r_mb* mbs, int addr)
{
    return __ldg(reinterpret_cast<const unsigned int*>(mbs + addr));
}

// The host no longer walks the macroblocks at submit time.  residual_kernel, which reads every header of a picture before
// any other kernel of the wave does, checks the description once and REPAIRS the HBM copy where a value would lead a
// kernel out of bounds (slice index, QPs, level range, motion index); the other kernels read repaired headers and carry
// no checks of their own beyond index masks.  One warp per MB; lanes share the work.
__device__ __forceinline__ void validate_and_repair(MbHdr& h, const DevPicture& pic, int addr, int lane, uint32_t* err)
{
    const bool intra = h.intra();
    bool bad_hdr = h.slice_idx >= pic.num_slices;
    bad_hdr |= ((unsigned)h.qp_y > 51u) | ((unsigned)h.qp_c[0] > 51u) | ((unsigned)h.qp_c[1] > 51u);
    const bool bad_levels = h.coeff_offset > pic.stream_words || (uint32_t)h.coeff_count > pic.stream_words - h.coeff_offset;
    const bool bad_type = h.mb_type > H264R_MB_IPCM || h.mb_type == 11 || (intra ? h.mb_type < H264R_MB_I4x4 : h.mb_type > H264R_MB_8x8);
    bool bad_motion = false, bad_ref = false;
    if (!intra) {
        const uint32_t code = h.packed & 15u, first = h.packed >> 4;
        const uint32_t n = code == 1 ? 1u : (code == 2 || code == 3 ? 2u : (code == 4 ? 4u : (code == 5 ? 16u : 0u)));
        bad_motion = n == 0 || first > pic.stream_words || 3u * n > pic.stream_words - first;
        if (!bad_motion && (uint32_t)lane < n) {
            const uint32_t rw = __ldg(pic.stream + first + 3u * lane + 2);      // ref_idx[0..1] | ref_pic[0..1] << 16
            const int r0 = (int8_t)(rw >> 16), r1 = (int8_t)(rw >> 24);
            bad_ref = r0 < -1 || r0 >= pic.num_refs || r1 < -1 || r1 >= pic.num_refs;
        }
        bad_ref = __any_sync(0xFFFFFFFFu, bad_ref);
        const uint32_t pm = h.u1;                                               // sub_mb_pred_mode[0..3]
        bad_hdr |= ((pm & 0xFF) > 2u) | (((pm >> 8) & 0xFF) > 2u) | (((pm >> 16) & 0xFF) > 2u) | ((pm >> 24) > 2u);
    }
    if (!(bad_hdr | bad_levels | bad_type | bad_motion | bad_ref)) return;
    h.slice_idx = min(h.slice_idx, pic.num_slices - 1);
    h.qp_y = clip3i(0, 51, h.qp_y); h.qp_c[0] = clip3i(0, 51, h.qp_c[0]); h.qp_c[1] = clip3i(0, 51, h.qp_c[1]);
    if (bad_levels) { h.coeff_count = 0; h.coeff_offset = 0; }
    if (bad_motion) h.packed = pic.stream_words >= 3u ? 1u : 0u;                // entry 0 of the stream, whatever it holds: in bounds
    if (!intra) h.u1 = __vminu4(h.u1, 0x02020202u);
    if (lane == 0) {
        uint32_t* w = reinterpret_cast<uint32_t*>(const_cast<h264r_mb*>(pic.mbs) + addr);
        w[0] = (uint32_t)h.mb_type | (uint32_t)h.flags << 8 | (uint32_t)h.slice_idx << 16;
        w[1] = (uint32_t)h.cbp_luma | (uint32_t)h.cbp_chroma << 8 | (uint32_t)h.qp_y << 16 | (uint32_t)h.qp_c[0] << 24;
        w[2] = (uint32_t)h.qp_c[1] | (uint32_t)h.i16mode << 8 | (uint32_t)h.cmode << 16 | (w[2] & 0xFF000000u);
        w[3] = (uint32_t)h.cbp_blks | (uint32_t)h.coeff_count << 16;
        w[4] = h.coeff_offset; w[6] = h.u1; w[7] = h.packed;
        report_error(err, (bad_hdr | bad_type ? ERR_HEADER : 0u) | (bad_levels ? ERR_LEVEL : 0u) | (bad_motion | bad_ref ? ERR_MOTION : 0u));
    }
}

// Packed motion (h264r_pack_motion): the distinct motion entries of an MB, three stream words each = mv[0], mv[1]
// (int16 x, y), ref_idx[0], ref_idx[1], ref_pic[0], ref_pic[1].  Layout code 1: one entry | 2: rows 0-1 / rows 2-3 |
// 3: columns 0-1 / columns 2-3 | 4: quadrants | 5: all sixteen 4x4 blocks.
__device__ __forceinline__ int packed_entry_within(uint32_t packed, int b)
{
    const int code = packed & 15, row2 = b >> 3, col2 = (b >> 1) & 1;
    return code == 5 ? b : ((code == 2 || code == 4) ? row2 << (code == 4) : 0) + ((code == 3 || code == 4) ? col2 : 0);
}
// word index of the entry that covers block b (in bounds: validate_and_repair)
__device__ __forceinline__ uint32_t packed_entry_word(uint32_t packed, int b)
{
    return (packed >> 4) + 3u * (uint32_t)packed_entry_within(packed, b);
}

// ---------------------------------------------------------------------------------------------------
// residual: dequantisation + DC Hadamard + inverse transform.  Uncoded parts come out as 0, so reconstruction is
// always clip(pred + res) (equal to the reference's "copy prediction" branches, transform.cc:926-934, 1070-1073).
//
// Coefficient scratch of one MB in shared memory (ints): luma 16 rows of pitch 20, chroma 2 planes x 8 rows of pitch 12,
// planes 104 apart.  The padded pitches put the 128-bit row accesses of a quarter warp (eight 4x4 blocks, or the eight
// rows of an 8x8 block) on eight distinct bank groups.
constexpr int kResP = 20, kResCP = 12, kResCPlane = 104, kResC = 16 * kResP, kResInts = kResC + 2 * kResCPlane;

// 4x4 inverse transform in registers: d[row][col] -> residual samples (transform.cc:597-640 itrans4x4 + rounding)
__device__ __forceinline__ void idct4_regs(int (&d)[4][4])
{
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int e0 = d[i][0] + d[i][2], e1 = d[i][0] - d[i][2], e2 = (d[i][1] >> 1) - d[i][3], e3 = d[i][1] + (d[i][3] >> 1);
        d[i][0] = e0 + e3; d[i][1] = e1 + e2; d[i][2] = e1 - e2; d[i][3] = e0 - e3;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int g0 = d[0][j] + d[2][j], g1 = d[0][j] - d[2][j], g2 = (d[1][j] >> 1) - d[3][j], g3 = d[1][j] + (d[3][j] >> 1);
        d[0][j] = (g0 + g3 + 32) >> 6;
        d[1][j] = (g1 + g2 + 32) >> 6;
        d[2][j] = (g1 - g2 + 32) >> 6;
        d[3][j] = (g0 - g3 + 32) >> 6;
    }
}
// a 4x4 block of the scratch <-> registers (rows are 16-byte aligned in both the luma and the chroma layout)
__device__ __forceinline__ void load_block4(const int* blk, int pitch, int (&d)[4][4])
{
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int4 v = *reinterpret_cast<const int4*>(blk + r * pitch);
        d[r][0] = v.x; d[r][1] = v.y; d[r][2] = v.z; d[r][3] = v.w;
    }
}
__device__ __forceinline__ void store_block4(int* blk, int pitch, const int (&d)[4][4])
{
#pragma unroll
    for (int r = 0; r < 4; ++r) *reinterpret_cast<int4*>(blk + r * pitch) = make_int4(d[r][0], d[r][1], d[r][2], d[r][3]);
}

// one 8-point pass of the 8x8 inverse transform (transform.cc:642-733) over p[0], p[stride], ...
__device__ __forceinline__ void idct8_1d(int* p, int stride, bool final_pass)
{
    int d0 = p[0], d1 = p[stride], d2 = p[2 * stride], d3 = p[3 * stride];
    int d4 = p[4 * stride], d5 = p[5 * stride], d6 = p[6 * stride], d7 = p[7 * stride];
    int e0 = d0 + d4;
    int e1 = -d3 + d5 - d7 - (d7 >> 1);
    int e2 = d0 - d4;
    int e3 = d1 + d7 - d3 - (d3 >> 1);
    int e4 = (d2 >> 1) - d6;
    int e5 = -d1 + d7 + d5 + (d5 >> 1);
    int e6 = d2 + (d6 >> 1);
    int e7 = d3 + d5 + d1 + (d1 >> 1);
    int f0 = e0 + e6, f1 = e1 + (e7 >> 2), f2 = e2 + e4, f3 = e3 + (e5 >> 2);
    int f4 = e2 - e4, f5 = (e3 >> 2) - e5, f6 = e0 - e6, f7 = e7 - (e1 >> 2);
    int o0 = f0 + f7, o1 = f2 + f5, o2 = f4 + f3, o3 = f6 + f1, o4 = f6 - f1, o5 = f4 - f3, o6 = f2 - f5, o7 = f0 - f7;
    if (final_pass) {
        o0 = (o0 + 32) >> 6; o1 = (o1 + 32) >> 6; o2 = (o2 + 32) >> 6; o3 = (o3 + 32) >> 6;
        o4 = (o4 + 32) >> 6; o5 = (o5 + 32) >> 6; o6 = (o6 + 32) >> 6; o7 = (o7 + 32) >> 6;
    }
    p[0] = o0; p[stride] = o1; p[2 * stride] = o2; p[3 * stride] = o3;
    p[4 * stride] = o4; p[5 * stride] = o5; p[6 * stride] = o6; p[7 * stride] = o7;
}

// One transmitted level -> the scratch (transform.cc:394-456: the AC levels are dequantised when they are stored, the DC
// levels of Intra16x16 luma and of chroma stay raw until their Hadamard).  Returns the bit of the 4x4 block that received
// the level (0..15 luma raster, 16..23 chroma), 0 if the entry carries none.
//   ctl  = cbp_luma | cbp_chroma << 8 | QpC[0] << 16 | QpC[1] << 24
//   mode = inter | t8 << 1 | i16 << 2 | (QpY / 6) << 8 | (QpY % 6) << 16
__device__ __forceinline__ uint32_t scatter_ctl(const MbHdr& h) { return (uint32_t)h.cbp_luma | (uint32_t)h.cbp_chroma << 8 | (uint32_t)h.qp_c[0] << 16 | (uint32_t)h.qp_c[1] << 24; }
__device__ __forceinline__ uint32_t scatter_mode(const MbHdr& h, int inter)
{
    const int per = h.qp_y / 6, rem = h.qp_y - per * 6;
    return (uint32_t)inter | (h.t8() ? 2u : 0u) | (h.mb_type == H264R_MB_I16x16 ? 4u : 0u) | (uint32_t)per << 8 | (uint32_t)rem << 16;
}
__device__ __forceinline__ unsigned scatter_level(uint32_t e, uint32_t ctl, uint32_t mode, const h264r_slice* __restrict__ sl, int* cof, uint32_t* err)
{
    const int p = (int)(e & 0xFFFFu), l = (int)(int16_t)(e >> 16);
    if (p >= H264R_COEFFS_PER_MB) { report_error(err, ERR_LEVEL); return 0u; }
    if (l == 0) return 0u;
    const int inter = mode & 1, per = (mode >> 8) & 0xFF, rem = mode >> 16;
    int val = 0, at;
    unsigned bit;
    if (p < 256) {
        const int x = p & 15, y = p >> 4;
        at = p + (y << 2);                                   // y * kResP + x
        if (mode & 4u) {
            if (((x | y) & 3) == 0) val = l;
            else val = ((l * (int)__ldg(&sl->level_scale_4x4[0][0][rem][(y & 3) * 4 + (x & 3)])) * (1 << per) + 8) >> 4;
        } else if ((ctl >> ((y >> 3) * 2 + (x >> 3))) & 1) {                             // quirk 6
            if (mode & 2u) val = ((l * (int)__ldg(&sl->level_scale_8x8[inter][rem][(y & 7) * 8 + (x & 7)])) * (1 << per) + 32) >> 6;
            else           val = ((l * (int)__ldg(&sl->level_scale_4x4[inter][0][rem][(y & 3) * 4 + (x & 3)])) * (1 << per) + 8) >> 4;
        }
        bit = 1u << ((y >> 2) * 4 + (x >> 2));
    } else {
        const int c = p - 256, pl = c >> 6, x = c & 7, y = (c >> 3) & 7;
        at = kResC + pl * kResCPlane + y * kResCP + x;
        if (!((ctl >> 8) & 0xFF)) return 0u;
        if (((x | y) & 3) == 0) val = l;
        else {
            const int qc = (ctl >> (16 + 8 * pl)) & 0xFF, cper = qc / 6, crem = qc - cper * 6;
            val = ((l * (int)__ldg(&sl->level_scale_4x4[inter][pl + 1][crem][(y & 3) * 4 + (x & 3)])) * (1 << cper) + 8) >> 4;
        }
        bit = 1u << (16 + pl * 4 + (y >> 2) * 2 + (x >> 2));
    }
    cof[at] = val;
    return bit;
}

// Chroma DC of 4x4 block q (raster, 2x2 blocks) of one plane: the 2x2 Hadamard + dequantisation of transform_chroma_dc
// (transform.cc:858-910) evaluated for this block's position only; c00..c11 = the four raw DC levels of the plane.
__device__ __forceinline__ int chroma_dc_of_block(int q, int c00, int c01, int c10, int c11, int scale, int cper)
{
    const int a = (q & 1) ? c00 - c01 : c00 + c01, b = (q & 1) ? c10 - c11 : c10 + c11;
    const int f = (q & 2) ? a - b : a + b;
    return ((f * scale) * (1 << cper)) >> 5;
}

// eight residual samples -> four int16 pairs clamped to [-255, 255] (clip(pred + res) cannot tell the difference)
__device__ __forceinline__ uint32_t pack_res2(int lo, int hi)
{
    uint32_t pr;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(pr) : "r"(hi), "r"(lo));
    return __vmaxs2(__vmins2(pr, 0x00FF00FFu), 0xFF01FF01u);
}

// ---------------------------------------------------------------------------------------------------
// deblock descriptor

// Tables 8-16 / 8-17 (deblock.cc:294-324) packed for the descriptor: by indexA alpha | tc0[bS=1] << 13 | tc0[2] << 18 |
// tc0[3] << 23, by indexB beta << 8.  Plain global arrays read through the L1 (the indexes differ from lane to lane).
__device__ const uint32_t c_thr_a[52] = {
#define TA(al, t1, t2, t3) ((al) | (t1) << 13 | (t2) << 18 | (t3) << 23)
    TA(0,0,0,0), TA(0,0,0,0), TA(0,0,0,0), TA(0,0,0,0), TA(0,0,0,0), TA(0,0,0,0), TA(0,0,0,0), TA(0,0,0,0),
    TA(0,0,0,0), TA(0,0,0,0), TA(0,0,0,0), TA(0,0,0,0), TA(0,0,0,0), TA(0,0,0,0), TA(0,0,0,0), TA(0,0,0,0),
    TA(4,0,0,0), TA(4,0,0,1), TA(5,0,0,1), TA(6,0,0,1), TA(7,0,0,1), TA(8,0,1,1), TA(9,0,1,1), TA(10,1,1,1),
    TA(12,1,1,1), TA(13,1,1,1), TA(15,1,1,1), TA(17,1,1,2), TA(20,1,1,2), TA(22,1,1,2), TA(25,1,1,2), TA(28,1,2,3),
    TA(32,1,2,3), TA(36,2,2,3), TA(40,2,2,4), TA(45,2,3,4), TA(50,2,3,4), TA(56,3,3,5), TA(63,3,4,6), TA(71,3,4,6),
    TA(80,4,5,7), TA(90,4,5,8), TA(101,4,6,9), TA(113,5,7,10), TA(127,6,8,11), TA(144,6,8,13), TA(162,7,10,14), TA(182,8,11,16),
    TA(203,9,12,18), TA(226,10,13,20), TA(255,11,15,23), TA(255,13,17,25)
#undef TA
};
__device__ const uint8_t c_thr_beta[52] = {
    0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,2,2,3,3,3,3,4,4,4,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13,14,14,15,15,16,16,17,17,18,18 };

// qp of plane pl (0 Y, 1 Cb, 2 Cr) from header words 1 and 2 (cbp_luma | cbp_chroma << 8 | qp_y << 16 | qp_c[0] << 24; qp_c[1] | ...)
__device__ __forceinline__ int qp_of_plane(uint32_t w1, uint32_t w2, int pl)
{
    return (int)(int8_t)(pl == 0 ? w1 >> 16 : (pl == 1 ? w1 >> 24 : w2));
}
// filter thresholds of one (plane, edge type) pair: the qPav / indexA / indexB part of filter_edge (deblock.cc:469-474)
__device__ __forceinline__ uint32_t deblock_threshold_word(int qp_p, int qp_q, int foa, int fob)
{
    const int qPav = (qp_p + qp_q + 1) >> 1;
    const int ia = clip3i(0, 51, qPav + foa), ib = clip3i(0, 51, qPav + fob);
    return __ldg(&c_thr_a[ia]) | (uint32_t)__ldg(&c_thr_beta[ib]) << 8;
}

// |mv_x| differs by four quarter samples or more, or |mv_y| by mvlimit (4; 2 in field pictures, deblock.cc:86, 164); a, b = packed int16 pairs
__device__ __forceinline__ int mv_differs(uint32_t a, uint32_t b, int mvlimit)
{
    const int dx = (int)(int16_t)(a & 0xFFFF) - (int)(int16_t)(b & 0xFFFF), dy = (int)(int16_t)(a >> 16) - (int)(int16_t)(b >> 16);
    return (abs(dx) >= 4) | (abs(dy) >= mvlimit);
}
// bs_compare_mvs, deblock.cc:35-75, on two motion entries held in registers (mv[0], mv[1], ref_idx[0..1] | ref_pic[0..1] << 16)
__device__ __forceinline__ int bs_compare(uint32_t mp0, uint32_t mp1, uint32_t rp, uint32_t mq0, uint32_t mq1, uint32_t rq, int mvlimit)
{
    const int p0 = (int8_t)(rp >> 16), p1 = (int8_t)(rp >> 24), q0 = (int8_t)(rq >> 16), q1 = (int8_t)(rq >> 24);
    if (!((p0 == q0 && p1 == q1) || (p0 == q1 && p1 == q0))) return 1;
    if (p0 != p1) {
        if (p0 == q0) return mv_differs(mp0, mq0, mvlimit) | mv_differs(mp1, mq1, mvlimit);
        return mv_differs(mp0, mq1, mvlimit) | mv_differs(mp1, mq0, mvlimit);
    }
    return (mv_differs(mp0, mq0, mvlimit) | mv_differs(mp1, mq1, mvlimit)) & (mv_differs(mp0, mq1, mvlimit) | mv_differs(mp1, mq0, mvlimit));
}

} // namespace h264r
#endif
