// sm_100a kernels of the H.264 macroblock-reconstruction path.
//
//   residual_kernel           : one warp per MB with levels: dequantisation, DC Hadamards, 4x4 / 8x8 inverse transforms
//                               (transform.cc:394-456, 460-554, 597-733, 825-910) -> int16 residual plane
//   recon_inter2_kernel       : two inter MBs per warp, one 4x4 block per lane, all pictures of a wave in one grid:
//                               motion compensation, weighted prediction, residual add (inter_prediction.cc:53-406,
//                               448-536; decoder.cc:217-262; transform.cc:913-984)
//                               (recon_inter_kernel: the one-MB-per-warp variant, -DH264R_INTER_TWO_MB=0)
//   recon_intra_kernel        : all-intra pictures, one warp per MB ROW; rows form a 2:1 wavefront (MB(x,y) needs (x-1,y),
//                               (x-1,y-1), (x,y-1), (x+1,y-1)) and talk through mailboxes (intra_prediction.cc:137-904)
//   recon_intra_sparse_kernel : the intra MBs of P/B pictures, one warp each, per-MB epoch stamps between intra neighbours
//   deblock_prep_kernel       : boundary strengths and alpha / beta / tc0 per MB (deblock.cc:35-289, 469-474)
//   deblock_kernel            : row wavefront, two pictures per warp, vertical then horizontal edges in place, rows talk
//                               through mailboxes (deblock.cc:327-552)
//
// Arithmetic follows the reference (src/codec/h264/decoder/{transform,inter_prediction,intra_prediction,
// deblock}.cc); the line-by-line citations live in the CPU restatement oracle/port_recon.c, whose structure
// these kernels mirror.  All sample arithmetic is int32; results are bit-exact by construction.
#include "device_types.h"
#include "mc_core.cuh"

#include <stdint.h>
#include <stddef.h>

namespace h264r {

constexpr int kWarpsPerCta = 4;
// resident CTAs per SM the register allocation aims at (tuned on B200, scripts/tune.sh)
#ifndef H264R_INTER_CTAS
#define H264R_INTER_CTAS 10
#endif
#ifndef H264R_INTRA_CTAS
#define H264R_INTRA_CTAS 4
#endif
#ifndef H264R_INTER_TWO_MB
#define H264R_INTER_TWO_MB 1
#endif
#ifndef H264R_RESID_CTAS
#define H264R_RESID_CTAS 14
#endif
#ifndef H264R_PREP_CTAS
#define H264R_PREP_CTAS 12
#endif
#ifndef H264R_PREP_UNROLL
#define H264R_PREP_UNROLL 0
#endif
#ifndef H264R_DEBLOCK_CTAS
#define H264R_DEBLOCK_CTAS 4
#endif

// ---------------------------------------------------------------------------------------------------
// small helpers

__device__ __forceinline__ int clip3i(int lo, int hi, int v) { return min(max(v, lo), hi); }
__device__ __forceinline__ int clip255(int v) { return min(max(v, 0), 255); }
__device__ __forceinline__ int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }

__device__ __forceinline__ uint32_t ldcg_u32(const void* p) { return __ldcg(reinterpret_cast<const unsigned int*>(p)); }
__device__ __forceinline__ uint8_t  ldcg_u8(const uint8_t* p) { return __ldcg(p); }

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

struct MbHdr {
    int mb_type, flags, slice_idx, cbp_luma, cbp_chroma, qp_y, qp_c[2], i16mode, cmode, cbp_blks;
    uint32_t coeff_offset, u0, u1;        // u0/u1: the 8-byte union (intra modes | sub_mb_type, sub_mb_pred_mode)
    uint32_t packed;                      // inter MBs: first packed motion entry << 4 | layout code (engine.cu pack_motion)
    int coeff_count;
    __device__ __forceinline__ bool intra() const { return flags & H264R_MB_FLAG_INTRA; }
    __device__ __forceinline__ bool t8() const { return flags & H264R_MB_FLAG_T8x8; }
    // a residual plane exists for this MB (written by residual_kernel)
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

// Packed motion (engine.cu pack_motion): the distinct motion entries of an MB, 12 bytes each = mv[0], mv[1] (int16 x, y),
// ref_idx[0], ref_idx[1], ref_pic[0], ref_pic[1].  Layout code 1: one entry | 2: rows 0-1 / rows 2-3 | 3: columns 0-1 /
// columns 2-3 | 4: quadrants | 5: all sixteen 4x4 blocks.  Returns the three words of the entry that covers block b.
__device__ __forceinline__ int packed_entry_index(uint32_t packed, int b)
{
    const int code = packed & 15, row2 = b >> 3, col2 = (b >> 1) & 1;
    const int within = code == 5 ? b : ((code == 2 || code == 4) ? row2 << (code == 4) : 0) + ((code == 3 || code == 4) ? col2 : 0);
    return (int)(packed >> 4) + within;
}
__device__ __forceinline__ const uint32_t* packed_entry(const uint8_t* packed_motion, uint32_t packed, int b)
{
    return reinterpret_cast<const uint32_t*>(packed_motion) + (size_t)packed_entry_index(packed, b) * 3;
}

// ---------------------------------------------------------------------------------------------------
// residual: dequantisation + DC Hadamard + inverse transform, whole MB by one warp into res[384] (int32,
// Y 16x16 stride 16 | Cb 8x8 | Cr 8x8).  Uncoded parts come out as 0, so reconstruction is always
// clip(pred + res) (equal to the reference's "copy prediction" branches, transform.cc:926-934, 1070-1073).

__device__ __forceinline__ void idct4_inplace(int* d, int s)
{
    int f[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int d0 = d[i * s], d1 = d[i * s + 1], d2 = d[i * s + 2], d3 = d[i * s + 3];
        int e0 = d0 + d2, e1 = d0 - d2, e2 = (d1 >> 1) - d3, e3 = d1 + (d3 >> 1);
        f[i][0] = e0 + e3; f[i][1] = e1 + e2; f[i][2] = e1 - e2; f[i][3] = e0 - e3;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int f0 = f[0][j], f1 = f[1][j], f2 = f[2][j], f3 = f[3][j];
        int g0 = f0 + f2, g1 = f0 - f2, g2 = (f1 >> 1) - f3, g3 = f1 + (f3 >> 1);
        d[0 * s + j] = (g0 + g3 + 32) >> 6;
        d[1 * s + j] = (g1 + g2 + 32) >> 6;
        d[2 * s + j] = (g1 - g2 + 32) >> 6;
        d[3 * s + j] = (g0 - g3 + 32) >> 6;
    }
}

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

// residual_kernel: one warp per MB that received levels.  Scatters the dequantised levels into shared memory,
// runs the DC Hadamards and the inverse transforms, and writes the MB's 384 residual samples (int16, clamped to
// [-255, 255]: clip(pred + res) cannot tell the difference) to the picture's residual plane.
//
// Shared-memory layout (ints): luma 16 rows of pitch 20, chroma 2 planes x 8 rows of pitch 12, planes 104 apart.  The
// padded pitches put the 128-bit row accesses of a quarter warp (eight 4x4 blocks, or the eight rows of an 8x8 block)
// on eight distinct bank groups; with the raster pitches 16 / 8 they fell two to four on one, and at 14 resident CTAs
// per SM the kernel waits on the shared-memory pipe (ncu v36: mio_throttle 4.5 cycles per issue).
constexpr int kResP = 20, kResCP = 12, kResCPlane = 104, kResC = 16 * kResP, kResInts = kResC + 2 * kResCPlane;
struct __align__(16) ResidSmem { int cof[kResInts]; };

#ifndef H264R_RESID_WARPS
#define H264R_RESID_WARPS 4
#endif
constexpr int kResidWarps = H264R_RESID_WARPS;
__global__ void __launch_bounds__(kResidWarps * 32, H264R_RESID_CTAS * 4 / H264R_RESID_WARPS)
residual_kernel(const DevPicture* __restrict__ pics, FrameGeom g)
{
    __shared__ __align__(16) ResidSmem smem_all[kResidWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nmb = g.width_mbs * g.height_mbs;
    const int addr = blockIdx.x * kResidWarps + warp;     // grid = (ceil(nmb / 4), 1, pictures): no index divisions
    if (addr >= nmb) return;
    const DevPicture& pic = pics[blockIdx.z];
    const MbHdr h = load_hdr(pic.mbs, addr);
    if (!h.has_resid()) return;
    ResidSmem& sm = smem_all[warp];
    int* res = sm.cof;
    const h264r_slice* __restrict__ sl = pic.slices + h.slice_idx;
    const int inter = h.intra() ? 0 : 1;
    const bool t8 = h.t8();
    const bool i16 = h.mb_type == H264R_MB_I16x16;
    const int per = h.qp_y / 6, rem = h.qp_y - per * 6;

#pragma unroll
    for (int k = 0; k < (kResInts / 4 + 31) / 32; ++k)
        if (lane + 32 * k < kResInts / 4) reinterpret_cast<int4*>(res)[lane + 32 * k] = make_int4(0, 0, 0, 0);
    __syncwarp();

    // scatter: dequantise at coeff_luma_ac / coeff_chroma_ac time (transform.cc:394-456); DC levels stay raw
    const h264r_level* __restrict__ lv = pic.levels + h.coeff_offset;
    unsigned nz = 0;                                         // bit b: 4x4 block b (0..15 luma raster, 16..23 chroma) has a level
    for (int i = lane; i < h.coeff_count; i += 32) {
        const uint32_t e = __ldg(lv + i);
        const int p = (int)(e & 0xFFFFu), l = (int)(int16_t)(e >> 16);
        if (p >= 384 || l == 0) continue;
        int val = 0, at;
        if (p < 256) {
            const int x = p & 15, y = p >> 4;
            at = p + (y << 2);                                   // y * kResP + x
            if (i16) {
                if (((x | y) & 3) == 0) val = l;
                else val = ((l * (int)__ldg(&sl->level_scale_4x4[0][0][rem][(y & 3) * 4 + (x & 3)])) * (1 << per) + 8) >> 4;
            } else if ((h.cbp_luma >> ((y >> 3) * 2 + (x >> 3))) & 1) {                      // quirk 6
                if (t8) val = ((l * (int)__ldg(&sl->level_scale_8x8[inter][rem][(y & 7) * 8 + (x & 7)])) * (1 << per) + 32) >> 6;
                else    val = ((l * (int)__ldg(&sl->level_scale_4x4[inter][0][rem][(y & 3) * 4 + (x & 3)])) * (1 << per) + 8) >> 4;
            }
            nz |= 1u << ((y >> 2) * 4 + (x >> 2));
        } else {
            const int c = p - 256, pl = c >> 6, x = c & 7, y = (c >> 3) & 7;
            at = kResC + pl * kResCPlane + y * kResCP + x;
            if (!h.cbp_chroma) continue;
            if (((x | y) & 3) == 0) val = l;
            else {
                const int qc = pl ? h.qp_c[1] : h.qp_c[0], cper = qc / 6, crem = qc - cper * 6;   // (no dynamic index: keeps h in registers)
                val = ((l * (int)__ldg(&sl->level_scale_4x4[inter][pl + 1][crem][(y & 3) * 4 + (x & 3)])) * (1 << cper) + 8) >> 4;
            }
            nz |= 1u << (16 + pl * 4 + (y >> 2) * 2 + (x >> 2));
        }
        res[at] = val;
    }
    nz = __reduce_or_sync(0xFFFFFFFFu, nz);
    __syncwarp();

    // DC transforms (transform_luma_dc :825-856, transform_chroma_dc :858-910)
    if (i16) {
        if (lane == 0) {
            int c[4][4], e[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) c[i][j] = res[i * 4 * kResP + j * 4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int a0 = c[i][0] + c[i][2], a1 = c[i][0] - c[i][2], a2 = c[i][1] - c[i][3], a3 = c[i][1] + c[i][3];
                e[i][0] = a0 + a3; e[i][1] = a1 + a2; e[i][2] = a1 - a2; e[i][3] = a0 - a3;
            }
            const int scale = (int)__ldg(&sl->level_scale_4x4[0][0][rem][0]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int a0 = e[0][j] + e[2][j], a1 = e[0][j] - e[2][j], a2 = e[1][j] - e[3][j], a3 = e[1][j] + e[3][j];
                int f[4] = { a0 + a3, a1 + a2, a1 - a2, a0 - a3 };
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    res[i * 4 * kResP + j * 4] = h.qp_y >= 36 ? (f[i] * scale) * (1 << (per - 6))
                                                       : (f[i] * scale + (1 << (5 - per))) >> (6 - per);
            }
        }
        nz |= 0xFFFFu;                                       // the DC Hadamard spreads into every luma block
    }
    if (h.cbp_chroma && (nz >> 16)) {
        if (lane == 1 || lane == 2) {
            const int pl = lane - 1, qc = pl ? h.qp_c[1] : h.qp_c[0], cper = qc / 6, crem = qc - cper * 6;
            int* c = res + kResC + pl * kResCPlane;             // DC positions (0,0) (0,4) (4,0) (4,4)
            int c00 = c[0], c01 = c[4], c10 = c[4 * kResCP], c11 = c[4 * kResCP + 4];
            int e00 = c00 + c01, e01 = c00 - c01, e10 = c10 + c11, e11 = c10 - c11;
            const int scale = (int)__ldg(&sl->level_scale_4x4[inter][pl + 1][crem][0]);
            c[0]  = (((e00 + e10) * scale) * (1 << cper)) >> 5;
            c[4]  = (((e01 + e11) * scale) * (1 << cper)) >> 5;
            c[4 * kResCP]     = (((e00 - e10) * scale) * (1 << cper)) >> 5;
            c[4 * kResCP + 4] = (((e01 - e11) * scale) * (1 << cper)) >> 5;
        }
        nz |= 0xFF0000u;
    }
    __syncwarp();

    // inverse transforms, only where something is non-zero
    if (t8) {
        const int b = lane >> 3, i = lane & 7;
        int* blk = res + (b >> 1) * 8 * kResP + (b & 1) * 8;
        const unsigned m8 = 0x33u << ((b >> 1) * 8 + (b & 1) * 2);          // the four 4x4 blocks of 8x8 block b
        if (nz & m8) idct8_1d(blk + i * kResP, 1, false);
        __syncwarp();
        if (nz & m8) idct8_1d(blk + i, kResP, true);
        if (lane < 8 && ((nz >> (16 + lane)) & 1)) idct4_inplace(res + kResC + (lane >> 2) * kResCPlane + ((lane >> 1) & 1) * 4 * kResCP + (lane & 1) * 4, kResCP);
    } else if (lane < 24) {
        // one instruction stream for the sixteen luma blocks (lanes 0..15, row pitch 16) and the eight chroma blocks
        const int c = lane - 16;
        int* const blk = lane < 16 ? res + (lane >> 2) * 4 * kResP + (lane & 3) * 4
                                   : res + kResC + (c >> 2) * kResCPlane + ((c >> 1) & 1) * 4 * kResCP + (c & 1) * 4;
        if ((nz >> lane) & 1) idct4_inplace(blk, lane < 16 ? kResP : kResCP);
    }
    __syncwarp();

    // 384 x int16 = 48 x 16 B; saturating pack to int16 pairs, then the [-255, 255] clamp on both halves at once
    uint4* out = reinterpret_cast<uint4*>(pic.resid + (size_t)addr * H264R_COEFFS_PER_MB);
    auto pack8 = [&](const int* r, int v) {              // eight consecutive samples of a row -> one 16-byte store
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t pr;
            asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(pr) : "r"(r[2 * k + 1]), "r"(r[2 * k]));
            w[k] = __vmaxs2(__vmins2(pr, 0x00FF00FFu), 0xFF01FF01u);
        }
        out[v] = make_uint4(w[0], w[1], w[2], w[3]);
    };
    pack8(res + (lane >> 1) * kResP + (lane & 1) * 8, lane);                                       // luma row lane >> 1, half lane & 1
    if (lane < 16) pack8(res + kResC + (lane >> 3) * kResCPlane + (lane & 7) * kResCP, 32 + lane);   // plane lane >> 3, row lane & 7
}

// ---------------------------------------------------------------------------------------------------
// inter prediction

// Decoder::mb_pred_inter partition walk (decoder.cc:217-262) for 4x4 block `blk`: returns the block whose motion
// entry the reference reads (partition origin), the prediction direction, and whether the partition covers the
// whole 8x8 quadrant of the block.  Partition steps in 4x4 units per type 0..7 ({0,0},{4,4},{4,2},{2,4},{2,2},{2,1},
// {1,2},{1,1}) are nibbles of two constants.
__device__ __forceinline__ void partition_of_block(const MbHdr& h, int is_b, int direct_spatial, const uint32_t* refs,
                                                   int direct8x8, int blk, int& origin, int& dir, bool& covers8x8)
{
    const int bx = blk & 3, by = blk >> 2;
    int sh0 = (0x11222440u >> (4 * (h.mb_type & 7))) & 7, sv0 = (0x12124240u >> (4 * (h.mb_type & 7))) & 7;
    if (h.mb_type == 0) sh0 = sv0 = is_b ? 2 : 4;
    const int i0 = bx & ~(sh0 - 1), j0 = by & ~(sv0 - 1);
    const int b8 = 2 * (j0 >> 1) + (i0 >> 1);
    const int mode = (h.u0 >> (8 * b8)) & 0xFF;
    int pd = (h.u1 >> (8 * b8)) & 0xFF;
    int sh4 = (0x11222440u >> (4 * (mode & 7))) & 7, sv4 = (0x12124240u >> (4 * (mode & 7))) & 7;
    if (mode == 0) sh4 = sv4 = direct8x8 ? 2 : 1;
    if (is_b && h.mb_type == H264R_MB_8x8 && direct_spatial) {
        const int b = j0 * 4 + i0;
        const uint32_t rw = refs[b];                       // ref_idx[0] | ref_idx[1] << 8 | ref_pic[0] << 16 | ref_pic[1] << 24
        pd = (int8_t)(rw >> 8) < 0 ? 0 : ((int8_t)rw < 0 ? 1 : 2);
    }
    const int i = bx & ~(sh4 - 1), j = by & ~(sv4 - 1);   // partitions are aligned to their own size
    origin = j * 4 + i;
    dir = pd;
    covers8x8 = sh4 >= 2 && sv4 >= 2;
}

// Reference windows in shared memory.  Interior windows are fetched as aligned 32-bit words: the first sample x0 of
// a window row then sits at byte offset x0 & 3.  Windows touching the picture border (rare) are fetched sample by
// sample with clamped coordinates -- bit-identical to the reference's padded planes + block pre-clamp, SURVEY.md
// 8a -- and start at byte offset 0; that path is kept out of line.
__device__ __noinline__ void load_window_border(uint32_t* win, int pitch_words, const uint8_t* __restrict__ plane, int pitch,
                                                int W, int H, int x0, int y0, int ncols, int nrows, int first_row, int row_step)
{
    for (int row = first_row; row < nrows; row += row_step) {
        const uint8_t* src = plane + (uint32_t)(clip3i(0, H - 1, y0 + row) * pitch);
        uint8_t* dst = reinterpret_cast<uint8_t*>(win + row * pitch_words);
        for (int c = 0; c < ncols; ++c) dst[c] = __ldg(src + clip3i(0, W - 1, x0 + c));
    }
}
__device__ __forceinline__ uint32_t ldg_u32(const uint8_t* p) { return __ldg(reinterpret_cast<const unsigned int*>(p)); }
// 4-byte global -> shared copy that never touches a register (LDGSTS); completion: cp_async_wait_all()
__device__ __forceinline__ void cp_async4(uint32_t* dst, const uint8_t* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// per-warp scratch, in 32-bit words.  Per 8x8 quadrant: luma 146 words = uniform quadrant 13 rows x 4 words | split
// quadrant 4 blocks x (9 rows x 4 words), one row pitch for both so that row offsets are immediates; chroma 50 words =
// uniform 2 planes x (5 rows x 2 words) | split 4 blocks x 2 planes x (3 rows x 2 words), + 1 word the funnel shifts
// may touch.  146 = 2 (mod 8): the four quadrant groups of a warp read disjoint banks.
constexpr int kLumaQ = 146, kChromaQ = 50;
struct __align__(16) InterSmem {
    uint32_t luma[4 * kLumaQ + 2];
    uint32_t chroma[4 * kChromaQ + 2];
    uint32_t mv[2][16];                   // the MB's sixteen motion entries (from the packed form): mv x | y << 16 per list
    uint32_t refs[16];                    // ref_idx[0] | ref_idx[1] << 8 | ref_pic[0] << 16 | ref_pic[1] << 24
};
static_assert(sizeof(InterSmem) % 16 == 0, "InterSmem alignment");

// grid = (ceil(width_mbs / 4), height_mbs, pictures of the wave): one warp per macroblock, no index divisions
__global__ void __launch_bounds__(kWarpsPerCta * 32, H264R_INTER_CTAS)
recon_inter_kernel(const DevPicture* __restrict__ pics, FrameGeom g, int direct8x8)
{
    __shared__ __align__(16) InterSmem smem_all[kWarpsPerCta];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mbx = blockIdx.x * kWarpsPerCta + warp, mby = blockIdx.y;
    if (mbx >= g.width_mbs) return;
    const DevPicture& pic = pics[blockIdx.z];
    if (!pic.has_inter) return;
    const int addr = mby * g.width_mbs + mbx;
    const MbHdr h = load_hdr(pic.mbs, addr);
    if (h.intra()) return;
    InterSmem& sm = smem_all[warp];
    const h264r_slice* __restrict__ sl = pic.slices + h.slice_idx;
    const int wY = g.width_mbs * 16, hY = g.height_mbs * 16, wC = wY >> 1, hC = hY >> 1;

    if (lane < 16) {                                     // the MB's sixteen motion entries, from the packed form
        const uint32_t* e = packed_entry(pic.packed_motion, h.packed, lane);
        const uint32_t m0 = __ldg(e), m1 = __ldg(e + 1), m2 = __ldg(e + 2);
        sm.mv[0][lane] = m0; sm.mv[1][lane] = m1; sm.refs[lane] = m2;
    }
    // slice-level parameters (one 12-byte read, broadcast)
    const uint32_t s0 = __ldg(reinterpret_cast<const uint32_t*>(sl)), s1 = __ldg(reinterpret_cast<const uint32_t*>(sl) + 1),
                   s2 = __ldg(reinterpret_cast<const uint32_t*>(sl) + 2);
    const bool is_b = (s0 & 0xFF) == H264R_B_SLICE;
    const int denom_y = s1 & 0xFF, denom_c = (s1 >> 8) & 0xFF, wp_flag = (s1 >> 16) & 0xFF, bipred_idc = (s1 >> 24) & 0xFF;
    const int direct_spatial = (s2 >> 8) & 0xFF;
    __syncwarp();

    // lanes 8q..8q+7 work on 8x8 quadrant q: lane r = 4x4 block sb = r>>1 of the quadrant, luma rows 2*(r&1)..+1 (a 4x2
    // patch) and the block's 2x2 chroma patch of plane r&1.  If one partition covers the quadrant ("uniform") the eight
    // lanes share one 13x13 luma / 5x5 chroma window; otherwise every block has its own 9x9 / 3x3 window.
    const int q = lane >> 3, r = lane & 7, qx = q & 1, qy = q >> 1;
    const int sb = r >> 1, half = r & 1;
    const int blk = (qy * 2 + (sb >> 1)) * 4 + qx * 2 + (sb & 1);
    int origin, pd; bool uni;
    partition_of_block(h, is_b, direct_spatial, sm.refs, direct8x8, blk, origin, pd, uni);

    // residual of this lane's samples (issued early; consumed at the end)
    const int lx = (blk & 3) * 4, ly = (blk >> 2) * 4 + half * 2;          // luma position in the MB
    const int cxx = (blk & 3) * 2, cyy = (blk >> 2) * 2;                    // chroma position in the MB (plane `half`)
    uint2 resY0 = make_uint2(0, 0), resY1 = make_uint2(0, 0); uint32_t resC0 = 0, resC1 = 0;
    if (h.has_resid()) {
        const int16_t* __restrict__ rs = pic.resid + (size_t)addr * H264R_COEFFS_PER_MB;
        resY0 = __ldg(reinterpret_cast<const uint2*>(rs + ly * 16 + lx));
        resY1 = __ldg(reinterpret_cast<const uint2*>(rs + (ly + 1) * 16 + lx));
        resC0 = __ldg(reinterpret_cast<const uint32_t*>(rs + 256 + half * 64 + cyy * 8 + cxx));
        resC1 = __ldg(reinterpret_cast<const uint32_t*>(rs + 256 + half * 64 + (cyy + 1) * 8 + cxx));
    }

    uint32_t* const lq = sm.luma + q * kLumaQ;
    uint32_t* const cq = sm.chroma + q * kChromaQ;
    const int pitch_y = g.pitch_y, pitch_c = g.pitch_c;

    // samples of the (up to) two lists, packed bytes: cur = last list done, prev = the one before
    uint32_t curY0 = 0, curY1 = 0, curC = 0, prevY0 = 0, prevY1 = 0, prevC = 0;
    int ref_cur = 0, ref_prev = 0;
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {
        const bool active = k == 0 || pd == 2;
        if (k == 1 && !__any_sync(0xFFFFFFFFu, active)) break;
        const int list = pd == 2 ? k : pd;
        int vx = 0, vy = 0, refidx = 0;
        const uint32_t* wl = lq; const uint32_t* wc = cq;
        int loff = 2, coff = 0;
        if (active) {
            // the entry names the reference picture itself (ref_pic = slot of pic_params.ref_frames, the identity the
            // deblocking rule compares): RefPicList[list][ref_idx] resolved by the parser side, one load less in the chain
            const uint32_t rw = sm.refs[origin], mvw = sm.mv[list][origin];
            refidx = (int)(int8_t)(rw >> (8 * list));
            const int slot = (int)(int8_t)(rw >> (16 + 8 * list));
            const uint8_t* __restrict__ rbase = pic.ref[slot & 31];
            const int mvx = (int)(int16_t)(mvw & 0xFFFF), mvy = (int)(int16_t)(mvw >> 16);
            vx = (mbx * 16 + (blk & 3) * 4) * 4 + mvx; vy = (mby * 16 + (blk >> 2) * 4) * 4 + mvy;   // this block's position
            if (uni) {
                const int qvx = (mbx * 16 + qx * 8) * 4 + mvx, qvy = (mby * 16 + qy * 8) * 4 + mvy;
                const int x0 = (qvx >> 2) - 2, y0 = (qvy >> 2) - 2, cx0 = qvx >> 3, cy0 = qvy >> 3;
                const int xa = x0 & ~3, cxa = cx0 & ~3;
                const bool in_y = xa >= 0 && xa + 16 <= wY && y0 >= 0 && y0 + 13 <= hY;
                const bool in_c = cxa >= 0 && cxa + 8 <= wC && cy0 >= 0 && cy0 + 5 <= hC;
                if (in_y) {                                     // 13 rows x 4 words: lane = (row parity, word)
                    const int col = r & 3, rsel = r >> 2;
                    const uint8_t* src = rbase + (uint32_t)((y0 + rsel) * pitch_y + xa + col * 4);
                    uint32_t* dst = lq + rsel * 4 + col;
                    uint32_t v[7];
#pragma unroll
                    for (int i = 0; i < 7; ++i) if (i < 6 || rsel == 0) v[i] = ldg_u32(src + (uint32_t)(i * 2 * pitch_y));
#pragma unroll
                    for (int i = 0; i < 7; ++i) if (i < 6 || rsel == 0) dst[i * 8] = v[i];
                } else load_window_border(lq, 4, rbase, pitch_y, wY, hY, x0, y0, 13, 13, r, 8);
                if (in_c) {                                     // 2 planes x 5 rows x 2 words: lane = (row parity, plane, word)
                    const int col = r & 1, pl = (r >> 1) & 1, rsel = r >> 2;
                    const uint8_t* src = rbase + (pl ? g.off_cr : g.off_cb) + (uint32_t)((cy0 + rsel) * pitch_c + cxa + col * 4);
                    uint32_t* dst = cq + pl * 10 + rsel * 2 + col;
                    uint32_t v[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) if (i < 2 || rsel == 0) v[i] = ldg_u32(src + (uint32_t)(i * 2 * pitch_c));
#pragma unroll
                    for (int i = 0; i < 3; ++i) if (i < 2 || rsel == 0) dst[i * 4] = v[i];
                } else {
                    const int pl = r & 1;
                    load_window_border(cq + pl * 10, 2, rbase + (pl ? g.off_cr : g.off_cb), pitch_c, wC, hC, cx0, cy0, 5, 5, r >> 1, 4);
                }
                wl = lq + ((sb >> 1) * 4 + half * 2) * 4;
                loff = 2 + (sb & 1) * 4 + (in_y ? x0 & 3 : 0);
                wc = cq + half * 10 + (sb >> 1) * 2 * 2;
                coff = (sb & 1) * 2 + (in_c ? cx0 & 3 : 0);
            } else {
                const int x0 = (vx >> 2) - 2, y0 = (vy >> 2) - 2, cx0 = vx >> 3, cy0 = vy >> 3;
                const int xa = x0 & ~3, cxa = cx0 & ~3;
                const bool in_y = xa >= 0 && xa + 12 <= wY && y0 >= 0 && y0 + 9 <= hY;
                const bool in_c = cxa >= 0 && cxa + 8 <= wC && cy0 >= 0 && cy0 + 3 <= hC;
                uint32_t* const lb = lq + sb * 36;
                uint32_t* const cb = cq + sb * 12 + half * 6;
                const uint8_t* const cplane = rbase + (half ? g.off_cr : g.off_cb);
                if (in_y) {                                     // 9 rows x 3 words: the two lanes of the block take alternate rows
                    const uint8_t* src = rbase + (uint32_t)((y0 + half) * pitch_y + xa);
                    uint32_t* dst = lb + half * 4;
                    uint32_t v[5][3];
#pragma unroll
                    for (int i = 0; i < 5; ++i) if (i < 4 || half == 0)
#pragma unroll
                        for (int c = 0; c < 3; ++c) v[i][c] = ldg_u32(src + (uint32_t)(i * 2 * pitch_y) + c * 4);
#pragma unroll
                    for (int i = 0; i < 5; ++i) if (i < 4 || half == 0)
#pragma unroll
                        for (int c = 0; c < 3; ++c) dst[i * 8 + c] = v[i][c];
                } else load_window_border(lb, 4, rbase, pitch_y, wY, hY, x0, y0, 9, 9, half, 2);
                if (in_c) {                                     // 3 rows x 2 words of this lane's plane
                    const uint8_t* src = cplane + (uint32_t)(cy0 * pitch_c + cxa);
                    uint32_t v[3][2];
#pragma unroll
                    for (int i = 0; i < 3; ++i) { v[i][0] = ldg_u32(src + (uint32_t)(i * pitch_c)); v[i][1] = ldg_u32(src + (uint32_t)(i * pitch_c) + 4); }
#pragma unroll
                    for (int i = 0; i < 3; ++i) { cb[i * 2] = v[i][0]; cb[i * 2 + 1] = v[i][1]; }
                } else load_window_border(cb, 2, cplane, pitch_c, wC, hC, cx0, cy0, 3, 3, 0, 1);
                wl = lb + half * 2 * 4;
                loff = 2 + (in_y ? x0 & 3 : 0);
                wc = cb;
                coff = in_c ? cx0 & 3 : 0;
            }
        }
        __syncwarp();
        {
            // stage masks of the warp: every lane runs the stages some lane needs (warp-uniform branches, no divergence
            // bookkeeping); lanes without a second list contribute nothing and keep their samples
            const int xf = vx & 3, yf = vy & 3;
            unsigned hm, cm;
            mc_luma_masks(xf, yf, hm, cm);
            const unsigned whm = __reduce_or_sync(0xFFFFFFFFu, active ? hm : 0u), wcm = __reduce_or_sync(0xFFFFFFFFu, active ? cm : 0u);
            uint32_t y0, y1;
            mc_luma_patch_4x2(wl, loff, xf, yf, whm, wcm, y0, y1);
            const uint32_t c = mc_chroma_patch_2x2(wc, coff, vx & 7, vy & 7);
            if (active) {
                prevY0 = curY0; prevY1 = curY1; prevC = curC; ref_prev = ref_cur; ref_cur = refidx;
                curY0 = y0; curY1 = y1; curC = c;
            }
        }
        __syncwarp();
    }

    // weighted sample prediction (mc_prediction / bi_prediction, inter_prediction.cc:53-156), residual, store
    const bool uni_weighted = (wp_flag && !is_b) || (bipred_idc == 1 && is_b);
    const int ref0 = pd == 2 ? ref_prev : ref_cur, ref1 = ref_cur;
    const int mode = pd != 2 ? (uni_weighted ? 1 : 0) : (bipred_idc == 0 ? 2 : 3);
    const uint32_t p0Y0 = pd == 2 ? prevY0 : curY0, p0Y1 = pd == 2 ? prevY1 : curY1, p0C = pd == 2 ? prevC : curC;
    uint32_t outY0, outY1, outC;
    {
        int wgt[2][2] = { { 0, 0 }, { 0, 0 } }, off[2] = { 0, 0 };                 // [luma, chroma plane `half`][list]
        if (mode == 1) {
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                const int pl = part ? 1 + half : 0;
                wgt[part][0] = (int)(int8_t)__ldg(&sl->wp_weight[pd][pl][ref0 & 31]);
                off[part] = (int)(int8_t)__ldg(&sl->wp_offset[pd][pl][ref0 & 31]);
            }
        } else if (mode == 3) {
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                const int pl = part ? 1 + half : 0;
                if (bipred_idc == 1) {
                    wgt[part][0] = (int)(int8_t)__ldg(&sl->wp_weight[0][pl][ref0 & 31]);
                    wgt[part][1] = (int)(int8_t)__ldg(&sl->wp_weight[1][pl][ref1 & 31]);
                    off[part] = ((int)(int8_t)__ldg(&sl->wp_offset[0][pl][ref0 & 31]) + (int)(int8_t)__ldg(&sl->wp_offset[1][pl][ref1 & 31]) + 1) >> 1;
                } else {
                    wgt[part][1] = (int)__ldg(&sl->implicit_w1[ref0 & 31][ref1 & 31]);
                    wgt[part][0] = 64 - wgt[part][1];
                }
            }
        }
        outY0 = mc_weight_recon4(mode, p0Y0, curY0, wgt[0][0], wgt[0][1], denom_y, off[0], resY0.x, resY0.y);
        outY1 = mc_weight_recon4(mode, p0Y1, curY1, wgt[0][0], wgt[0][1], denom_y, off[0], resY1.x, resY1.y);
        outC  = mc_weight_recon4(mode, p0C,  curC,  wgt[1][0], wgt[1][1], denom_c, off[1], resC0, resC1);
    }
    uint8_t* dY = pic.dst + (uint32_t)((mby * 16 + ly) * pitch_y + mbx * 16 + lx);
    uint8_t* dC = pic.dst + (half ? g.off_cr : g.off_cb) + (uint32_t)((mby * 8 + cyy) * pitch_c + mbx * 8 + cxx);
    *reinterpret_cast<uint32_t*>(dY) = outY0;
    *reinterpret_cast<uint32_t*>(dY + pitch_y) = outY1;
    *reinterpret_cast<uint16_t*>(dC) = (uint16_t)(outC & 0xFFFF);
    *reinterpret_cast<uint16_t*>(dC + pitch_c) = (uint16_t)(outC >> 16);
}

// ---------------------------------------------------------------------------------------------------
// inter prediction, two macroblocks per warp
//
// Lanes 0..15 reconstruct MB 2j, lanes 16..31 MB 2j+1 of a row; lane b of a half owns 4x4 luma block b (raster) and the
// 2x2 chroma patches of both planes under it.  Against one-MB-per-warp with 4x2 patches: the per-MB work that is the
// same for every lane (header, slice, partition walk, addressing, weights, loop control: two thirds of that kernel's
// instructions) is issued once for two MBs, and a 4x4 patch filters 9 window rows for 4 output rows where two 4x2
// patches filter 14.  If one partition covers an 8x8 quadrant its four lanes share one 13x13 luma / 5x5 chroma window,
// otherwise every block has its own 9x9 / 3x3 window (same window layout per quadrant as above).
struct __align__(16) Inter2Smem {
    uint32_t luma[2][4 * kLumaQ + 2];
    uint32_t chroma[2][4 * kChromaQ + 2];
};

__device__ __forceinline__ void partition_of_block2(const MbHdr& h, int is_b, int direct_spatial, const uint8_t* pm, int direct8x8,
                                                    int blk, int& origin, int& dir, bool& covers8x8)
{
    const int bx = blk & 3, by = blk >> 2;
    int sh0 = (0x11222440u >> (4 * (h.mb_type & 7))) & 7, sv0 = (0x12124240u >> (4 * (h.mb_type & 7))) & 7;
    if (h.mb_type == 0) sh0 = sv0 = is_b ? 2 : 4;
    const int i0 = bx & ~(sh0 - 1), j0 = by & ~(sv0 - 1);
    const int b8 = 2 * (j0 >> 1) + (i0 >> 1);
    const int mode = (h.u0 >> (8 * b8)) & 0xFF;
    int pd = (h.u1 >> (8 * b8)) & 0xFF;
    int sh4 = (0x11222440u >> (4 * (mode & 7))) & 7, sv4 = (0x12124240u >> (4 * (mode & 7))) & 7;
    if (mode == 0) sh4 = sv4 = direct8x8 ? 2 : 1;
    if (is_b && h.mb_type == H264R_MB_8x8 && direct_spatial) {
        const uint32_t rw = __ldg(packed_entry(pm, h.packed, j0 * 4 + i0) + 2);
        pd = (int8_t)(rw >> 8) < 0 ? 0 : ((int8_t)rw < 0 ? 1 : 2);
    }
    const int i = bx & ~(sh4 - 1), j = by & ~(sv4 - 1);   // partitions are aligned to their own size
    origin = j * 4 + i;
    dir = pd;
    covers8x8 = sh4 >= 2 && sv4 >= 2;
}

#ifndef H264R_INTER2_WARPS
#define H264R_INTER2_WARPS 2
#endif
#ifndef H264R_INTER2_CTAS
#define H264R_INTER2_CTAS (28 / H264R_INTER2_WARPS)
#endif
constexpr int kInter2Warps = H264R_INTER2_WARPS;       // warps per CTA (each warp: two MBs)
// grid = (ceil(width_mbs / 8), height_mbs, pictures of the wave)
__global__ void __launch_bounds__(kInter2Warps * 32, H264R_INTER2_CTAS)
recon_inter2_kernel(const DevPicture* __restrict__ pics, FrameGeom g, int direct8x8)
{
    __shared__ __align__(16) Inter2Smem smem_all[kInter2Warps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = lane >> 4, b = lane & 15, bx = b & 3, by = b >> 2;
    const int mbx = (blockIdx.x * kInter2Warps + warp) * 2 + m, mby = blockIdx.y;
    const DevPicture& pic = pics[blockIdx.z];
    if (!pic.has_inter) return;
    const int W = g.width_mbs;
    if ((blockIdx.x * kInter2Warps + warp) * 2 >= W) return;
    const int addr = mby * W + min(mbx, W - 1);
    const MbHdr h = load_hdr(pic.mbs, addr);
    const bool valid = mbx < W && !h.intra();
    if (!__any_sync(0xFFFFFFFFu, valid)) return;
    Inter2Smem& sm = smem_all[warp];
    const h264r_slice* __restrict__ sl = pic.slices + h.slice_idx;
    const int wY = W * 16, hY = g.height_mbs * 16, wC = wY >> 1, hC = hY >> 1;
    const uint32_t s0 = __ldg(reinterpret_cast<const uint32_t*>(sl)), s1 = __ldg(reinterpret_cast<const uint32_t*>(sl) + 1),
                   s2 = __ldg(reinterpret_cast<const uint32_t*>(sl) + 2);
    const bool is_b = (s0 & 0xFF) == H264R_B_SLICE;
    const int denom_y = s1 & 0xFF, denom_c = (s1 >> 8) & 0xFF, wp_flag = (s1 >> 16) & 0xFF, bipred_idc = (s1 >> 24) & 0xFF;
    const int direct_spatial = (s2 >> 8) & 0xFF;

    // the residual is needed last: pull its six lines into the L2 now (no registers held), the loads at the end then
    // cost an L2 hit instead of one more HBM round trip on the warp's dependent chain
    if (valid && h.has_resid() && b < 6) prefetch_l2(pic.resid + (size_t)addr * H264R_COEFFS_PER_MB + b * 64);
    int origin = 0, pd = 0; bool uni = true;
    uint32_t mvw0 = 0, mvw1 = 0, rw = 0;                 // the motion entry of this block's partition
    if (valid) {
        partition_of_block2(h, is_b, direct_spatial, pic.packed_motion, direct8x8, b, origin, pd, uni);
        const uint32_t* e = packed_entry(pic.packed_motion, h.packed, origin);
        mvw0 = __ldg(e); mvw1 = __ldg(e + 1); rw = __ldg(e + 2);
    }
    const int q = (by >> 1) * 2 + (bx >> 1), sb = (by & 1) * 2 + (bx & 1);        // quadrant, block inside the quadrant
    uint32_t* const lq = sm.luma[m] + q * kLumaQ;
    uint32_t* const cq = sm.chroma[m] + q * kChromaQ;
    const int pitch_y = g.pitch_y, pitch_c = g.pitch_c;

    // samples of the (up to) two lists, packed bytes: cur = last list done, prev = the one before
    uint32_t curY[4] = { 0, 0, 0, 0 }, prevY[4] = { 0, 0, 0, 0 }, curC[2] = { 0, 0 }, prevC[2] = { 0, 0 };
    int ref_cur = 0, ref_prev = 0;
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {
        const bool active = valid && (k == 0 || pd == 2);
        if (k == 1 && !__any_sync(0xFFFFFFFFu, active)) break;
        const int list = pd == 2 ? k : pd;
        int vx = 0, vy = 0, refidx = 0;
        const uint32_t* wl = lq; const uint32_t* wc0 = cq; const uint32_t* wc1 = cq;
        int loff = 2, coff = 0;
        if (active) {
            refidx = (int)(int8_t)(rw >> (8 * list));
            const int slot = (int)(int8_t)(rw >> (16 + 8 * list));
            const uint8_t* __restrict__ rbase = pic.ref[slot & 31];
            const uint32_t mvw = list ? mvw1 : mvw0;
            const int mvx = (int)(int16_t)(mvw & 0xFFFF), mvy = (int)(int16_t)(mvw >> 16);
            vx = (mbx * 16 + bx * 4) * 4 + mvx; vy = (mby * 16 + by * 4) * 4 + mvy;       // this block's position
            if (uni) {
                const int qvx = (mbx * 16 + (bx >> 1) * 8) * 4 + mvx, qvy = (mby * 16 + (by >> 1) * 8) * 4 + mvy;
                const int x0 = (qvx >> 2) - 2, y0 = (qvy >> 2) - 2, cx0 = qvx >> 3, cy0 = qvy >> 3;
                const int xa = x0 & ~3, cxa = cx0 & ~3;
                const bool in_y = xa >= 0 && xa + 16 <= wY && y0 >= 0 && y0 + 13 <= hY;
                const bool in_c = cxa >= 0 && cxa + 8 <= wC && cy0 >= 0 && cy0 + 5 <= hC;
                if (in_y) {                                     // 13 rows x 4 words: lane = word column
                    const uint8_t* src = rbase + (uint32_t)(y0 * pitch_y + xa + sb * 4);
#pragma unroll
                    for (int i = 0; i < 13; ++i) cp_async4(lq + i * 4 + sb, src + (uint32_t)(i * pitch_y));
                } else load_window_border(lq, 4, rbase, pitch_y, wY, hY, x0, y0, 13, 13, sb, 4);
                {                                               // 2 planes x 5 rows x 2 words: lane = (plane, word column)
                    const int pl = sb >> 1, col = sb & 1;
                    const uint8_t* cplane = rbase + (pl ? g.off_cr : g.off_cb);
                    if (in_c) {
                        const uint8_t* src = cplane + (uint32_t)(cy0 * pitch_c + cxa + col * 4);
#pragma unroll
                        for (int i = 0; i < 5; ++i) cp_async4(cq + pl * 10 + i * 2 + col, src + (uint32_t)(i * pitch_c));
                    } else load_window_border(cq + pl * 10, 2, cplane, pitch_c, wC, hC, cx0, cy0, 5, 5, col, 2);
                }
                wl = lq + (sb >> 1) * 4 * 4;
                loff = 2 + (sb & 1) * 4 + (in_y ? x0 & 3 : 0);
                wc0 = cq + (sb >> 1) * 2 * 2; wc1 = wc0 + 10;
                coff = (sb & 1) * 2 + (in_c ? cx0 & 3 : 0);
            } else {
                const int x0 = (vx >> 2) - 2, y0 = (vy >> 2) - 2, cx0 = vx >> 3, cy0 = vy >> 3;
                const int xa = x0 & ~3, cxa = cx0 & ~3;
                const bool in_y = xa >= 0 && xa + 12 <= wY && y0 >= 0 && y0 + 9 <= hY;
                const bool in_c = cxa >= 0 && cxa + 8 <= wC && cy0 >= 0 && cy0 + 3 <= hC;
                uint32_t* const lb = lq + sb * 36;
                uint32_t* const cb = cq + sb * 12;
                if (in_y) {                                     // 9 rows x 3 words
                    const uint8_t* src = rbase + (uint32_t)(y0 * pitch_y + xa);
#pragma unroll
                    for (int i = 0; i < 9; ++i)
#pragma unroll
                        for (int c = 0; c < 3; ++c) cp_async4(lb + i * 4 + c, src + (uint32_t)(i * pitch_y) + c * 4);
                } else load_window_border(lb, 4, rbase, pitch_y, wY, hY, x0, y0, 9, 9, 0, 1);
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {                // 3 rows x 2 words per plane
                    const uint8_t* cplane = rbase + (pl ? g.off_cr : g.off_cb);
                    if (in_c) {
                        const uint8_t* src = cplane + (uint32_t)(cy0 * pitch_c + cxa);
#pragma unroll
                        for (int i = 0; i < 3; ++i) { cp_async4(cb + pl * 6 + i * 2, src + (uint32_t)(i * pitch_c)); cp_async4(cb + pl * 6 + i * 2 + 1, src + (uint32_t)(i * pitch_c) + 4); }
                    } else load_window_border(cb + pl * 6, 2, cplane, pitch_c, wC, hC, cx0, cy0, 3, 3, 0, 1);
                }
                wl = lb;
                loff = 2 + (in_y ? x0 & 3 : 0);
                wc0 = cb; wc1 = cb + 6;
                coff = in_c ? cx0 & 3 : 0;
            }
        }
        cp_async_wait_all();                               // this lane's window copies have landed
        __syncwarp();
        {
            const int xf = vx & 3, yf = vy & 3;
            unsigned hm, cm;
            mc_luma_masks_r<4>(xf, yf, hm, cm);
            const unsigned whm = __reduce_or_sync(0xFFFFFFFFu, active ? hm : 0u), wcm = __reduce_or_sync(0xFFFFFFFFu, active ? cm : 0u);
            uint32_t y[4];
            mc_luma_patch<4>(wl, loff, xf, yf, whm, wcm, y);
            const uint32_t c0 = mc_chroma_patch_2x2(wc0, coff, vx & 7, vy & 7), c1 = mc_chroma_patch_2x2(wc1, coff, vx & 7, vy & 7);
            if (active) {
#pragma unroll
                for (int r = 0; r < 4; ++r) { prevY[r] = curY[r]; curY[r] = y[r]; }
                prevC[0] = curC[0]; prevC[1] = curC[1]; curC[0] = c0; curC[1] = c1;
                ref_prev = ref_cur; ref_cur = refidx;
            }
        }
        __syncwarp();
    }
    if (!valid) return;

    // weighted sample prediction (mc_prediction / bi_prediction, inter_prediction.cc:53-156), residual, store
    const bool uni_weighted = (wp_flag && !is_b) || (bipred_idc == 1 && is_b);
    const int ref0 = pd == 2 ? ref_prev : ref_cur, ref1 = ref_cur;
    const int mode = pd != 2 ? (uni_weighted ? 1 : 0) : (bipred_idc == 0 ? 2 : 3);
    int wgt[3][2] = { { 0, 0 }, { 0, 0 }, { 0, 0 } }, off[3] = { 0, 0, 0 };             // [Y, Cb, Cr][list]
    if (mode == 1) {
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
            wgt[pl][0] = (int)(int8_t)__ldg(&sl->wp_weight[pd][pl][ref0 & 31]);
            off[pl] = (int)(int8_t)__ldg(&sl->wp_offset[pd][pl][ref0 & 31]);
        }
    } else if (mode == 3) {
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
            if (bipred_idc == 1) {
                wgt[pl][0] = (int)(int8_t)__ldg(&sl->wp_weight[0][pl][ref0 & 31]);
                wgt[pl][1] = (int)(int8_t)__ldg(&sl->wp_weight[1][pl][ref1 & 31]);
                off[pl] = ((int)(int8_t)__ldg(&sl->wp_offset[0][pl][ref0 & 31]) + (int)(int8_t)__ldg(&sl->wp_offset[1][pl][ref1 & 31]) + 1) >> 1;
            } else {
                wgt[pl][1] = (int)__ldg(&sl->implicit_w1[ref0 & 31][ref1 & 31]);
                wgt[pl][0] = 64 - wgt[pl][1];
            }
        }
    }
    const bool has_res = h.has_resid();
    const int16_t* __restrict__ rs = pic.resid + (size_t)addr * H264R_COEFFS_PER_MB;
    uint8_t* dY = pic.dst + (uint32_t)((mby * 16 + by * 4) * pitch_y + mbx * 16 + bx * 4);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint2 res = has_res ? __ldg(reinterpret_cast<const uint2*>(rs + (by * 4 + r) * 16 + bx * 4)) : make_uint2(0, 0);
        const uint32_t p0 = pd == 2 ? prevY[r] : curY[r];
        *reinterpret_cast<uint32_t*>(dY + (uint32_t)(r * pitch_y)) = mc_weight_recon4(mode, p0, curY[r], wgt[0][0], wgt[0][1], denom_y, off[0], res.x, res.y);
    }
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
        uint32_t r0 = 0, r1 = 0;
        if (has_res) {
            r0 = __ldg(reinterpret_cast<const uint32_t*>(rs + 256 + pl * 64 + (by * 2) * 8 + bx * 2));
            r1 = __ldg(reinterpret_cast<const uint32_t*>(rs + 256 + pl * 64 + (by * 2 + 1) * 8 + bx * 2));
        }
        const uint32_t p0 = pd == 2 ? prevC[pl] : curC[pl];
        const uint32_t o = mc_weight_recon4(mode, p0, curC[pl], wgt[1 + pl][0], wgt[1 + pl][1], denom_c, off[1 + pl], r0, r1);
        uint8_t* dC = pic.dst + (pl ? g.off_cr : g.off_cb) + (uint32_t)((mby * 8 + by * 2) * pitch_c + mbx * 8 + bx * 2);
        *reinterpret_cast<uint16_t*>(dC) = (uint16_t)(o & 0xFFFF);
        *reinterpret_cast<uint16_t*>(dC + pitch_c) = (uint16_t)(o >> 16);
    }
}

// ---------------------------------------------------------------------------------------------------
// row wavefront plumbing

// Mailbox word: 4 samples + the launch epoch in one 64-bit store / load (single-copy atomic): the data arrives with its
// own flag, so neither side needs a fence (the low-latency protocol of collective libraries).  Epochs make clearing
// unnecessary; the intra wavefront tags its words with bit 31 so that they never pass for deblock words of the same wave.
__device__ __forceinline__ void st_mbox(uint64_t* p, uint32_t data, uint32_t epoch)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"((uint64_t)data | ((uint64_t)epoch << 32)) : "memory");
}
__device__ __forceinline__ uint64_t ld_mbox(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}


// ---------------------------------------------------------------------------------------------------
// intra prediction (wavefront)

// luma tile: rows -1..15, cols -4..27 -> index (y+1)*32 + (x+4); 17 rows x 32 B
// chroma tile per plane: rows -1..7, cols -4..11 -> index (y+1)*16 + (x+4); 9 rows x 16 B
struct __align__(16) IntraSmem {
    __align__(16) int16_t res[384];              // this MB's residual (zero when it has none)
    __align__(16) uint8_t ty[17 * 32];
    __align__(16) uint8_t tc[2][9 * 16];
    __align__(4) uint8_t f8[32];                 // Intra8x8 filtered reference samples p': [7 - i] = p'(-1, i), [8] = p'(-1, -1), [12 + i] = p'(i, -1)
};

#define TY(x, y) sm.ty[((y) + 1) * 32 + (x) + 4]
#define TC(pl, x, y) sm.tc[pl][((y) + 1) * 16 + (x) + 4]

__device__ __forceinline__ bool nb_avail(const h264r_mb* mbs, int W, int H, int cur, uint32_t cur_w0, int nx, int ny, bool need_intra)
{
    if (nx < 0 || nx >= W || ny < 0 || ny >= H) return false;
    const int nb = ny * W + nx;
    if (nb >= cur) return false;
    const uint32_t w0 = load_hdr_word0(mbs, nb);
    if ((w0 >> 16) != (cur_w0 >> 16)) return false;
    if (need_intra && !((w0 >> 8) & H264R_MB_FLAG_INTRA)) return false;
    return true;
}

// One of the nine directional predictors at sample (x, y) of an n x n block.  T(i), L(i): reference samples
// with T(-1) == L(-1) the corner; tmax = last valid top index (2n-1, or n-1 when C is substituted).
template <typename TF, typename LF>
__device__ __forceinline__ int pred_dir_sample(int mode, int n, int x, int y, int dcv, TF T, LF L)
{
    switch (mode) {
    case 0: return T(x);
    case 1: return L(y);
    case 2: return dcv;
    case 3:
        if (x == n - 1 && y == n - 1) return (T(x + y) + 3 * T(x + y + 1) + 2) >> 2;
        return (T(x + y) + 2 * T(x + y + 1) + T(x + y + 2) + 2) >> 2;
    case 4:
        if (x > y) return (T(x - y - 2) + 2 * T(x - y - 1) + T(x - y) + 2) >> 2;
        if (x < y) return (L(y - x - 2) + 2 * L(y - x - 1) + L(y - x) + 2) >> 2;
        return (T(0) + 2 * T(-1) + L(0) + 2) >> 2;
    case 5: {
        const int z = 2 * x - y;
        if (z >= 0 && (z & 1) == 0) return (T(x - (y >> 1) - 1) + T(x - (y >> 1)) + 1) >> 1;
        if (z >= 0) return (T(x - (y >> 1) - 2) + 2 * T(x - (y >> 1) - 1) + T(x - (y >> 1)) + 2) >> 2;
        if (z == -1) return (L(0) + 2 * T(-1) + T(0) + 2) >> 2;
        return (L(y - 2 * x - 1) + 2 * L(y - 2 * x - 2) + L(y - 2 * x - 3) + 2) >> 2; }
    case 6: {
        const int z = 2 * y - x;
        if (z >= 0 && (z & 1) == 0) return (L(y - (x >> 1) - 1) + L(y - (x >> 1)) + 1) >> 1;
        if (z >= 0) return (L(y - (x >> 1) - 2) + 2 * L(y - (x >> 1) - 1) + L(y - (x >> 1)) + 2) >> 2;
        if (z == -1) return (L(0) + 2 * T(-1) + T(0) + 2) >> 2;
        return (T(x - 2 * y - 1) + 2 * T(x - 2 * y - 2) + T(x - 2 * y - 3) + 2) >> 2; }
    case 7:
        if ((y & 1) == 0) return (T(x + (y >> 1)) + T(x + (y >> 1) + 1) + 1) >> 1;
        return (T(x + (y >> 1)) + 2 * T(x + (y >> 1) + 1) + T(x + (y >> 1) + 2) + 2) >> 2;
    default: {
        const int z = x + 2 * y, m = 2 * n - 3;
        if (z < m && (z & 1) == 0) return (L(y + (x >> 1)) + L(y + (x >> 1) + 1) + 1) >> 1;
        if (z < m) return (L(y + (x >> 1)) + 2 * L(y + (x >> 1) + 1) + L(y + (x >> 1) + 2) + 2) >> 2;
        if (z == m) return (L(n - 2) + 3 * L(n - 1) + 2) >> 2;
        return L(n - 1); }
    }
}

// Intra16x16 / chroma whole-plane predictors (intra_prediction.cc:668-735, 798-894) at sample (x, y).
// mode numbering here: 0 V, 1 H, 2 DC, 3 plane.  T/L as above, n = 16 or 8.
template <typename TF, typename LF>
__device__ __forceinline__ void plane_params(int n, bool chroma, TF T, LF L, int& a, int& b, int& c)
{
    const int hn = n >> 1;
    int Hs = 0, Vs = 0;
    for (int i = 0; i < hn; ++i) {
        Hs += (i + 1) * (T(hn + i) - T(hn - 2 - i));
        Vs += (i + 1) * (L(hn + i) - L(hn - 2 - i));
    }
    a = 16 * (L(n - 1) + T(n - 1));
    b = chroma ? (34 * Hs + 32) >> 6 : (5 * Hs + 32) >> 6;
    c = chroma ? (34 * Vs + 32) >> 6 : (5 * Vs + 32) >> 6;
}

template <typename TF, typename LF>
__device__ __forceinline__ int dc_value(int n, int log2n, bool a, bool b, TF T, LF L)
{
    if (!a && !b) return 128;
    int sum = 0;
    if (a) for (int y = 0; y < n; ++y) sum += L(y);
    if (b) for (int x = 0; x < n; ++x) sum += T(x);
    const int shift = log2n - 1 + (a ? 1 : 0) + (b ? 1 : 0);
    const int round = (a ? n >> 1 : 0) + (b ? n >> 1 : 0);
    return (sum + round) >> shift;
}

// Intra4x4 directional predictors as data: for mode m and sample position pos = y * 4 + x,
//     pred = (S[a] + 2 * S[b] + S[c] + 2) >> 2,
// S[o] = the tile sample at byte offset o from the block origin (row pitch 32: p(i,-1) = -32 + i, p(-1,j) = 32 j - 1,
// the corner = -33); a | b << 8 | c << 16 as signed bytes.  Two-tap averages are (a, b, a), copies (a, a, a), the
// "(x + 3y + 2) >> 2" end cases (a, b, b): the nine-way switch of intra_prediction.cc:187-356 becomes one table read.
// Generated from pred_dir_sample() (the I8x8 path still evaluates it directly); mode 2 (DC) is computed, not tabulated.
__device__ const uint32_t c_i4_pred[9 * 16] = {
    0xE0E0E0, 0xE1E1E1, 0xE2E2E2, 0xE3E3E3, 0xE0E0E0, 0xE1E1E1, 0xE2E2E2, 0xE3E3E3, 0xE0E0E0, 0xE1E1E1, 0xE2E2E2, 0xE3E3E3, 0xE0E0E0, 0xE1E1E1, 0xE2E2E2, 0xE3E3E3,
    0xFFFFFF, 0xFFFFFF, 0xFFFFFF, 0xFFFFFF, 0x1F1F1F, 0x1F1F1F, 0x1F1F1F, 0x1F1F1F, 0x3F3F3F, 0x3F3F3F, 0x3F3F3F, 0x3F3F3F, 0x5F5F5F, 0x5F5F5F, 0x5F5F5F, 0x5F5F5F,
    0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000,
    0xE2E1E0, 0xE3E2E1, 0xE4E3E2, 0xE5E4E3, 0xE3E2E1, 0xE4E3E2, 0xE5E4E3, 0xE6E5E4, 0xE4E3E2, 0xE5E4E3, 0xE6E5E4, 0xE7E6E5, 0xE5E4E3, 0xE6E5E4, 0xE7E6E5, 0xE7E7E6,
    0xFFDFE0, 0xE1E0DF, 0xE2E1E0, 0xE3E2E1, 0x1FFFDF, 0xFFDFE0, 0xE1E0DF, 0xE2E1E0, 0x3F1FFF, 0x1FFFDF, 0xFFDFE0, 0xE1E0DF, 0x5F3F1F, 0x3F1FFF, 0x1FFFDF, 0xFFDFE0,
    0xDFE0DF, 0xE0E1E0, 0xE1E2E1, 0xE2E3E2, 0xE0DFFF, 0xE1E0DF, 0xE2E1E0, 0xE3E2E1, 0xDFFF1F, 0xDFE0DF, 0xE0E1E0, 0xE1E2E1, 0xFF1F3F, 0xE0DFFF, 0xE1E0DF, 0xE2E1E0,
    0xDFFFDF, 0xE0DFFF, 0xDFE0E1, 0xE0E1E2, 0xFF1FFF, 0x1FFFDF, 0xDFFFDF, 0xE0DFFF, 0x1F3F1F, 0x3F1FFF, 0xFF1FFF, 0x1FFFDF, 0x3F5F3F, 0x5F3F1F, 0x1F3F1F, 0x3F1FFF,
    0xE0E1E0, 0xE1E2E1, 0xE2E3E2, 0xE3E4E3, 0xE2E1E0, 0xE3E2E1, 0xE4E3E2, 0xE5E4E3, 0xE1E2E1, 0xE2E3E2, 0xE3E4E3, 0xE4E5E4, 0xE3E2E1, 0xE4E3E2, 0xE5E4E3, 0xE6E5E4,
    0xFF1FFF, 0x3F1FFF, 0x1F3F1F, 0x5F3F1F, 0x1F3F1F, 0x5F3F1F, 0x3F5F3F, 0x5F5F3F, 0x3F5F3F, 0x5F5F3F, 0x5F5F5F, 0x5F5F5F, 0x5F5F5F, 0x5F5F5F, 0x5F5F5F, 0x5F5F5F,
};

// Intra8x8 directional predictors as data (scripts/gen_intra_tables.py): for mode m and sample (x, y),
//     pred = (F[a] + 2 * F[b] + F[c] + 2) >> 2,   F = IntraSmem::f8, entry = a | b << 8 | c << 16,
// which replaces the nine-way switch of intra_prediction.cc:449-621 evaluated twice per lane (the switch was 20 KB of
// code in a kernel whose warps run through it once per MB: it did not fit the 32 KB instruction cache).
__device__ const uint32_t c_i8_pred[9 * 64] = {
    0x0C0C0C, 0x0D0D0D, 0x0E0E0E, 0x0F0F0F, 0x101010, 0x111111, 0x121212, 0x131313, 0x0C0C0C, 0x0D0D0D, 0x0E0E0E, 0x0F0F0F, 0x101010, 0x111111, 0x121212, 0x131313,
    0x0C0C0C, 0x0D0D0D, 0x0E0E0E, 0x0F0F0F, 0x101010, 0x111111, 0x121212, 0x131313, 0x0C0C0C, 0x0D0D0D, 0x0E0E0E, 0x0F0F0F, 0x101010, 0x111111, 0x121212, 0x131313,
    0x0C0C0C, 0x0D0D0D, 0x0E0E0E, 0x0F0F0F, 0x101010, 0x111111, 0x121212, 0x131313, 0x0C0C0C, 0x0D0D0D, 0x0E0E0E, 0x0F0F0F, 0x101010, 0x111111, 0x121212, 0x131313,
    0x0C0C0C, 0x0D0D0D, 0x0E0E0E, 0x0F0F0F, 0x101010, 0x111111, 0x121212, 0x131313, 0x0C0C0C, 0x0D0D0D, 0x0E0E0E, 0x0F0F0F, 0x101010, 0x111111, 0x121212, 0x131313,
    0x070707, 0x070707, 0x070707, 0x070707, 0x070707, 0x070707, 0x070707, 0x070707, 0x060606, 0x060606, 0x060606, 0x060606, 0x060606, 0x060606, 0x060606, 0x060606,
    0x050505, 0x050505, 0x050505, 0x050505, 0x050505, 0x050505, 0x050505, 0x050505, 0x040404, 0x040404, 0x040404, 0x040404, 0x040404, 0x040404, 0x040404, 0x040404,
    0x030303, 0x030303, 0x030303, 0x030303, 0x030303, 0x030303, 0x030303, 0x030303, 0x020202, 0x020202, 0x020202, 0x020202, 0x020202, 0x020202, 0x020202, 0x020202,
    0x010101, 0x010101, 0x010101, 0x010101, 0x010101, 0x010101, 0x010101, 0x010101, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000,
    0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000,
    0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000,
    0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000,
    0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000,
    0x0E0D0C, 0x0F0E0D, 0x100F0E, 0x11100F, 0x121110, 0x131211, 0x141312, 0x151413, 0x0F0E0D, 0x100F0E, 0x11100F, 0x121110, 0x131211, 0x141312, 0x151413, 0x161514,
    0x100F0E, 0x11100F, 0x121110, 0x131211, 0x141312, 0x151413, 0x161514, 0x171615, 0x11100F, 0x121110, 0x131211, 0x141312, 0x151413, 0x161514, 0x171615, 0x181716,
    0x121110, 0x131211, 0x141312, 0x151413, 0x161514, 0x171615, 0x181716, 0x191817, 0x131211, 0x141312, 0x151413, 0x161514, 0x171615, 0x181716, 0x191817, 0x1A1918,
    0x141312, 0x151413, 0x161514, 0x171615, 0x181716, 0x191817, 0x1A1918, 0x1B1A19, 0x151413, 0x161514, 0x171615, 0x181716, 0x191817, 0x1A1918, 0x1B1A19, 0x1B1B1A,
    0x07080C, 0x0D0C08, 0x0E0D0C, 0x0F0E0D, 0x100F0E, 0x11100F, 0x121110, 0x131211, 0x060708, 0x07080C, 0x0D0C08, 0x0E0D0C, 0x0F0E0D, 0x100F0E, 0x11100F, 0x121110,
    0x050607, 0x060708, 0x07080C, 0x0D0C08, 0x0E0D0C, 0x0F0E0D, 0x100F0E, 0x11100F, 0x040506, 0x050607, 0x060708, 0x07080C, 0x0D0C08, 0x0E0D0C, 0x0F0E0D, 0x100F0E,
    0x030405, 0x040506, 0x050607, 0x060708, 0x07080C, 0x0D0C08, 0x0E0D0C, 0x0F0E0D, 0x020304, 0x030405, 0x040506, 0x050607, 0x060708, 0x07080C, 0x0D0C08, 0x0E0D0C,
    0x010203, 0x020304, 0x030405, 0x040506, 0x050607, 0x060708, 0x07080C, 0x0D0C08, 0x000102, 0x010203, 0x020304, 0x030405, 0x040506, 0x050607, 0x060708, 0x07080C,
    0x080C08, 0x0C0D0C, 0x0D0E0D, 0x0E0F0E, 0x0F100F, 0x101110, 0x111211, 0x121312, 0x0C0807, 0x0D0C08, 0x0E0D0C, 0x0F0E0D, 0x100F0E, 0x11100F, 0x121110, 0x131211,
    0x080706, 0x080C08, 0x0C0D0C, 0x0D0E0D, 0x0E0F0E, 0x0F100F, 0x101110, 0x111211, 0x070605, 0x0C0807, 0x0D0C08, 0x0E0D0C, 0x0F0E0D, 0x100F0E, 0x11100F, 0x121110,
    0x060504, 0x080706, 0x080C08, 0x0C0D0C, 0x0D0E0D, 0x0E0F0E, 0x0F100F, 0x101110, 0x050403, 0x070605, 0x0C0807, 0x0D0C08, 0x0E0D0C, 0x0F0E0D, 0x100F0E, 0x11100F,
    0x040302, 0x060504, 0x080706, 0x080C08, 0x0C0D0C, 0x0D0E0D, 0x0E0F0E, 0x0F100F, 0x030201, 0x050403, 0x070605, 0x0C0807, 0x0D0C08, 0x0E0D0C, 0x0F0E0D, 0x100F0E,
    0x080708, 0x0C0807, 0x080C0D, 0x0C0D0E, 0x0D0E0F, 0x0E0F10, 0x0F1011, 0x101112, 0x070607, 0x060708, 0x080708, 0x0C0807, 0x080C0D, 0x0C0D0E, 0x0D0E0F, 0x0E0F10,
    0x060506, 0x050607, 0x070607, 0x060708, 0x080708, 0x0C0807, 0x080C0D, 0x0C0D0E, 0x050405, 0x040506, 0x060506, 0x050607, 0x070607, 0x060708, 0x080708, 0x0C0807,
    0x040304, 0x030405, 0x050405, 0x040506, 0x060506, 0x050607, 0x070607, 0x060708, 0x030203, 0x020304, 0x040304, 0x030405, 0x050405, 0x040506, 0x060506, 0x050607,
    0x020102, 0x010203, 0x030203, 0x020304, 0x040304, 0x030405, 0x050405, 0x040506, 0x010001, 0x000102, 0x020102, 0x010203, 0x030203, 0x020304, 0x040304, 0x030405,
    0x0C0D0C, 0x0D0E0D, 0x0E0F0E, 0x0F100F, 0x101110, 0x111211, 0x121312, 0x131413, 0x0E0D0C, 0x0F0E0D, 0x100F0E, 0x11100F, 0x121110, 0x131211, 0x141312, 0x151413,
    0x0D0E0D, 0x0E0F0E, 0x0F100F, 0x101110, 0x111211, 0x121312, 0x131413, 0x141514, 0x0F0E0D, 0x100F0E, 0x11100F, 0x121110, 0x131211, 0x141312, 0x151413, 0x161514,
    0x0E0F0E, 0x0F100F, 0x101110, 0x111211, 0x121312, 0x131413, 0x141514, 0x151615, 0x100F0E, 0x11100F, 0x121110, 0x131211, 0x141312, 0x151413, 0x161514, 0x171615,
    0x0F100F, 0x101110, 0x111211, 0x121312, 0x131413, 0x141514, 0x151615, 0x161716, 0x11100F, 0x121110, 0x131211, 0x141312, 0x151413, 0x161514, 0x171615, 0x181716,
    0x070607, 0x050607, 0x060506, 0x040506, 0x050405, 0x030405, 0x040304, 0x020304, 0x060506, 0x040506, 0x050405, 0x030405, 0x040304, 0x020304, 0x030203, 0x010203,
    0x050405, 0x030405, 0x040304, 0x020304, 0x030203, 0x010203, 0x020102, 0x000102, 0x040304, 0x020304, 0x030203, 0x010203, 0x020102, 0x000102, 0x010001, 0x000001,
    0x030203, 0x010203, 0x020102, 0x000102, 0x010001, 0x000001, 0x000000, 0x000000, 0x020102, 0x000102, 0x010001, 0x000001, 0x000000, 0x000000, 0x000000, 0x000000,
    0x010001, 0x000001, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000, 0x000000,
};

// Everything about an intra MB that does not depend on its neighbours being reconstructed: header, the header words
// of the four neighbouring MBs (lane & 3 = 0 left, 1 top, 2 top-left, 3 top-right), the residual, the slice's
// constrained_intra_pred_flag.  Loaded ahead of time (next MB of the row / before the dependency wait).
struct IntraPre {
    MbHdr h;
    uint32_t nbw;                   // header word 0 of neighbour (lane & 3), 0xFFFFFFFF outside the picture
    uint4 r0, r1;                   // residual chunks lane and 32 + lane (lanes 0..15) of the MB's 48 x 16 bytes
    int ci;
};
__device__ __forceinline__ void intra_prefetch(const DevPicture& pic, const FrameGeom& g, int mbx, int mby, int lane, IntraPre& p)
{
    const int W = g.width_mbs, addr = mby * W + mbx;
    p.h = load_hdr(pic.mbs, addr);
    const int k = lane & 3;
    const int nx = mbx + (k == 3 ? 1 : (k == 1 ? 0 : -1)), ny = mby - (k == 0 ? 0 : 1);
    p.nbw = 0xFFFFFFFFu;
    if (nx >= 0 && nx < W && ny >= 0) p.nbw = load_hdr_word0(pic.mbs, ny * W + nx);
    const uint4* rsrc = reinterpret_cast<const uint4*>(pic.resid + (size_t)addr * H264R_COEFFS_PER_MB);
    p.r0 = __ldg(rsrc + lane);
    p.r1 = lane < 16 ? __ldg(rsrc + 32 + lane) : make_uint4(0, 0, 0, 0);
    p.ci = (int)__ldg(&(pic.slices + p.h.slice_idx)->constrained_intra_pred_flag);
}

// Reconstruction of one intra macroblock by one warp (mb_pred_intra / mb_pred_ipcm, decoder.cc:149-215): neighbour
// availability, neighbour samples of the current unfiltered picture, prediction + residual block by block through a
// shared-memory tile, store.  The caller has made sure that the neighbouring MBs are reconstructed and visible.
// kRowMode (row wavefront): the caller has put the samples above the MB (from the mailboxes of the row above) and the
// left column (carried in the tile from the previous MB of the row) into the tiles; otherwise they are read from the frame.
template <bool kRowMode>
__device__ __forceinline__ void intra_reconstruct_mb(const DevPicture& pic, const FrameGeom& g, IntraSmem& sm, const IntraPre& pre,
                                                     int mbx, int mby, int lane)
{
    const MbHdr& h = pre.h;
    const int W = g.width_mbs;
    uint8_t* const dY = pic.dst;
    uint8_t* const dC[2] = { pic.dst + g.off_cb, pic.dst + g.off_cr };
    const int px = mbx * 16, py = mby * 16, cx = mbx * 8, cy = mby * 8;

    if (h.mb_type == H264R_MB_IPCM) {                 // mb_pred_ipcm, decoder.cc:149-168
        const h264r_level* __restrict__ lv = pic.levels + h.coeff_offset;
        for (int i = lane; i < h.coeff_count; i += 32) {
            const uint32_t e = __ldg(lv + i);
            const int p = (int)(e & 0xFFFFu), v = (int)(e >> 16) & 0xFF;
            if (p < 256) {
                dY[(size_t)(py + (p >> 4)) * g.pitch_y + px + (p & 15)] = (uint8_t)v;
                if (kRowMode) TY(p & 15, p >> 4) = (uint8_t)v;       // the row wavefront carries the MB in its tile
            } else if (p < 384) {
                const int pl = (p - 256) >> 6, q = (p - 256) & 63;
                dC[pl][(size_t)(cy + (q >> 3)) * g.pitch_c + cx + (q & 7)] = (uint8_t)v;
                if (kRowMode) TC(pl, q & 7, q >> 3) = (uint8_t)v;
            }
        }
        if (kRowMode) __syncwarp();
        return;
    }

    // availability of the four neighbouring MBs (neighbour.cc:123-175 + slice_nr + constrained intra): the neighbour
    // exists, belongs to the same slice and -- with constrained_intra_pred -- is an intra MB.  All four precede the MB
    // in raster order.
    const uint32_t w0 = (uint32_t)h.mb_type | (uint32_t)h.flags << 8 | (uint32_t)h.slice_idx << 16;
    bool av[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t nw = __shfl_sync(0xFFFFFFFFu, pre.nbw, k);
        av[k] = nw != 0xFFFFFFFFu && (nw >> 16) == (w0 >> 16) && (!pre.ci || ((nw >> 8) & H264R_MB_FLAG_INTRA));
    }
    const bool aL = av[0], aT = av[1], aTL = av[2], aTR = av[3];

    // neighbour samples of the current, unfiltered picture -> tiles (L1-bypassing loads: other SMs wrote them)
    if (!kRowMode && mby > 0) {
        if (lane < 8) {                                // luma top row, cols -4..27
            const int x = px - 4 + lane * 4;
            uint32_t v = 0;
            if (x >= 0 && x < W * 16) v = ldcg_u32(dY + (size_t)(py - 1) * g.pitch_y + x);
            reinterpret_cast<uint32_t*>(sm.ty)[lane] = v;
        } else if (lane < 16) {                        // chroma top rows, cols -4..11
            const int c = lane - 8, pl = c >> 2, x = cx - 4 + (c & 3) * 4;
            uint32_t v = 0;
            if (x >= 0 && x < W * 8) v = ldcg_u32(dC[pl] + (size_t)(cy - 1) * g.pitch_c + x);
            reinterpret_cast<uint32_t*>(sm.tc[pl])[c & 3] = v;
        }
    }
    if (!kRowMode && mbx > 0) {
        if (lane < 16) TY(-1, lane) = ldcg_u8(dY + (size_t)(py + lane) * g.pitch_y + px - 1);
        else { const int c = lane - 16, pl = c >> 3, y = c & 7; TC(pl, -1, y) = ldcg_u8(dC[pl] + (size_t)(cy + y) * g.pitch_c + cx - 1); }
    }
    {   // residual plane written by residual_kernel (48 x 16 B), or zeros
        const bool has = h.has_resid();
        reinterpret_cast<uint4*>(sm.res)[lane] = has ? pre.r0 : make_uint4(0, 0, 0, 0);
        if (lane < 16) reinterpret_cast<uint4*>(sm.res)[32 + lane] = has ? pre.r1 : make_uint4(0, 0, 0, 0);
    }
    __syncwarp();                                      // tiles and residual visible

    // ---- luma ----
    if (h.mb_type == H264R_MB_I16x16) {
        auto T = [&](int i) { return (int)TY(i, -1); };
        auto L = [&](int i) { return (int)TY(-1, i); };
        const int y = lane >> 1, x0 = (lane & 1) * 8;
        int pa = 0, pb = 0, pc = 0, dcv = 0;
        if (h.i16mode == 3) plane_params(16, false, T, L, pa, pb, pc);
        else if (h.i16mode == 2) dcv = dc_value(16, 4, aL, aT, T, L);
        int v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int x = x0 + i;
            int p;
            if (h.i16mode == 0) p = T(x);
            else if (h.i16mode == 1) p = L(y);
            else if (h.i16mode == 2) p = dcv;
            else p = clip255((pa + pb * (x - 7) + pc * (y - 7) + 16) >> 5);
            v[i] = clip255(p + sm.res[y * 16 + x]);
        }
        __syncwarp();                                  // all lanes have read the border before the tile is written
#pragma unroll
        for (int i = 0; i < 8; ++i) TY(x0 + i, y) = (uint8_t)v[i];
    } else if (h.mb_type != H264R_MB_I8x8) {
        // I_4x4: sixteen blocks in coding order, each waiting for the previous one through the tile; lane = sample
        const int x = lane & 3, y = (lane >> 2) & 3;
#pragma unroll 1
        for (int k = 0; k < 16; ++k) {
            const int xO = ((k >> 2) & 1) * 8 + (k & 1) * 4, yO = (k >> 3) * 8 + ((k >> 1) & 1) * 4;
            const int mode = ((k < 8 ? h.u0 >> (4 * k) : h.u1 >> (4 * (k - 8)))) & 15;
            const bool avA = xO > 0 ? true : aL;
            const bool avB = yO > 0 ? true : aT;
            bool avC;
            if (yO == 0) avC = (xO + 4 < 16) ? aT : aTR;
            else avC = xO + 4 < 16;
            if (xO == 4 && (yO == 4 || yO == 12)) avC = false;
            const uint8_t* const blk = &TY(xO, yO);
            const int r = sm.res[(yO + y) * 16 + xO + x];
            int pv;
            if (mode == 2) {                           // DC (intra_prediction.cc:206-232)
                const int top = __dp4a(*reinterpret_cast<const uint32_t*>(blk - 32), 0x01010101u, 0u);
                const int left = (int)blk[-1] + blk[31] + blk[63] + blk[95];
                pv = avA && avB ? (top + left + 4) >> 3 : (avA ? (left + 2) >> 2 : (avB ? (top + 2) >> 2 : 128));
            } else {
                const uint32_t e = __ldg(&c_i4_pred[min(mode, 8) * 16 + (lane & 15)]);
                int oa = (int)(int8_t)(e & 0xFF), ob = (int)(int8_t)((e >> 8) & 0xFF), oc = (int)(int8_t)((e >> 16) & 0xFF);
                if (!avC) {                            // p(x,-1), x = 4..7 -> p(3,-1) (intra_prediction.cc:182-185)
                    if (oa > -29 && oa < -1) oa = -29;
                    if (ob > -29 && ob < -1) ob = -29;
                    if (oc > -29 && oc < -1) oc = -29;
                }
                pv = ((int)blk[oa] + 2 * (int)blk[ob] + (int)blk[oc] + 2) >> 2;
            }
            const int v = clip255(pv + r);
            if (lane < 16) TY(xO + x, yO + y) = (uint8_t)v;      // the block never reads its own samples: no barrier before
            __syncwarp();
        }
    } else {
        // I_8x8: four blocks in coding order; lane = samples (2 (lane & 3), lane >> 2) and the one to its right
        const int y = lane >> 2, x0 = (lane & 3) * 2;
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            const int xO = (k & 1) * 8, yO = (k >> 1) * 8;
            const int mode = (h.u0 >> (4 * k)) & 15;
            const bool avA = xO > 0 ? true : aL;
            const bool avB = yO > 0 ? true : aT;
            const bool avD = (xO > 0 && yO > 0) ? true : (xO > 0 ? aT : (yO > 0 ? aL : aTL));
            const bool avC = k == 0 ? aT : (k == 1 ? aTR : k == 2);          // block 3 never has a top-right neighbour (:376)
            const int tmax = avC ? 15 : 7;             // C substitution: p(x,-1) = p(7,-1) for x >= 8 (:404-407)

            // reference sample filtering (Intra8x8::filtering, intra_prediction.cc:413-447)
            auto To = [&](int i) { return (int)TY(xO + min(i, tmax), yO - 1); };
            auto Lo = [&](int i) { return (int)TY(xO - 1, yO + i); };
            if (lane < 16) {                           // p'(lane, -1)
                int f = 0;
                if (avB) {
                    if (lane == 0) f = avD ? (To(-1) + 2 * To(0) + To(1) + 2) >> 2 : (3 * To(0) + To(1) + 2) >> 2;
                    else if (lane == 15) f = (To(14) + 3 * To(15) + 2) >> 2;
                    else f = (To(lane - 1) + 2 * To(lane) + To(lane + 1) + 2) >> 2;
                }
                sm.f8[12 + lane] = (uint8_t)f;
            } else if (lane < 24) {                    // p'(-1, i)
                const int i = lane - 16;
                int f = 0;
                if (avA) {
                    if (i == 0) f = avD ? (Lo(-1) + 2 * Lo(0) + Lo(1) + 2) >> 2 : (3 * Lo(0) + Lo(1) + 2) >> 2;
                    else if (i == 7) f = (Lo(6) + 3 * Lo(7) + 2) >> 2;
                    else f = (Lo(i - 1) + 2 * Lo(i) + Lo(i + 1) + 2) >> 2;
                }
                sm.f8[7 - i] = (uint8_t)f;
            } else if (lane == 24) {                   // p'(-1, -1)
                int f = 0;
                if (avD) {
                    const int c = To(-1);
                    if (avA && avB) f = (To(0) + 2 * c + Lo(0) + 2) >> 2;
                    else if (avB) f = (3 * c + To(0) + 2) >> 2;
                    else if (avA) f = (3 * c + Lo(0) + 2) >> 2;
                    else f = c;
                }
                sm.f8[8] = (uint8_t)f;
            }
            __syncwarp();
            int p0, p1;
            if (mode == 2) {                           // DC (intra_prediction.cc:466-492)
                const uint32_t* fw = reinterpret_cast<const uint32_t*>(sm.f8);
                const int left = __dp4a(fw[0], 0x01010101u, __dp4a(fw[1], 0x01010101u, 0u));
                const int top = __dp4a(fw[3], 0x01010101u, __dp4a(fw[4], 0x01010101u, 0u));
                p0 = p1 = avA && avB ? (left + top + 8) >> 4 : (avA ? (left + 4) >> 3 : (avB ? (top + 4) >> 3 : 128));
            } else {
                const uint2 e = __ldg(reinterpret_cast<const uint2*>(&c_i8_pred[min(mode, 8) * 64 + y * 8 + x0]));
                p0 = ((int)sm.f8[e.x & 0xFF] + 2 * (int)sm.f8[(e.x >> 8) & 0xFF] + (int)sm.f8[(e.x >> 16) & 0xFF] + 2) >> 2;
                p1 = ((int)sm.f8[e.y & 0xFF] + 2 * (int)sm.f8[(e.y >> 8) & 0xFF] + (int)sm.f8[(e.y >> 16) & 0xFF] + 2) >> 2;
            }
            const uint32_t r2 = *reinterpret_cast<const uint32_t*>(&sm.res[(yO + y) * 16 + xO + x0]);
            const int v0 = clip255(p0 + (int)(int16_t)(r2 & 0xFFFF)), v1 = clip255(p1 + (int)(int16_t)(r2 >> 16));
            *reinterpret_cast<uint16_t*>(&TY(xO + x0, yO + y)) = (uint16_t)(v0 | v1 << 8);    // the block never reads its own samples
            __syncwarp();
        }
    }

    // ---- chroma: lanes 0..15 Cb, 16..31 Cr; 4 samples per lane ----
    {
        const int pl = lane >> 4, l16 = lane & 15, y = l16 >> 1, x0 = (l16 & 1) * 4;
        auto T = [&](int i) { return (int)TC(pl, i, -1); };
        auto L = [&](int i) { return (int)TC(pl, -1, i); };
        const int m = h.cmode;                         // 0 DC, 1 H, 2 V, 3 plane
        int pa = 0, pb = 0, pc = 0, dcv = 0;
        if (m == 3) plane_params(8, true, T, L, pa, pb, pc);
        else if (m == 0) {                             // DC of this lane's 4x4 block (intra_prediction.cc:825-849)
            const int xO = x0, yO = y & 4;
            bool a, b;
            if ((xO == 0 && yO == 0) || (xO > 0 && yO > 0)) { a = aL; b = aT; }
            else if (xO > 0) { a = aT ? false : aL; b = aT; }
            else { a = aL; b = aL ? false : aT; }
            auto T4 = [&](int i) { return T(xO + i); };
            auto L4 = [&](int i) { return L(yO + i); };
            dcv = dc_value(4, 2, a, b, T4, L4);
        }
        int v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = x0 + i;
            int p;
            if (m == 0) p = dcv;
            else if (m == 1) p = L(y);
            else if (m == 2) p = T(x);
            else p = clip255((pa + pb * (x - 3) + pc * (y - 3) + 16) >> 5);
            v[i] = clip255(p + sm.res[256 + pl * 64 + y * 8 + x]);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) TC(pl, x0 + i, y) = (uint8_t)v[i];
    }
    __syncwarp();

    // ---- store the reconstructed MB ----
    if (lane < 16) {
        const uint32_t* r = reinterpret_cast<const uint32_t*>(&TY(0, lane));
        *reinterpret_cast<uint4*>(dY + (size_t)(py + lane) * g.pitch_y + px) = make_uint4(r[0], r[1], r[2], r[3]);
    } else {
        const int c = lane - 16, pl = c >> 3, y = c & 7;
        const uint32_t* r = reinterpret_cast<const uint32_t*>(&TC(pl, 0, y));
        *reinterpret_cast<uint2*>(dC[pl] + (size_t)(cy + y) * g.pitch_c + cx) = make_uint2(r[0], r[1]);
    }
}

// Mailbox of an intra MB: its bottom sample rows, 8 words = luma row 15 (4 words) | Cb row 7 (2) | Cr row 7 (2).
constexpr int kIntraBoxWords = 8;
constexpr uint32_t kIntraEpochTag = 0x80000000u;

// All-intra pictures: one warp per MB row, rows form the 2:1 wavefront.  MB (x, y) needs, from the row above, the bottom
// row of MB x, the first eight bottom samples of MB x+1 and the last bottom sample of MB x-1: 24 consecutive mailbox
// words, polled by 24 lanes with one load each.  The left column never leaves the tile.  No fence, no progress counter.
__global__ void __launch_bounds__(kWarpsPerCta * 32, H264R_INTRA_CTAS)
recon_intra_kernel(const DevPicture* __restrict__ pics, int num_pics, int* tickets, FrameGeom g, uint32_t epoch)
{
    __shared__ __align__(16) IntraSmem smem_all[kWarpsPerCta];
    __shared__ int s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(&tickets[0], 1);
    __syncthreads();
    const int W = g.width_mbs, H = g.height_mbs;
    const int groups = (H + kWarpsPerCta - 1) / kWarpsPerCta;
    // tickets run row-group-major over the pictures of the wave: a CTA's predecessor (same picture, previous
    // row group) took its ticket num_pics tickets earlier, so it is normally far ahead and nobody spins
    const int rg = s_ticket / num_pics, pic_i = s_ticket - rg * num_pics;
    if (rg >= groups) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mby = rg * kWarpsPerCta + warp;
    if (mby >= H) return;
    const DevPicture& pic = pics[pic_i];
    if (!pic.has_intra || pic.has_inter) return;         // mixed pictures: recon_intra_sparse_kernel
    IntraSmem& sm = smem_all[warp];
    const uint32_t tag = epoch | kIntraEpochTag;
    uint64_t* const box_out = pic.mbox + (size_t)mby * W * kIntraBoxWords;
    const uint64_t* const box_in = pic.mbox + (size_t)(mby > 0 ? mby - 1 : 0) * W * kIntraBoxWords;
    const bool has_below = mby + 1 < H;

    IntraPre nxt;
    intra_prefetch(pic, g, 0, mby, lane, nxt);
    for (int mbx = 0; mbx < W; ++mbx) {
        const IntraPre cur = nxt;
        // the 24 words around MB mbx of the row above: lane j = word j & 7 of MB mbx - 1 + (j >> 3)
        uint64_t t = 0;
        const int bx = mbx - 1 + (lane >> 3), bw = lane & 7;
        // needed: the corner samples of MB mbx-1 (words 3, 5, 7), everything of MB mbx, the first two luma words of MB mbx+1
        const bool need = mby > 0 && lane < 24 && bx >= 0 && bx < W &&
                          (lane < 8 ? (bw == 3 || bw == 5 || bw == 7) : (lane < 16 ? true : bw < 2));
        if (need) t = ld_mbox(box_in + (size_t)bx * kIntraBoxWords + bw);
        if (mbx + 1 < W) intra_prefetch(pic, g, mbx + 1, mby, lane, nxt);      // lands while this MB is reconstructed
        __syncwarp();                                      // the previous MB's tile has been stored and posted
        if (mbx > 0) {                                     // left column = the previous MB's last column, still in the tile
            if (lane < 16) TY(-1, lane) = TY(15, lane);
            else { const int c = lane - 16, pl = c >> 3, y = c & 7; TC(pl, -1, y) = TC(pl, 7, y); }
        }
        if (mby > 0) {
            bool waiting = need && (uint32_t)(t >> 32) != tag;
            unsigned ns = 16;
            while (__any_sync(0xFFFFFFFFu, waiting)) {
                if (waiting) {
                    __nanosleep(ns); if (ns < 128) ns *= 2;
                    t = ld_mbox(box_in + (size_t)bx * kIntraBoxWords + bw);
                    waiting = (uint32_t)(t >> 32) != tag;
                }
            }
            // tile row -1: luma words 0..7 = columns -4..27, chroma words 0..3 = columns -4..11 (0 outside the picture)
            const uint32_t v = need ? (uint32_t)t : 0u;
            if (lane < 8) {
                if (bw == 3) reinterpret_cast<uint32_t*>(sm.ty)[0] = v;
                else if (bw == 5) reinterpret_cast<uint32_t*>(sm.tc[0])[0] = v;
                else if (bw == 7) reinterpret_cast<uint32_t*>(sm.tc[1])[0] = v;
            } else if (lane < 16) {
                if (bw < 4) reinterpret_cast<uint32_t*>(sm.ty)[1 + bw] = v;
                else reinterpret_cast<uint32_t*>(sm.tc[(bw - 4) >> 1])[1 + (bw & 1)] = v;
            } else if (lane < 24 && bw < 2) reinterpret_cast<uint32_t*>(sm.ty)[5 + bw] = v;
        }
        intra_reconstruct_mb<true>(pic, g, sm, cur, mbx, mby, lane);
        // post the MB's bottom rows (the tile is final: the MB's own stores read it after a __syncwarp)
        if (has_below && lane < 8) {
            const uint32_t w = lane < 4 ? reinterpret_cast<const uint32_t*>(&TY(0, 15))[lane]
                                        : reinterpret_cast<const uint32_t*>(&TC((lane - 4) >> 1, 0, 7))[lane & 1];
            st_mbox(box_out + (size_t)mbx * kIntraBoxWords + lane, w, tag);
        }
    }
}

// Intra MBs of pictures that also have inter MBs (P/B pictures: a few percent of the MBs, mostly isolated).  One warp
// per intra MB, taken in raster order from the picture's address list; the warp waits only for those of its four
// neighbours (left, top-left, top, top-right) that are intra MBs themselves -- inter neighbours were reconstructed by
// recon_inter_kernel.  Completion is an epoch stamp per MB (no clearing between launches).  Tickets interleave the
// pictures of the wave and run in raster order inside a picture, so a warp only ever waits for warps that already
// hold a ticket.
#ifndef H264R_SPARSE_PER_WARP
#define H264R_SPARSE_PER_WARP 1
#endif
// A warp takes H264R_SPARSE_PER_WARP consecutive entries of the list: header, neighbour headers and residual of the
// next MB are in flight while the current one is reconstructed (the kernel is bound by those dependent loads).
#ifndef H264R_SPARSE_WARPS
#define H264R_SPARSE_WARPS 4
#endif
#ifndef H264R_SPARSE_CTAS
#define H264R_SPARSE_CTAS (48 / H264R_SPARSE_WARPS)
#endif
constexpr int kSparseWarps = H264R_SPARSE_WARPS;
__global__ void __launch_bounds__(kSparseWarps * 32, H264R_SPARSE_CTAS)
recon_intra_sparse_kernel(const DevPicture* __restrict__ pics, int num_pics, int* tickets, FrameGeom g, uint32_t epoch)
{
    __shared__ __align__(16) IntraSmem smem_all[kSparseWarps];
    __shared__ int s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(&tickets[2], 1);
    __syncthreads();
    const int grp = s_ticket / num_pics, pic_i = s_ticket - grp * num_pics;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const DevPicture& pic = pics[pic_i];
    const int first = (grp * kSparseWarps + warp) * H264R_SPARSE_PER_WARP;
    const int n = min(H264R_SPARSE_PER_WARP, pic.intra_count - first);
    if (n <= 0) return;
    const int W = g.width_mbs;
    // the warp's addresses: lane i holds entry i
    const int my_addr = lane < n ? (int)__ldg(pic.intra_list + first + lane) : 0;
    IntraPre nxt;
    int addr = __shfl_sync(0xFFFFFFFFu, my_addr, 0);
    intra_prefetch(pic, g, addr % W, addr / W, lane, nxt);          // header, neighbour headers, residual: all in flight at once
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const IntraPre pre = nxt;
        const int mby = addr / W, mbx = addr - mby * W, cur_addr = addr;
        if (i + 1 < n) {
            addr = __shfl_sync(0xFFFFFFFFu, my_addr, i + 1);
            intra_prefetch(pic, g, addr % W, addr / W, lane, nxt);
        }
        if (lane < 4 && pre.nbw != 0xFFFFFFFFu && ((pre.nbw >> 8) & H264R_MB_FLAG_INTRA)) {
            const int nx = mbx + (lane == 3 ? 1 : (lane == 1 ? 0 : -1)), ny = mby - (lane == 0 ? 0 : 1);   // left, top, top-left, top-right
            const int* flag = reinterpret_cast<const int*>(pic.mb_done + ny * W + nx);
            unsigned ns = 16;
            while ((uint32_t)ld_acquire(flag) != epoch) { __nanosleep(ns); if (ns < 256) ns *= 2; }
        }
        __syncwarp();
        intra_reconstruct_mb<false>(pic, g, smem_all[warp], pre, mbx, mby, lane);
        __syncwarp();
        if (lane == 0) st_release(reinterpret_cast<int*>(pic.mb_done + cur_addr), (int)epoch);
    }
}

// ---------------------------------------------------------------------------------------------------
// deblocking (wavefront)

__constant__ uint8_t c_alpha[52] = {
    0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,4,4,5,6,7,8,9,10,12,13,15,17,20,22,25,28,32,36,40,45,50,56,63,71,80,90,101,113,127,144,162,182,203,226,255,255 };
__constant__ uint8_t c_beta[52] = {
    0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,2,2,3,3,3,3,4,4,4,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13,14,14,15,15,16,16,17,17,18,18 };
__constant__ uint8_t c_tc0[52][3] = {
    {0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},
    {0,0,1},{0,0,1},{0,0,1},{0,0,1},{0,1,1},{0,1,1},{1,1,1},{1,1,1},{1,1,1},{1,1,1},{1,1,2},{1,1,2},{1,1,2},{1,1,2},{1,2,3},{1,2,3},{2,2,3},
    {2,2,4},{2,3,4},{2,3,4},{3,3,5},{3,4,6},{3,4,6},{4,5,7},{4,5,8},{4,6,9},{5,7,10},{6,8,11},{6,8,13},{7,10,14},{8,11,16},{9,12,18},
    {10,13,20},{11,15,23},{13,17,25} };

// ---- pass 1 (fully parallel): per-MB deblock descriptor = boundary strengths + filter thresholds ----

// |mv_x| or |mv_y| differ by four quarter samples or more (mvlimit 4, frame pictures); a, b = packed int16 pairs
__device__ __forceinline__ int mv_differs(uint32_t a, uint32_t b)
{
    const int dx = (int)(int16_t)(a & 0xFFFF) - (int)(int16_t)(b & 0xFFFF), dy = (int)(int16_t)(a >> 16) - (int)(int16_t)(b >> 16);
    return (abs(dx) >= 4) | (abs(dy) >= 4);
}
// bs_compare_mvs, deblock.cc:35-75, on two packed motion entries (words: mv[0], mv[1], ref_idx[0..1] | ref_pic[0..1] << 16)
__device__ __forceinline__ int bs_compare(const uint32_t* ep, const uint32_t* eq)
{
    if (ep == eq) return 0;                                // the same entry: same pictures, same vectors
    const uint32_t rp = __ldg(ep + 2), rq = __ldg(eq + 2);
    const int p0 = (int8_t)(rp >> 16), p1 = (int8_t)(rp >> 24), q0 = (int8_t)(rq >> 16), q1 = (int8_t)(rq >> 24);
    if (!((p0 == q0 && p1 == q1) || (p0 == q1 && p1 == q0))) return 1;
    const uint32_t mp0 = __ldg(ep), mp1 = __ldg(ep + 1), mq0 = __ldg(eq), mq1 = __ldg(eq + 1);
    if (p0 != p1) {
        if (p0 == q0) return mv_differs(mp0, mq0) | mv_differs(mp1, mq1);
        return mv_differs(mp0, mq1) | mv_differs(mp1, mq0);
    }
    return (mv_differs(mp0, mq0) | mv_differs(mp1, mq1)) & (mv_differs(mp0, mq1) | mv_differs(mp1, mq0));
}

// Deblock::strength (deblock.cc:78-289) + the qPav/indexA/indexB/alpha/beta/tc0 part of filter_edge (deblock.cc:469-474,
// tables :294-324).  One THREAD per MB (the work is scalar: 32 strengths and 9 threshold sets out of three MB headers).
struct HdrLite { int mb_type, flags, slice_idx, qp_y, qp_c[2], cbp_blks; uint32_t packed; };
__device__ __forceinline__ HdrLite load_hdr_lite(const h264r_mb* mbs, int addr)
{
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(mbs + addr));
    HdrLite h;
    h.mb_type = a.x & 0xFF; h.flags = (a.x >> 8) & 0xFF; h.slice_idx = a.x >> 16;
    h.qp_y = (int)(int8_t)(a.y >> 16); h.qp_c[0] = (int)(int8_t)(a.y >> 24); h.qp_c[1] = (int)(int8_t)(a.z & 0xFF);
    h.cbp_blks = a.w & 0xFFFF;
    h.packed = (h.flags & H264R_MB_FLAG_INTRA) ? 0u : __ldg(reinterpret_cast<const unsigned int*>(mbs + addr) + 7);
    return h;
}

__global__ void __launch_bounds__(128, H264R_PREP_CTAS)
deblock_prep_kernel(const DevPicture* __restrict__ pics, int num_pics, FrameGeom g)
{
    const int W = g.width_mbs, nmb = W * g.height_mbs;
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= (long long)num_pics * nmb) return;
    const int pic_i = (int)(gi / nmb), q = (int)(gi - (long long)pic_i * nmb);
    const DevPicture& pic = pics[pic_i];
    if (!pic.run_deblock) return;
    const int mbx = q % W, mby = q / W;
    const HdrLite Q = load_hdr_lite(pic.mbs, q);
    const h264r_slice* sl = pic.slices + Q.slice_idx;
    const int idc = __ldg(&sl->disable_deblocking_filter_idc);
    uint4* out = reinterpret_cast<uint4*>(pic.desc + q);
    if (idc == 1) { out[0] = make_uint4(0, 0, 0, 0); return; }           // no strengths: the thresholds are never read

    bool left = mbx > 0, top = mby > 0;
    HdrLite PL = Q, PT = Q;
    if (left) { PL = load_hdr_lite(pic.mbs, q - 1); if (idc == 2 && PL.slice_idx != Q.slice_idx) left = false; }
    if (top)  { PT = load_hdr_lite(pic.mbs, q - W); if (idc == 2 && PT.slice_idx != Q.slice_idx) top = false; }
    const bool q_intra = Q.flags & H264R_MB_FLAG_INTRA, t8 = Q.flags & H264R_MB_FLAG_T8x8;
    const bool p_skip = __ldg(&sl->slice_type) == H264R_P_SLICE && Q.mb_type == 0;

    // (loops kept rolled: unrolled, the kernel was 77 KB of code for a 32 KB instruction cache)
    uint32_t bs0 = 0, bs1 = 0, bs2 = 0, bs3 = 0;
    auto bs_or = [&](int wi, uint32_t v) { if (wi == 0) bs0 |= v; else if (wi == 1) bs1 |= v; else if (wi == 2) bs2 |= v; else bs3 |= v; };
#if H264R_PREP_UNROLL < 2
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int dir = 0; dir < 2; ++dir) {
        const bool mbedge = dir == 0 ? left : top;
        const HdrLite& PN = dir == 0 ? PL : PT;
#if H264R_PREP_UNROLL < 1
#pragma unroll 1
#else
#pragma unroll
#endif
        for (int e = 0; e < 4; ++e) {
            const bool on = e == 0 ? mbedge : !(t8 && (e & 1));
            if (!on) continue;
            if (e > 0 && p_skip) continue;
            const int wi = dir * 2 + (e >> 1), sh = (e & 1) * 16;
            const bool p_intra = e ? q_intra : (PN.flags & H264R_MB_FLAG_INTRA) != 0;
            if (p_intra || q_intra) { bs_or(wi, (e == 0 ? 0x4444u : 0x3333u) << sh); continue; }
            const int pcbp = e ? Q.cbp_blks : PN.cbp_blks;
            const bool same_part = e > 0 && (Q.mb_type == 1 || Q.mb_type == (dir == 0 ? 2 : 3));
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                const int blkQ = dir == 0 ? k4 * 4 + e : e * 4 + k4;
                const int blkP = dir == 0 ? k4 * 4 + (e ? e - 1 : 3) : (e ? e - 1 : 3) * 4 + k4;
                uint32_t v = 0;
                if (((Q.cbp_blks >> blkQ) & 1) || ((pcbp >> blkP) & 1)) v = 2;
                else if (!same_part && bs_compare(packed_entry(pic.packed_motion, e ? Q.packed : PN.packed, blkP),
                                                   packed_entry(pic.packed_motion, Q.packed, blkQ))) v = 1;
                bs_or(wi, v << (sh + k4 * 4));
            }
        }
    }
    out[0] = make_uint4(bs0, bs1, bs2, bs3);
    // thresholds: type 0 = left MB edge, 1 = internal edge, 2 = top MB edge
    const int foa = (int)(int8_t)__ldg(&sl->filter_offset_a), fob = (int)(int8_t)__ldg(&sl->filter_offset_b);
    uint32_t w[12];                                       // [plane][type], contiguous: no padding words
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const HdrLite& P = t == 0 ? PL : (t == 2 ? PT : Q);
            const int qp_p = pl ? P.qp_c[pl - 1] : P.qp_y, qp_q = pl ? Q.qp_c[pl - 1] : Q.qp_y;
            const int qPav = (qp_p + qp_q + 1) >> 1;
            const int ia = clip3i(0, 51, qPav + foa), ib = clip3i(0, 51, qPav + fob);
            w[pl * 3 + t] = (uint32_t)c_alpha[ia] | (uint32_t)c_beta[ib] << 8 | (uint32_t)c_tc0[ia][0] << 13 |
                   (uint32_t)c_tc0[ia][1] << 18 | (uint32_t)c_tc0[ia][2] << 23;
        }
    }
    out[1] = make_uint4(w[0], w[1], w[2], w[3]);
    out[2] = make_uint4(w[4], w[5], w[6], w[7]);
    reinterpret_cast<uint32_t*>(out)[12] = w[8];
}

// ---- pass 2 (row wavefront): filtering, in place ----
//
// One warp filters the SAME macroblock row of TWO pictures of the wave: lanes 0..15 picture A, lanes 16..31 picture B
// (one instruction stream, independent data: the filter is a data-dependent scalar recipe per line, so a picture can
// keep only 16 lanes busy).  Lane l of a half owns luma line l and chroma line l & 7 of plane l >> 3.
//   vertical edges  : the lane's row lives in registers (left 4 samples carried from the previous MB + own 16 / 8),
//                     the four (two) edges are filtered in sequence without touching memory;
//   horizontal edges: the row goes through a shared-memory tile (transposition), the lane then owns a column.
//
// Rows talk through MAILBOXES, not through the frame (the low-latency protocol of collective libraries: data and
// flag travel in the same 64-bit word, so neither side needs a fence).  The last four luma rows and the last two
// rows of each chroma plane of MB (x, y) -- the only samples MB (x, y+1) reads or changes -- are final once the row's
// warp has filtered the left edge of MB (x+1, y).  At that point the warp posts them as 24 words of
// { 4 samples, launch epoch } (st.relaxed.gpu.u64, single-copy atomic); the warp of row y+1 polls the 24 words of
// mailbox (x, y), filters its top edge on them, and is the ONLY writer of luma rows 13..15 / chroma row 7 of row y in
// the frame (row y's own warp stores rows 0..12 / 0..6 unless it is the last row).  Every frame byte therefore has
// one writer per launch, there is no release/acquire pair in the kernel, and the samples above an MB arrive with the
// notification instead of one more round trip after it.  Epoch stamps make clearing unnecessary.
//
// Shared-memory tile per half: luma 16 rows x 48 B (own 16 samples at byte 16; 48 keeps the 128-bit row accesses of a
// quarter warp on distinct banks), chroma 2 planes x 8 rows x 16 B (own 8 samples at byte 8), plus the mailbox words
// of the MB above (4 luma rows x 16 B, 2 x 2 chroma rows x 8 B).  The halves are skewed so that the byte accesses of
// the column pass (all 32 lanes in one wavefront) fall on disjoint banks.
constexpr int kTileP = 48;                                   // luma tile row pitch
constexpr int kTileHalfY = 16 * kTileP + 16;                 // half B starts 4 banks further
constexpr int kTilePlaneC = 8 * 16 + 8;                      // chroma plane stride (2 banks further)
constexpr int kTileHalfC = 2 * kTilePlaneC + 16;             // = 288 = 32 (mod 128)
struct __align__(16) DeblockSmem {
    uint8_t y[2 * kTileHalfY];
    uint8_t c[2 * kTileHalfC];
    uint8_t top_y[2][4 * 16];
    uint8_t top_c[2][2][2 * 8];
};
constexpr int kMboxWords = 24;                               // per MB: 16 luma words (rows 12..15) + 8 chroma words (rows 6, 7 of Cb, Cr)

// filter_strong / filter_normal (deblock.cc:327-415) on the samples across one edge: p[0] = p0 ... p[3] = p3.
template <bool kChroma>
__device__ __forceinline__ void filter_edge(int bS, uint32_t par, int (&p)[4], int (&q)[4])
{
    if (bS == 0) return;
    const int alpha = par & 0xFF, beta = (par >> 8) & 31;
    const int p0 = p[0], p1 = p[1], q0 = q[0], q1 = q[1];
    if (!(abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta)) return;
    const int tc0 = (par >> (8 + 5 * bS)) & 31;                  // bS 1..3 (unused for bS 4)
    if (kChroma) {
        if (bS == 4) {
            p[0] = (2 * p1 + p0 + q1 + 2) >> 2;
            q[0] = (2 * q1 + q0 + p1 + 2) >> 2;
        } else {
            const int tc = tc0 + 1;
            const int delta = clip3i(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
            p[0] = clip255(p0 + delta);
            q[0] = clip255(q0 - delta);
        }
        return;
    }
    const int p2 = p[2], q2 = q[2];
    const bool ap = abs(p2 - p0) < beta, aq = abs(q2 - q0) < beta;
    if (bS == 4) {
        const bool small = abs(p0 - q0) < (alpha >> 2) + 2;
        if (ap && small) {
            p[0] = (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3;
            p[1] = (p2 + p1 + p0 + q0 + 2) >> 2;
            p[2] = (2 * p[3] + 3 * p2 + p1 + p0 + q0 + 4) >> 3;
        } else p[0] = (2 * p1 + p0 + q1 + 2) >> 2;
        if (aq && small) {
            q[0] = (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3;
            q[1] = (p0 + q0 + q1 + q2 + 2) >> 2;
            q[2] = (2 * q[3] + 3 * q2 + q1 + q0 + p0 + 4) >> 3;
        } else q[0] = (2 * q1 + q0 + p1 + 2) >> 2;
        return;
    }
    const int tc = tc0 + (ap ? 1 : 0) + (aq ? 1 : 0);
    const int delta = clip3i(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
    p[0] = clip255(p0 + delta);
    q[0] = clip255(q0 - delta);
    if (ap) p[1] = p1 + clip3i(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - (p1 * 2)) >> 1);
    if (aq) q[1] = q1 + clip3i(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - (q1 * 2)) >> 1);
}

__device__ __forceinline__ void unpack4(uint32_t w, int* v)
{
    v[0] = w & 0xFF; v[1] = __byte_perm(w, 0, 0x4441); v[2] = __byte_perm(w, 0, 0x4442); v[3] = w >> 24;
}
__device__ __forceinline__ uint32_t pack4(const int* v)
{
    return __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
}

__global__ void __launch_bounds__(kWarpsPerCta * 32, H264R_DEBLOCK_CTAS)
deblock_kernel(const DevPicture* __restrict__ pics, int num_pics, int* tickets, FrameGeom g, uint32_t epoch)
{
    __shared__ __align__(16) DeblockSmem smem_all[kWarpsPerCta];
    __shared__ int s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(&tickets[1], 1);
    __syncthreads();
    const int W = g.width_mbs, H = g.height_mbs;
    const int groups = (H + kWarpsPerCta - 1) / kWarpsPerCta;
    const int npairs = (num_pics + 1) >> 1;
    const int rg = s_ticket / npairs, pair = s_ticket - rg * npairs;            // row-group-major, see recon_intra_kernel
    if (rg >= groups) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mby = rg * kWarpsPerCta + warp;
    if (mby >= H) return;
    const int half = lane >> 4, l = lane & 15, cpl = l >> 3, cl = l & 7;
    const int pic_i = pair * 2 + half;
    const DevPicture& pic = pics[min(pic_i, num_pics - 1)];
    const bool enabled = pic_i < num_pics && pic.run_deblock;                   // this half has a picture to filter
    if (!__any_sync(0xFFFFFFFFu, enabled)) return;
    DeblockSmem& sm = smem_all[warp];
    uint8_t* const TY = sm.y + half * kTileHalfY;
    uint8_t* const TCh = sm.c + half * kTileHalfC;         // both planes of this half
    uint8_t* const TC = TCh + cpl * kTilePlaneC;           // this lane's plane
    uint8_t* const TOPY = sm.top_y[half];
    uint8_t* const TOPC = sm.top_c[half][0];
    uint8_t* const dY = pic.dst;
    uint8_t* const dC = pic.dst + (cpl ? g.off_cr : g.off_cb);
    const int pitch_y = g.pitch_y, pitch_c = g.pitch_c;
    const uint4* const desc = reinterpret_cast<const uint4*>(pic.desc + (size_t)mby * W);
    const int py = mby * 16, cy = mby * 8;
    const int gshY = (l >> 2) * 4, gshC = (cl >> 1) * 4;   // nibble position of this lane's 4-sample group
    const bool has_above = mby > 0, has_below = mby + 1 < H;                    // warp-uniform
    const bool own_y = enabled && (l <= 12 || !has_below);                      // frame rows this warp stores itself
    const bool own_c = enabled && (cl <= 6 || !has_below);
    uint64_t* const box_out = pic.mbox + (size_t)mby * W * kMboxWords;          // posted by this row
    const uint64_t* const box_in = pic.mbox + (size_t)(has_above ? mby - 1 : 0) * W * kMboxWords;
    // mailbox word this lane posts: luma word l = row 12 + (l >> 2), samples 4 (l & 3)..; chroma word l (l < 8) = plane
    // l >> 2, row 6 + ((l >> 1) & 1), samples 4 (l & 1)..  The last word of a row is final only after the next MB's
    // left edge: it comes from the lane that owns that row in the vertical pass.
    const uint8_t* const boxsrc_y = TY + (12 + (l >> 2)) * kTileP + 16 + 4 * (l & 3);
    const uint8_t* const boxsrc_c = TCh + ((l >> 2) & 1) * kTilePlaneC + (6 + ((l >> 1) & 1)) * 16 + 8 + 4 * (l & 1);
    const int boxlane_y = (lane & 16) + 12 + (l >> 2), boxlane_c = (lane & 16) + ((l >> 2) & 1) * 8 + 6 + ((l >> 1) & 1);

    // prefetch of MB 0: descriptor (strengths, luma thresholds, thresholds of this lane's chroma plane), own samples
    // thresholds: words 4..12 of the descriptor = [Y, Cb, Cr][left edge, internal, top edge]: two 128-bit loads and one word,
    // every loaded word used (a padded row per plane left a dead destination register that the compiler recycled at once: its
    // write then waited for the whole load, 24 % of the kernel's stall samples in ncu v28/v33)
    uint4 n_bs = make_uint4(0, 0, 0, 0), n_pa = n_bs, n_pb = n_bs, n_ownY = n_bs; uint2 n_ownC = make_uint2(0, 0); uint32_t n_pz = 0;
    if (enabled) {
        n_bs = __ldg(desc); n_pa = __ldg(desc + 1); n_pb = __ldg(desc + 2); n_pz = __ldg(reinterpret_cast<const unsigned int*>(desc) + 12);
        n_ownY = __ldcg(reinterpret_cast<const uint4*>(dY + (uint32_t)((py + l) * pitch_y)));
        n_ownC = __ldcg(reinterpret_cast<const uint2*>(dC + (uint32_t)((cy + cl) * pitch_c)));
    }
    uint32_t boxY = 0, boxC = 0;                          // this lane's mailbox words of the previous MB (after its horizontal pass)

    for (int mbx = 0; mbx < W; ++mbx) {
        const uint4 bs = n_bs, ownY = n_ownY; const uint2 ownC = n_ownC;
        const uint4 parY = make_uint4(n_pa.x, n_pa.y, n_pa.z, 0u);
        const uint4 parC = cpl ? make_uint4(n_pb.z, n_pb.w, n_pz, 0u) : make_uint4(n_pa.w, n_pb.x, n_pb.y, 0u);
        const int px = mbx * 16, cx = mbx * 8;

        // mailbox of the MB above: issued now, looked at after the vertical pass
        uint64_t t0 = 0, t1 = 0;
        if (has_above && enabled) {
            t0 = ld_mbox(box_in + mbx * kMboxWords + l);
            if (l < 8) t1 = ld_mbox(box_in + mbx * kMboxWords + 16 + l);
        }
        // previous MB's last four samples of this lane's rows (final but for this MB's left edge)
        const uint32_t carryY = *reinterpret_cast<const uint32_t*>(TY + l * kTileP + 16 + 12);
        const uint32_t carryC = *reinterpret_cast<const uint32_t*>(TC + cl * 16 + 8 + 4);

        // prefetch the next MB: independent of every other MB of this kernel
        if (mbx + 1 < W && enabled) {
            n_bs = __ldg(desc + (mbx + 1) * 4); n_pa = __ldg(desc + (mbx + 1) * 4 + 1); n_pb = __ldg(desc + (mbx + 1) * 4 + 2);
            n_pz = __ldg(reinterpret_cast<const unsigned int*>(desc + (mbx + 1) * 4) + 12);
            n_ownY = __ldcg(reinterpret_cast<const uint4*>(dY + (uint32_t)((py + l) * pitch_y + px + 16)));
            n_ownC = __ldcg(reinterpret_cast<const uint2*>(dC + (uint32_t)((cy + cl) * pitch_c + cx + 8)));
        }

        // ---- vertical edges, in registers ----
        uint32_t leftY, leftC;                             // the previous MB's last four samples after this MB's left edge
        {
            int v[20];
            unpack4(carryY, v); unpack4(ownY.x, v + 4); unpack4(ownY.y, v + 8); unpack4(ownY.z, v + 12); unpack4(ownY.w, v + 16);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int s = enabled ? ((e < 2 ? bs.x : bs.y) >> ((e & 1) * 16 + gshY)) & 7 : 0;
                int p[4] = { v[4 * e + 3], v[4 * e + 2], v[4 * e + 1], v[4 * e] }, q[4] = { v[4 * e + 4], v[4 * e + 5], v[4 * e + 6], v[4 * e + 7] };
                filter_edge<false>(s, e ? parY.y : parY.x, p, q);
                v[4 * e + 2] = p[1]; v[4 * e + 1] = p[2]; v[4 * e + 3] = p[0];
                v[4 * e + 4] = q[0]; v[4 * e + 5] = q[1]; v[4 * e + 6] = q[2];
            }
            *reinterpret_cast<uint4*>(TY + l * kTileP + 16) = make_uint4(pack4(v + 4), pack4(v + 8), pack4(v + 12), pack4(v + 16));
            leftY = pack4(v);
            if (own_y && (bs.x & 0xFFFF) && mbx > 0)          // columns 13..15 of the left MB
                *reinterpret_cast<uint32_t*>(dY + (uint32_t)((py + l) * pitch_y + px - 4)) = leftY;
        }
        {
            int v[12];
            unpack4(carryC, v); unpack4(ownC.x, v + 4); unpack4(ownC.y, v + 8);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int s = enabled ? ((e ? bs.y : bs.x) >> gshC) & 7 : 0;          // chroma edge e <- luma edge 2e
                int p[4] = { v[4 * e + 3], v[4 * e + 2], 0, 0 }, q[4] = { v[4 * e + 4], v[4 * e + 5], 0, 0 };
                filter_edge<true>(s, e ? parC.y : parC.x, p, q);
                v[4 * e + 3] = p[0]; v[4 * e + 4] = q[0];
            }
            *reinterpret_cast<uint2*>(TC + cl * 16 + 8) = make_uint2(pack4(v + 4), pack4(v + 8));
            leftC = pack4(v);
            if (own_c && (bs.x & 0xFFFF) && mbx > 0)
                *reinterpret_cast<uint32_t*>(dC + (uint32_t)((cy + cl) * pitch_c + cx - 4)) = leftC;
        }

        // ---- post the mailbox of the previous MB: its bottom rows are final now ----
        {
            const uint32_t fy = __shfl_sync(0xFFFFFFFFu, leftY, boxlane_y), fc = __shfl_sync(0xFFFFFFFFu, leftC, boxlane_c);
            if (has_below && enabled && mbx > 0) {
                st_mbox(box_out + (mbx - 1) * kMboxWords + l, (l & 3) == 3 ? fy : boxY, epoch);
                if (l < 8) st_mbox(box_out + (mbx - 1) * kMboxWords + 16 + l, (l & 1) ? fc : boxC, epoch);
            }
        }

        // ---- mailbox of the MB above: normally there already ----
        if (has_above) {
            bool waiting = enabled && ((uint32_t)(t0 >> 32) != epoch || (l < 8 && (uint32_t)(t1 >> 32) != epoch));
            unsigned ns = 16;
            while (__any_sync(0xFFFFFFFFu, waiting)) {
                if (waiting) {
                    __nanosleep(ns); if (ns < 128) ns *= 2;
                    t0 = ld_mbox(box_in + mbx * kMboxWords + l);
                    if (l < 8) t1 = ld_mbox(box_in + mbx * kMboxWords + 16 + l);
                    waiting = (uint32_t)(t0 >> 32) != epoch || (l < 8 && (uint32_t)(t1 >> 32) != epoch);
                }
            }
            reinterpret_cast<uint32_t*>(TOPY)[l] = (uint32_t)t0;           // row 12 + (l >> 2), samples 4 (l & 3)..
            if (l < 8) reinterpret_cast<uint32_t*>(TOPC)[l] = (uint32_t)t1;
        }
        __syncwarp();                                      // tile rows (vertical pass) and the rows above visible to the column owners

        // ---- horizontal edges: lane l = luma column l, chroma column cl of plane cpl ----
        uint32_t upY = 0, upC = 0;                         // samples of the MB above after the top edge
        {
            int v[20];
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = has_above ? TOPY[r * 16 + l] : 0;
#pragma unroll
            for (int r = 0; r < 16; ++r) v[4 + r] = TY[r * kTileP + 16 + l];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int s = enabled ? ((e < 2 ? bs.z : bs.w) >> ((e & 1) * 16 + gshY)) & 7 : 0;
                int p[4] = { v[4 * e + 3], v[4 * e + 2], v[4 * e + 1], v[4 * e] }, q[4] = { v[4 * e + 4], v[4 * e + 5], v[4 * e + 6], v[4 * e + 7] };
                filter_edge<false>(s, e ? parY.y : parY.z, p, q);
                v[4 * e + 2] = p[1]; v[4 * e + 1] = p[2]; v[4 * e + 3] = p[0];
                v[4 * e + 4] = q[0]; v[4 * e + 5] = q[1]; v[4 * e + 6] = q[2];
            }
#pragma unroll
            for (int r = 0; r < 15; ++r) TY[r * kTileP + 16 + l] = (uint8_t)v[4 + r];
            upY = (uint32_t)v[1] | (uint32_t)v[2] << 8 | (uint32_t)v[3] << 16;     // rows 13..15 of the MB above
        }
        {
            int v[10];                                     // rows -2, -1, 0..7
            v[0] = has_above ? TOPC[cpl * 16 + cl] : 0; v[1] = has_above ? TOPC[cpl * 16 + 8 + cl] : 0;
#pragma unroll
            for (int r = 0; r < 8; ++r) v[2 + r] = TC[r * 16 + 8 + cl];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int s = enabled ? ((e ? bs.w : bs.z) >> gshC) & 7 : 0;
                int p[4] = { v[4 * e + 1], v[4 * e], 0, 0 }, q[4] = { v[4 * e + 2], v[4 * e + 3], 0, 0 };
                filter_edge<true>(s, e ? parC.y : parC.z, p, q);
                v[4 * e + 1] = p[0]; v[4 * e + 2] = q[0];
            }
            TC[0 * 16 + 8 + cl] = (uint8_t)v[2]; TC[3 * 16 + 8 + cl] = (uint8_t)v[5]; TC[4 * 16 + 8 + cl] = (uint8_t)v[6];
            upC = (uint32_t)v[1];
        }
        __syncwarp();                                      // the tile holds the MB after both passes

        // ---- write back: rows 13..15 / row 7 of the MB above (this warp is their only writer), then the MB's own rows ----
        if (has_above && enabled) {
            uint8_t* ty = dY + (uint32_t)((py - 3) * pitch_y + px + l);
            ty[0] = (uint8_t)upY; ty[pitch_y] = (uint8_t)(upY >> 8); ty[2 * pitch_y] = (uint8_t)(upY >> 16);
            dC[(uint32_t)((cy - 1) * pitch_c + cx + cl)] = (uint8_t)upC;
        }
        if (own_y) *reinterpret_cast<uint4*>(dY + (uint32_t)((py + l) * pitch_y + px)) = *reinterpret_cast<const uint4*>(TY + l * kTileP + 16);
        if (own_c) *reinterpret_cast<uint2*>(dC + (uint32_t)((cy + cl) * pitch_c + cx)) = *reinterpret_cast<const uint2*>(TC + cl * 16 + 8);
        // this lane's mailbox words of the MB (their last samples are replaced after the next MB's left edge)
        boxY = *reinterpret_cast<const uint32_t*>(boxsrc_y);
        boxC = *reinterpret_cast<const uint32_t*>(boxsrc_c);
        __syncwarp();                                      // before the next vertical pass overwrites the tile rows
    }
    // the last MB of the row has no right neighbour: its bottom rows are final as they stand
    if (has_below && enabled) {
        st_mbox(box_out + (W - 1) * kMboxWords + l, boxY, epoch);
        if (l < 8) st_mbox(box_out + (W - 1) * kMboxWords + 16 + l, boxC, epoch);
    }
}

// ---------------------------------------------------------------------------------------------------

int launch_wave_kernel(const WaveLaunch& w, int which, cudaStream_t stream)
{
    const int nmb = w.geom.width_mbs * w.geom.height_mbs;
    const int threads = kWarpsPerCta * 32;
    const int groups = (w.geom.height_mbs + kWarpsPerCta - 1) / kWarpsPerCta;
    if (which == KERNEL_RESID) {
        const int n = 1;
        residual_kernel<<<dim3((nmb + kResidWarps - 1) / kResidWarps, 1, w.num_pics), kResidWarps * 32, 0, stream>>>(w.pics, w.geom);
        return n;
    }
    if (which == KERNEL_INTER) {
        if (!w.any_inter) return 0;
#if H264R_INTER_TWO_MB
        const dim3 grid((w.geom.width_mbs + 2 * kInter2Warps - 1) / (2 * kInter2Warps), w.geom.height_mbs, w.num_pics);
        recon_inter2_kernel<<<grid, kInter2Warps * 32, 0, stream>>>(w.pics, w.geom, w.direct8x8);
#else
        const dim3 grid((w.geom.width_mbs + kWarpsPerCta - 1) / kWarpsPerCta, w.geom.height_mbs, w.num_pics);
        recon_inter_kernel<<<grid, threads, 0, stream>>>(w.pics, w.geom, w.direct8x8);
#endif
        return 1;
    }
    if (which == KERNEL_INTRA) {
        int n = 0;
        if (w.any_intra_rows) {
            recon_intra_kernel<<<w.num_pics * groups, threads, 0, stream>>>(w.pics, w.num_pics, w.tickets, w.geom, w.epoch);
            ++n;
        }
        if (w.max_intra_sparse > 0) {
            const int per_cta = kSparseWarps * H264R_SPARSE_PER_WARP;
            const int grps = (w.max_intra_sparse + per_cta - 1) / per_cta;
            recon_intra_sparse_kernel<<<w.num_pics * grps, kSparseWarps * 32, 0, stream>>>(w.pics, w.num_pics, w.tickets, w.geom, w.epoch);
            ++n;
        }
        return n;
    }
    if (!w.any_deblock) return 0;
    if (which == KERNEL_DBPREP) {
        const long long total = (long long)w.num_pics * nmb;
        deblock_prep_kernel<<<(int)((total + 127) / 128), 128, 0, stream>>>(w.pics, w.num_pics, w.geom);
        return 1;
    }
    deblock_kernel<<<((w.num_pics + 1) / 2) * groups, threads, 0, stream>>>(w.pics, w.num_pics, w.tickets, w.geom, w.epoch);
    return 1;
}

} // namespace h264r
