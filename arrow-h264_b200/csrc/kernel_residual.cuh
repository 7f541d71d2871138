// residual_kernel: one warp per MB that received levels: scatter-dequantise the level list into shared memory, luma 4x4 /
// chroma 2x2 DC Hadamards, 4x4 / 8x8 inverse transforms only for blocks that received a level, saturating pack to int16
// -> the picture's residual plane [nmb][384] (transform.cc:394-456, 460-554, 597-733, 825-910).
//                  It is also the one place where the description of a picture is checked (validate_and_repair).
// deblock_prep_kernel: boundary strengths and alpha / beta / tc0 per MB (deblock.cc:35-289, 469-474), one thread per MB.
// Both need nothing but the picture description: they run on the side stream, a wave ahead of the reconstruction.
#ifndef H264R_KERNEL_RESIDUAL_CUH_
#define H264R_KERNEL_RESIDUAL_CUH_

#include "kernels_common.cuh"

namespace h264r {

#ifndef H264R_RESID_CTAS
#define H264R_RESID_CTAS 14
#endif
#ifndef H264R_RESID_WARPS
#define H264R_RESID_WARPS 4
#endif
#ifndef H264R_PREP_CTAS
#define H264R_PREP_CTAS 12
#endif

struct __align__(16) ResidSmem { int cof[kResInts]; };

// Residual of one MB by one warp into `res` (the coefficient scratch, kernels_common.cuh) and from there, clamped to
// [-255, 255] (clip(pred + res) cannot tell the difference), as 48 x 16 bytes to `out`.
__device__ __forceinline__ void residual_mb(const DevPicture& pic, const MbHdr& h, int* res, uint4* out, int lane, uint32_t* err)
{
    const h264r_slice* __restrict__ sl = pic.slices + h.slice_idx;
    const int inter = h.intra() ? 0 : 1;
    const bool t8 = h.t8();
    const bool i16 = h.mb_type == H264R_MB_I16x16;
    const int per = h.qp_y / 6, rem = h.qp_y - per * 6;
#pragma unroll
    for (int k = 0; k < (kResInts / 4 + 31) / 32; ++k)
        if (lane + 32 * k < kResInts / 4) reinterpret_cast<int4*>(res)[lane + 32 * k] = make_int4(0, 0, 0, 0);
    __syncwarp();
    const uint32_t* __restrict__ lv = pic.stream + h.coeff_offset;
    const uint32_t ctl = scatter_ctl(h), mode = scatter_mode(h, inter);
    unsigned nz = 0;
    for (int i = lane; i < h.coeff_count; i += 32) nz |= scatter_level(__ldg(lv + i), ctl, mode, sl, res, err);
    nz = __reduce_or_sync(0xFFFFFFFFu, nz);
    __syncwarp();

    // DC transforms (transform_luma_dc :825-856, transform_chroma_dc :858-910).  Luma (Intra16x16): lane j < 16 holds DC
    // (j >> 2, j & 3), both Hadamard passes run through shuffles (every lane busy instead of lane 0 alone).  Chroma: lanes
    // 16..23 = (plane, block); each evaluates the 2x2 Hadamard for its own block.
    int dc = 0;
    const bool cdc = h.cbp_chroma && (nz >> 16);
    if (i16) {
        const int j = lane & 15;
        const int c = res[(j >> 2) * 4 * kResP + (j & 3) * 4];
        const int base = lane & ~3, k = lane & 3;
        const int c0 = __shfl_sync(0xFFFFFFFFu, c, base), c1 = __shfl_sync(0xFFFFFFFFu, c, base + 1),
                  c2 = __shfl_sync(0xFFFFFFFFu, c, base + 2), c3 = __shfl_sync(0xFFFFFFFFu, c, base + 3);
        const int e = k == 0 ? c0 + c1 + c2 + c3 : (k == 1 ? c0 + c1 - c2 - c3 : (k == 2 ? c0 - c1 - c2 + c3 : c0 - c1 + c2 - c3));
        const int col = lane & 3, r = (lane >> 2) & 3, hb = lane & 16;
        const int e0 = __shfl_sync(0xFFFFFFFFu, e, hb + col), e1 = __shfl_sync(0xFFFFFFFFu, e, hb + 4 + col),
                  e2 = __shfl_sync(0xFFFFFFFFu, e, hb + 8 + col), e3 = __shfl_sync(0xFFFFFFFFu, e, hb + 12 + col);
        const int f = r == 0 ? e0 + e1 + e2 + e3 : (r == 1 ? e0 + e1 - e2 - e3 : (r == 2 ? e0 - e1 - e2 + e3 : e0 - e1 + e2 - e3));
        const int scale = (int)__ldg(&sl->level_scale_4x4[0][0][rem][0]);
        dc = h.qp_y >= 36 ? (f * scale) * (1 << (per - 6)) : (f * scale + (1 << (5 - per))) >> (6 - per);
        nz |= 0xFFFFu;                                       // the DC Hadamard spreads into every luma block
    }
    if (cdc) {
        if (lane >= 16 && lane < 24) {
            const int pl = (lane - 16) >> 2, qb = lane & 3;
            const int qc = pl ? h.qp_c[1] : h.qp_c[0], cper = qc / 6, crem = qc - cper * 6;
            const int* c = res + kResC + pl * kResCPlane;             // DC positions (0,0) (0,4) (4,0) (4,4)
            dc = chroma_dc_of_block(qb, c[0], c[4], c[4 * kResCP], c[4 * kResCP + 4], (int)__ldg(&sl->level_scale_4x4[inter][pl + 1][crem][0]), cper);
        }
        nz |= 0xFF0000u;
    }
    __syncwarp();                                            // every raw DC has been read
    if (i16 && lane < 16) res[(lane >> 2) * 4 * kResP + (lane & 3) * 4] = dc;
    if (cdc && lane >= 16 && lane < 24) {
        const int pl = (lane - 16) >> 2, qb = lane & 3;
        res[kResC + pl * kResCPlane + (qb >> 1) * 4 * kResCP + (qb & 1) * 4] = dc;
    }
    __syncwarp();

    // inverse transforms, only where something is non-zero
    if (t8) {
        const int b8 = lane >> 3, i = lane & 7;
        int* blk = res + (b8 >> 1) * 8 * kResP + (b8 & 1) * 8;
        const unsigned m8 = 0x33u << ((b8 >> 1) * 8 + (b8 & 1) * 2);          // the four 4x4 blocks of 8x8 block b8
        if (nz & m8) idct8_1d(blk + i * kResP, 1, false);
        __syncwarp();
        if (nz & m8) idct8_1d(blk + i, kResP, true);
        if (lane < 8 && ((nz >> (16 + lane)) & 1)) {
            int d[4][4];
            int* cb = res + kResC + (lane >> 2) * kResCPlane + ((lane >> 1) & 1) * 4 * kResCP + (lane & 1) * 4;
            load_block4(cb, kResCP, d); idct4_regs(d); store_block4(cb, kResCP, d);
        }
    } else if (lane < 24 && ((nz >> lane) & 1)) {
        // one instruction stream for the sixteen luma blocks (lanes 0..15) and the eight chroma blocks (lanes 16..23)
        const int c = lane - 16;
        int* const blk = lane < 16 ? res + (lane >> 2) * 4 * kResP + (lane & 3) * 4
                                   : res + kResC + (c >> 2) * kResCPlane + ((c >> 1) & 1) * 4 * kResCP + (c & 1) * 4;
        const int pitch = lane < 16 ? kResP : kResCP;
        int d[4][4];
        load_block4(blk, pitch, d); idct4_regs(d); store_block4(blk, pitch, d);
    }
    __syncwarp();

    // 384 x int16 = 48 x 16 B
    auto pack8 = [&](const int* r, int v) {              // eight consecutive samples of a row -> one 16-byte store
        out[v] = make_uint4(pack_res2(r[0], r[1]), pack_res2(r[2], r[3]), pack_res2(r[4], r[5]), pack_res2(r[6], r[7]));
    };
    pack8(res + (lane >> 1) * kResP + (lane & 1) * 8, lane);                                       // luma row lane >> 1, half lane & 1
    if (lane < 16) pack8(res + kResC + (lane >> 3) * kResCPlane + (lane & 7) * kResCP, 32 + lane);   // plane lane >> 3, row lane & 7
}

constexpr int kResidWarps = H264R_RESID_WARPS;
// grid = (ceil(nmb / warps), 1, pictures): no index divisions
__global__ void __launch_bounds__(kResidWarps * 32, H264R_RESID_CTAS * 4 / H264R_RESID_WARPS)
residual_kernel(const DevPicture* __restrict__ pics, FrameGeom g, uint32_t* err)
{
    __shared__ __align__(16) ResidSmem smem_all[kResidWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nmb = g.width_mbs * g.height_mbs;
    const int addr = blockIdx.x * kResidWarps + warp;
    if (addr >= nmb) return;
    const DevPicture& pic = pics[blockIdx.z];
    MbHdr h = load_hdr(pic.mbs, addr);
    validate_and_repair(h, pic, addr, lane, err);
    if (!h.has_resid()) return;
    residual_mb(pic, h, smem_all[warp].cof, reinterpret_cast<uint4*>(pic.resid + (size_t)addr * H264R_COEFFS_PER_MB), lane, err);
}

// ---- deblock descriptors (fully parallel): per MB boundary strengths + filter thresholds ----

// Deblock::strength (deblock.cc:78-289) + the qPav/indexA/indexB/alpha/beta/tc0 part of filter_edge (deblock.cc:469-474,
// tables :294-324).  One THREAD per MB (the work is scalar: 32 strengths and 9 threshold sets out of three MB headers).
struct HdrLite { int mb_type, flags, slice_idx, cbp_blks; uint32_t w1, w2, packed; };
__device__ __forceinline__ HdrLite load_hdr_lite(const h264r_mb* mbs, int addr)
{
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(mbs + addr));
    HdrLite h;
    h.mb_type = a.x & 0xFF; h.flags = (a.x >> 8) & 0xFF; h.slice_idx = a.x >> 16;
    h.w1 = a.y; h.w2 = a.z;
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
    const uint32_t s0 = __ldg(reinterpret_cast<const uint32_t*>(sl));      // slice_type | idc << 8 | FilterOffsetA << 16 | FilterOffsetB << 24
    const int idc = (s0 >> 8) & 0xFF;
    uint4* out = reinterpret_cast<uint4*>(pic.desc + q);
    if (idc == 1) { out[0] = make_uint4(0, 0, 0, 0); return; }           // no strengths: the thresholds are never read

    bool left = mbx > 0, top = mby > 0;
    HdrLite PL = Q, PT = Q;
    if (left) { PL = load_hdr_lite(pic.mbs, q - 1); if (idc == 2 && PL.slice_idx != Q.slice_idx) left = false; }
    if (top)  { PT = load_hdr_lite(pic.mbs, q - W); if (idc == 2 && PT.slice_idx != Q.slice_idx) top = false; }
    const bool q_intra = Q.flags & H264R_MB_FLAG_INTRA, t8 = Q.flags & H264R_MB_FLAG_T8x8;
    const bool p_skip = (s0 & 0xFF) == H264R_P_SLICE && Q.mb_type == 0;
    const int mvlimit = pic.field ? 2 : 4;

    // (loops kept rolled: unrolled, the kernel was 77 KB of code for a 32 KB instruction cache)
    uint32_t bs0 = 0, bs1 = 0, bs2 = 0, bs3 = 0;
    auto bs_or = [&](int wi, uint32_t v) { if (wi == 0) bs0 |= v; else if (wi == 1) bs1 |= v; else if (wi == 2) bs2 |= v; else bs3 |= v; };
#pragma unroll 1
    for (int dir = 0; dir < 2; ++dir) {
        const bool mbedge = dir == 0 ? left : top;
        const HdrLite& PN = dir == 0 ? PL : PT;
#pragma unroll 1
        for (int e = 0; e < 4; ++e) {
            const bool on = e == 0 ? mbedge : !(t8 && (e & 1));
            if (!on) continue;
            if (e > 0 && p_skip) continue;
            const int wi = dir * 2 + (e >> 1), sh = (e & 1) * 16;
            const bool p_intra = e ? q_intra : (PN.flags & H264R_MB_FLAG_INTRA) != 0;
            // bS 4 on MB edges -- in field pictures only on the vertical ones (cond_bS4, deblock.cc:106-107, 188-189)
            if (p_intra || q_intra) { bs_or(wi, (e == 0 && !(pic.field && dir == 1) ? 0x4444u : 0x3333u) << sh); continue; }
            const int pcbp = e ? Q.cbp_blks : PN.cbp_blks;
            const bool same_part = e > 0 && (Q.mb_type == 1 || Q.mb_type == (dir == 0 ? 2 : 3));
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                const int blkQ = dir == 0 ? k4 * 4 + e : e * 4 + k4;
                const int blkP = dir == 0 ? k4 * 4 + (e ? e - 1 : 3) : (e ? e - 1 : 3) * 4 + k4;
                uint32_t v = 0;
                if (((Q.cbp_blks >> blkQ) & 1) || ((pcbp >> blkP) & 1)) v = 2;
                else if (!same_part) {
                    const uint32_t wp = packed_entry_word(e ? Q.packed : PN.packed, blkP), wq = packed_entry_word(Q.packed, blkQ);
                    if (wp != wq) {                              // the same entry: same pictures, same vectors
                        const uint32_t* ep = pic.stream + wp; const uint32_t* eq = pic.stream + wq;
                        v = bs_compare(__ldg(ep), __ldg(ep + 1), __ldg(ep + 2), __ldg(eq), __ldg(eq + 1), __ldg(eq + 2), mvlimit);
                    }
                }
                bs_or(wi, v << (sh + k4 * 4));
            }
        }
    }
    out[0] = make_uint4(bs0, bs1, bs2, bs3);
    // thresholds: type 0 = left MB edge, 1 = internal edge, 2 = top MB edge
    const int foa = (int)(int8_t)(s0 >> 16), fob = (int)(int8_t)(s0 >> 24);
    uint32_t w[12];                                       // [plane][type], contiguous: no padding words
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const HdrLite& P = t == 0 ? PL : (t == 2 ? PT : Q);
            w[pl * 3 + t] = deblock_threshold_word(qp_of_plane(P.w1, P.w2, pl), qp_of_plane(Q.w1, Q.w2, pl), foa, fob);
        }
    }
    out[1] = make_uint4(w[0], w[1], w[2], w[3]);
    out[2] = make_uint4(w[4], w[5], w[6], w[7]);
    reinterpret_cast<uint32_t*>(out)[12] = w[8];
}

} // namespace h264r
#endif
