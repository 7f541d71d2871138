// recon_inter2_kernel: the inter macroblocks of a wave of pictures, TWO macroblocks per warp, one 4x4 block per lane:
// motion compensation, weighted prediction, residual add (inter_prediction.cc:53-406, 448-536; decoder.cc:217-262;
// transform.cc:913-984).
#ifndef H264R_KERNEL_INTER_CUH_
#define H264R_KERNEL_INTER_CUH_

#include "kernels_common.cuh"

namespace h264r {

#ifndef H264R_INTER2_WARPS
#define H264R_INTER2_WARPS 2
#endif
#ifndef H264R_INTER2_CTAS
#define H264R_INTER2_CTAS (28 / H264R_INTER2_WARPS)
#endif
// -DH264R_INTER_TMA=1: the windows of uniform 8x8 quadrants that lie inside the picture are fetched with TMA
// (cp.async.bulk.tensor, one 32 x 13 luma box and one 32 x 5 x 2 chroma box per quadrant, completion on an mbarrier)
// instead of 13 + 5 four-byte cp.async per lane.  A box must start on a 16-byte boundary, so it starts at x0 & ~15 and is 32
// bytes wide.  Measured against the cp.async build in round 2: DESIGN.md section 3.
#ifndef H264R_INTER_TMA
#define H264R_INTER_TMA 0
#endif

// Reference windows in shared memory.  Interior windows are fetched as aligned 32-bit words: the first sample x0 of
// a window row then sits at byte offset x0 & 3.  Windows touching the picture border (rare) are fetched sample by
// sample with clamped coordinates -- bit-identical to the reference's padded planes + block pre-clamp, SURVEY.md
// 8a -- and start at byte offset 0; that path is kept out of line.
__device__ __noinline__ void load_window_border(uint32_t* win, int pitch_words, const uint8_t* __restrict__ plane, int pitch,
                                                int W, int H, int x0, int y0, int ncols, int nrows, int first_row, int row_step)
{
    // word by word: four samples that lie inside the row are one unaligned 32-bit read (two aligned loads + funnel shift),
    // only words that straddle or leave the row are clamped sample by sample (W is a multiple of 8: the second aligned
    // word of an inside read is inside too)
    const int nwords = (ncols + 3) >> 2;
    for (int row = first_row; row < nrows; row += row_step) {
        const uint8_t* src = plane + (uint32_t)(clip3i(0, H - 1, y0 + row) * pitch);
        uint32_t* dst = win + row * pitch_words;
        for (int j = 0; j < nwords; ++j) {
            const int xs = x0 + 4 * j;
            uint32_t v;
            if (xs >= 0 && xs + 3 < W) {
                const uint32_t* p = reinterpret_cast<const uint32_t*>(src + (xs & ~3));
                const uint32_t lo = __ldg(p), hi = (xs & 3) ? __ldg(p + 1) : 0u;
                v = __funnelshift_r(lo, hi, (xs & 3) * 8);
            } else {
                v = (uint32_t)__ldg(src + clip3i(0, W - 1, xs)) | (uint32_t)__ldg(src + clip3i(0, W - 1, xs + 1)) << 8 |
                    (uint32_t)__ldg(src + clip3i(0, W - 1, xs + 2)) << 16 | (uint32_t)__ldg(src + clip3i(0, W - 1, xs + 3)) << 24;
            }
            dst[j] = v;
        }
    }
}
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
#if !H264R_INTER_TMA
constexpr int kLumaQ = 146, kChromaQ = 50, kChromaUniPitch = 2, kLumaUniPitch = 4;
struct __align__(16) Inter2Smem {
    uint32_t luma[2][4 * kLumaQ + 2];
    uint32_t chroma[2][4 * kChromaQ + 2];
};
#else
// TMA destinations must be 128-byte aligned: a quadrant owns 640 bytes of luma (the 32 x 13 box = row pitch 8 words, or four
// 9-row block windows of pitch 4) and 384 bytes of chroma (the 32 x 5 x 2 box = row pitch 8 words, or the split layout).  Every
// quadrant therefore starts on bank 0: the eight quadrants of a warp read the same banks (the cp.async layout staggers
// them by 146 words).
constexpr int kLumaQ = 160, kChromaQ = 96, kChromaUniPitch = 8, kLumaUniPitch = 8;
struct __align__(128) Inter2Smem {
    uint32_t luma[2][4 * kLumaQ];
    uint32_t chroma[2][4 * kChromaQ];
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar, uint32_t tx_bytes)
{
    if (tx_bytes) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(tx_bytes) : "memory");
    else          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_addr(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_addr(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// Decoder::mb_pred_inter partition walk (decoder.cc:217-262) for 4x4 block `blk`: returns the block whose motion
// entry the reference reads (partition origin), the prediction direction, and whether the partition covers the
// whole 8x8 quadrant of the block.  Partition steps in 4x4 units per type 0..7 ({0,0},{4,4},{4,2},{2,4},{2,2},{2,1},
// {1,2},{1,1}) are nibbles of two constants.
__device__ __forceinline__ void partition_of_block2(const MbHdr& h, int is_b, int direct_spatial, const DevPicture& pic, int direct8x8,
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
        const uint32_t rw = __ldg(pic.stream + packed_entry_word(h.packed, j0 * 4 + i0) + 2);
        pd = (int8_t)(rw >> 8) < 0 ? 0 : ((int8_t)rw < 0 ? 1 : 2);
    }
    const int i = bx & ~(sh4 - 1), j = by & ~(sv4 - 1);   // partitions are aligned to their own size
    origin = j * 4 + i;
    dir = pd;
    covers8x8 = sh4 >= 2 && sv4 >= 2;
}

constexpr int kInter2Warps = H264R_INTER2_WARPS;       // warps per CTA (each warp: two MBs)

// Lanes 0..15 reconstruct MB 2j, lanes 16..31 MB 2j+1 of a row; lane b of a half owns 4x4 luma block b (raster) and the
// 2x2 chroma patches of both planes under it.  The per-MB work that is the same for every lane (header, slice,
// partition walk, addressing, weights, loop control) is issued once for two MBs.  If one partition covers an 8x8
// quadrant its four lanes share one 13x13 luma / 5x5 chroma window, otherwise every block has its own 9x9 / 3x3 window.
// grid = (ceil(width_mbs / (2 * warps)), height_mbs, pictures of the wave)
// kField: some picture of the wave is a field picture (the chroma vector offset between fields of different parity is compiled
// in; waves of frame pictures run the instantiation without it)
template <bool kField>
__global__ void __launch_bounds__(kInter2Warps * 32, H264R_INTER2_CTAS)
recon_inter2_kernel(const DevPicture* __restrict__ pics, FrameGeom g)
{
#if H264R_INTER_TMA
    __shared__ Inter2Smem smem_all[kInter2Warps];
    __shared__ __align__(8) uint64_t bar_all[kInter2Warps];
#else
    __shared__ __align__(16) Inter2Smem smem_all[kInter2Warps];
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = lane >> 4, b = lane & 15, bx = b & 3, by = b >> 2;
    const int mbx = (blockIdx.x * kInter2Warps + warp) * 2 + m, mby = blockIdx.y;
    const DevPicture& pic = pics[blockIdx.z];
    if (pic.all_intra) return;
    const int W = g.width_mbs;
    if ((blockIdx.x * kInter2Warps + warp) * 2 >= W) return;
    const int addr = mby * W + min(mbx, W - 1);
    const MbHdr h = load_hdr(pic.mbs, addr);
    const bool valid = mbx < W && !h.intra();
    if (!__any_sync(0xFFFFFFFFu, valid)) return;
    Inter2Smem& sm = smem_all[warp];
#if H264R_INTER_TMA
    uint64_t* const bar = &bar_all[warp];                 // one mbarrier per warp: 32 arrivals + the bytes of the warp's boxes per phase
    if (lane == 0) mbar_init(bar, 32);
    __syncwarp();
    uint32_t phase = 0;
#endif
    const h264r_slice* __restrict__ sl = pic.slices + h.slice_idx;
    const int wY = W * 16, hY = g.height_mbs * 16, wC = wY >> 1, hC = hY >> 1;
    const uint32_t s0 = __ldg(reinterpret_cast<const uint32_t*>(sl)), s1 = __ldg(reinterpret_cast<const uint32_t*>(sl) + 1),
                   s2 = __ldg(reinterpret_cast<const uint32_t*>(sl) + 2);
    const bool is_b = (s0 & 0xFF) == H264R_B_SLICE;
    const int denom_y = s1 & 0xFF, denom_c = (s1 >> 8) & 0xFF, wp_flag = (s1 >> 16) & 0xFF, bipred_idc = (s1 >> 24) & 0xFF;
    const int direct_spatial = (s2 >> 8) & 0xFF;
    const bool has_res = valid && h.has_resid();

    // the residual is needed last: pull its six lines into the L2 now (no registers held), the loads at the end then
    // cost an L2 hit instead of one more HBM round trip on the warp's dependent chain
    if (has_res && b < 6) prefetch_l2(pic.resid + (size_t)addr * H264R_COEFFS_PER_MB + b * 64);

    int origin = 0, pd = 0; bool uni = true;
    uint32_t mvw0 = 0, mvw1 = 0, rw = 0;                 // the motion entry of this block's partition
    if (valid) {
        partition_of_block2(h, is_b, direct_spatial, pic, pic.direct8x8, b, origin, pd, uni);
        const uint32_t* e = pic.stream + packed_entry_word(h.packed, origin);
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
        int cdy = 0;                                       // field pictures: vertical chroma offset towards a field of the other parity
        const uint32_t* wl = lq; const uint32_t* wc0 = cq; const uint32_t* wc1 = cq;
        int loff = 2, coff = 0;
        int cpitch = 2, lpitch = 4;                        // row pitches of this lane's chroma / luma windows, words
#if H264R_INTER_TMA
        uint32_t tx = 0;                                   // bytes this lane's TMA boxes will deliver
#endif
        if (active) {
            refidx = (int)(int8_t)(rw >> (8 * list));
            // the entry names the reference picture itself (ref_pic = slot of pic_params.ref_frames, the identity the
            // deblocking rule compares); slots the picture does not have point at a dummy frame
            const int slot = (int)(int8_t)(rw >> (16 + 8 * list));
            const uint8_t* __restrict__ rbase = pic.ref[slot & 31];
            const uint32_t mvw = list ? mvw1 : mvw0;
            const int mvx = (int)(int16_t)(mvw & 0xFFFF), mvy = (int)(int16_t)(mvw >> 16);
            vx = (mbx * 16 + bx * 4) * 4 + mvx; vy = (mby * 16 + by * 4) * 4 + mvy;       // this block's position
            if (kField) cdy = ((pic.ref_opposite >> (slot & 31)) & 1) ? pic.chroma_dy : 0;     // get_block_chroma, inter_prediction.cc:352-354
            if (uni) {
                const int qvx = (mbx * 16 + (bx >> 1) * 8) * 4 + mvx, qvy = (mby * 16 + (by >> 1) * 8) * 4 + mvy;
                const int x0 = (qvx >> 2) - 2, y0 = (qvy >> 2) - 2, cx0 = qvx >> 3, cy0 = (qvy + cdy) >> 3;
                const int xa = x0 & ~3, cxa = cx0 & ~3;
                const bool in_y = xa >= 0 && xa + 16 <= wY && y0 >= 0 && y0 + 13 <= hY;
                const bool in_c = cxa >= 0 && cxa + 8 <= wC && cy0 >= 0 && cy0 + 5 <= hC;
#if H264R_INTER_TMA
                const FrameMaps* __restrict__ maps = pic.ref_maps[slot & 31];
                const bool tma = __ldg(&maps->ok) != 0;
                const bool tma_y = tma && x0 >= 0 && x0 + 13 <= wY && y0 >= 0 && y0 + 13 <= hY;
                const bool tma_c = tma && cx0 >= 0 && cx0 + 5 <= wC && cy0 >= 0 && cy0 + 5 <= hC;
                if (tma_y) {                                    // one 32 x 13 box from the 16-byte boundary below x0: lane sb = 0 of the quadrant
                    if (sb == 0) { tx += 32 * 13; mbar_arrive(bar, 32 * 13); tma_load_2d(lq, &maps->luma, x0 & ~15, y0, bar); }
                } else
#else
                const bool tma_y = false, tma_c = false;
#endif
                if (in_y) {                                     // 13 rows x 4 words: lane = word column
                    const uint8_t* src = rbase + (uint32_t)(y0 * pitch_y + xa + sb * 4);
#pragma unroll
                    for (int i = 0; i < 13; ++i) cp_async4(lq + i * kLumaUniPitch + sb, src + (uint32_t)(i * pitch_y));
                } else load_window_border(lq, kLumaUniPitch, rbase, pitch_y, wY, hY, x0, y0, 13, 13, sb, 4);
                {                                               // 2 planes x 5 rows x 2 (TMA build: 4) words: lane = (plane, word column)
                    const int pl = sb >> 1, col = sb & 1;
                    const uint8_t* cplane = rbase + (pl ? g.off_cr : g.off_cb);
#if H264R_INTER_TMA
                    if (tma_c) {                                // one 32 x 5 x 2 box (both planes): lane sb = 1 of the quadrant
                        if (sb == 1) { tx += 32 * 5 * 2; mbar_arrive(bar, 32 * 5 * 2); tma_load_3d(cq, &maps->chroma, cx0 & ~15, cy0, 0, bar); }
                    } else
#endif
                    if (in_c) {
                        const uint8_t* src = cplane + (uint32_t)(cy0 * pitch_c + cxa + col * 4);
#pragma unroll
                        for (int i = 0; i < 5; ++i) cp_async4(cq + pl * 5 * kChromaUniPitch + i * kChromaUniPitch + col, src + (uint32_t)(i * pitch_c));
                    } else load_window_border(cq + pl * 5 * kChromaUniPitch, kChromaUniPitch, cplane, pitch_c, wC, hC, cx0, cy0, 5, 5, col, 2);
                }
                lpitch = kLumaUniPitch;
                wl = lq + (sb >> 1) * 4 * kLumaUniPitch;
                loff = 2 + (sb & 1) * 4 + (tma_y ? x0 & 15 : (in_y ? x0 & 3 : 0));
                cpitch = kChromaUniPitch;
                wc0 = cq + (sb >> 1) * 2 * kChromaUniPitch; wc1 = wc0 + 5 * kChromaUniPitch;
                coff = (sb & 1) * 2 + (tma_c ? cx0 & 15 : (in_c ? cx0 & 3 : 0));
            } else {
                const int x0 = (vx >> 2) - 2, y0 = (vy >> 2) - 2, cx0 = vx >> 3, cy0 = (vy + cdy) >> 3;
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
#if H264R_INTER_TMA
        if (!tx) mbar_arrive(bar, 0);                      // (lanes that issued a box arrived with their expect_tx)
        mbar_wait(bar, phase);                             // every box of the warp has landed
        phase ^= 1;
#endif
        __syncwarp();
        {
            const int xf = vx & 3, yf = vy & 3;
            unsigned hm, cm;
            mc_luma_masks_r<4>(xf, yf, hm, cm);
            const unsigned whm = __reduce_or_sync(0xFFFFFFFFu, active ? hm : 0u), wcm = __reduce_or_sync(0xFFFFFFFFu, active ? cm : 0u);
            uint32_t y[4];
            mc_luma_patch<4>(wl, loff, xf, yf, whm, wcm, y, lpitch);
            const uint32_t c0 = mc_chroma_patch_2x2(wc0, coff, vx & 7, (vy + cdy) & 7, cpitch), c1 = mc_chroma_patch_2x2(wc1, coff, vx & 7, (vy + cdy) & 7, cpitch);
            if (active) {
#pragma unroll
                for (int r = 0; r < 4; ++r) { prevY[r] = curY[r]; curY[r] = y[r]; }
                prevC[0] = curC[0]; prevC[1] = curC[1]; curC[0] = c0; curC[1] = c1;
                ref_prev = ref_cur; ref_cur = refidx;
            }
        }
        __syncwarp();
#if H264R_INTER_TMA
        fence_proxy_async();                               // the windows were read through the generic proxy; the next boxes arrive through the async proxy
#endif
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

} // namespace h264r
#endif
