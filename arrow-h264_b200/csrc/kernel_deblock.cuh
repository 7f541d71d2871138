// deblock_kernel: the in-loop deblocking filter (deblock.cc:327-552) as a row wavefront, two pictures per warp, vertical
// then horizontal edges in place; rows talk through mailboxes.  Boundary strengths and thresholds come from the 64-byte
// descriptor the reconstruction kernels wrote for every macroblock.
#ifndef H264R_KERNEL_DEBLOCK_CUH_
#define H264R_KERNEL_DEBLOCK_CUH_

#include "kernels_common.cuh"
#include "kernel_intra.cuh"          // st_mbox / ld_mbox

namespace h264r {

#ifndef H264R_DEBLOCK_CTAS
#define H264R_DEBLOCK_CTAS (16 / H264R_WARPS_PER_CTA)      // 16 resident warps per SM (115 registers)
#endif
// 1: descriptor and samples of MB x + 1 are loaded while MB x is filtered (19 registers); 0: loaded at the start of their own
// step -- for builds that trade the prefetch for more resident warps (H264R_DEBLOCK_CTAS 6 / 8)
#ifndef H264R_DEBLOCK_PREFETCH
#define H264R_DEBLOCK_PREFETCH 1
#endif
// back-off of the mailbox poll, ns: first sleep and cap
#ifndef H264R_POLL_NS0
#define H264R_POLL_NS0 16
#endif
#ifndef H264R_POLL_NS1
#define H264R_POLL_NS1 128
#endif

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
#if !H264R_DEBLOCK_PREFETCH
        if (mbx > 0 && enabled) {
            n_bs = __ldg(desc + mbx * 4); n_pa = __ldg(desc + mbx * 4 + 1); n_pb = __ldg(desc + mbx * 4 + 2);
            n_pz = __ldg(reinterpret_cast<const unsigned int*>(desc + mbx * 4) + 12);
            n_ownY = __ldcg(reinterpret_cast<const uint4*>(dY + (uint32_t)((py + l) * pitch_y + mbx * 16)));
            n_ownC = __ldcg(reinterpret_cast<const uint2*>(dC + (uint32_t)((cy + cl) * pitch_c + mbx * 8)));
        }
#endif
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
        if (H264R_DEBLOCK_PREFETCH && mbx + 1 < W && enabled) {
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
            unsigned ns = H264R_POLL_NS0;
            while (__any_sync(0xFFFFFFFFu, waiting)) {
                if (waiting) {
                    __nanosleep(ns); if (ns < H264R_POLL_NS1) ns *= 2;
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

} // namespace h264r
#endif
