// Intra reconstruction (intra_prediction.cc:137-904, decoder.cc:149-215); the residual comes from residual_kernel's plane:
//   recon_intra_kernel        : all-intra pictures, one warp per MB ROW; rows form a 2:1 wavefront (MB(x,y) needs (x-1,y),
//                               (x-1,y-1), (x,y-1), (x+1,y-1)) and talk through mailboxes
//   intra_list_kernel         : raster-ordered list of the intra MBs of every P/B picture of a wave (no list from the host)
//   recon_intra_sparse_kernel : the intra MBs of P/B pictures, one warp each; per-MB epoch stamps between intra neighbours
#ifndef H264R_KERNEL_INTRA_CUH_
#define H264R_KERNEL_INTRA_CUH_

#include "kernels_common.cuh"

namespace h264r {

#ifndef H264R_INTRA_CTAS
#define H264R_INTRA_CTAS (16 / H264R_WARPS_PER_CTA)        // 16 resident warps per SM (122 registers)
#endif

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
        const uint32_t* __restrict__ lv = pic.stream + h.coeff_offset;
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
    if (!pic.all_intra) return;                          // pictures with P / B slices: recon_intra_sparse_kernel
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
        if (mbx + 1 < W) intra_prefetch(pic, g, mbx + 1, mby, lane, nxt);   // lands while this MB is reconstructed
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

// Intra MBs of pictures that also have inter MBs (P/B pictures: a few percent of the MBs, mostly isolated).  The host does
// not look at the macroblocks at all: intra_list_kernel (side stream, one CTA per picture) compacts the addresses of the
// intra MBs of a picture, in raster order, into the picture's device list.
constexpr int kListThreads = 256;
__global__ void __launch_bounds__(kListThreads)
intra_list_kernel(const DevPicture* __restrict__ pics, FrameGeom g, uint32_t* wave_max)
{
    __shared__ uint32_t s_warp[kListThreads / 32];
    __shared__ uint32_t s_base;
    const DevPicture& pic = pics[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (pic.all_intra) { if (tid == 0) *pic.intra_count = 0; return; }
    const int nmb = g.width_mbs * g.height_mbs, words = (nmb + 31) / 32;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int chunk = 0; chunk < words; chunk += kListThreads) {
        // thread t: the intra mask of MBs 32 (chunk + t) .. + 31 (32 independent loads of header word 0)
        const int first = (chunk + tid) * 32;
        uint32_t mask = 0;
        if (first < nmb) {
#pragma unroll 8
            for (int i = 0; i < 32; ++i)
                if (first + i < nmb && ((load_hdr_word0(pic.mbs, first + i) >> 8) & H264R_MB_FLAG_INTRA)) mask |= 1u << i;
        }
        // exclusive scan of the counts over the CTA
        const uint32_t n = __popc(mask);
        uint32_t incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += v; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t off = s_base + incl - n;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        uint32_t total = 0;
        for (int w = 0; w < kListThreads / 32; ++w) total += s_warp[w];
        while (mask) { pic.intra_list[off++] = (uint32_t)(first + __ffs(mask) - 1); mask &= mask - 1; }
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
    }
    if (tid == 0) { *pic.intra_count = s_base; atomicMax(wave_max, s_base); }
}

// recon_intra_sparse_kernel: persistent warps; a warp takes a ticket = one entry of one picture's list (one warp per intra
// MB), and takes its next ticket before it starts on the current MB (the atomic's round trip hides behind the work; no
// CTA-wide barrier couples the warps: the barrier version waited 15 cycles per issue on it).  Tickets interleave the
// pictures of the wave and run in raster order inside a picture; an MB waits only for those of its four neighbours (left,
// top-left, top, top-right) that are intra MBs themselves -- inter neighbours were reconstructed by recon_inter2_kernel --
// i.e. only for MBs of earlier tickets, which some running warp already holds: residency order cannot deadlock.
// Completion is an epoch stamp per MB (no clearing between launches).
#ifndef H264R_SPARSE_WARPS
#define H264R_SPARSE_WARPS 4
#endif
#ifndef H264R_SPARSE_CTAS
#define H264R_SPARSE_CTAS (48 / H264R_SPARSE_WARPS)
#endif
constexpr int kSparseWarps = H264R_SPARSE_WARPS;
__global__ void __launch_bounds__(kSparseWarps * 32, H264R_SPARSE_CTAS)
recon_intra_sparse_kernel(const DevPicture* __restrict__ pics, int num_pics, int* tickets, const uint32_t* __restrict__ wave_max,
                          FrameGeom g, uint32_t epoch)
{
    __shared__ __align__(16) IntraSmem smem_all[kSparseWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = g.width_mbs, nmb = W * g.height_mbs;
    const int most = (int)__ldg(wave_max);                 // entries of the picture with the most intra MBs
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(&tickets[2], 1);
    ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
#pragma unroll 1
    for (;;) {
        const int entry = ticket / num_pics, pic_i = ticket - entry * num_pics;
        if (entry >= most) return;
        int next = 0;
        if (lane == 0) next = atomicAdd(&tickets[2], 1);
        const DevPicture& pic = pics[pic_i];
        int addr = nmb;
        if (!pic.all_intra && entry < (int)__ldg(pic.intra_count)) addr = (int)__ldg(pic.intra_list + entry);
        if (addr < nmb) {
            const int mby = addr / W, mbx = addr - mby * W;
            IntraPre pre;
            intra_prefetch(pic, g, mbx, mby, lane, pre);       // header, neighbour headers, residual: all in flight at once
            if (lane < 4 && pre.nbw != 0xFFFFFFFFu && ((pre.nbw >> 8) & H264R_MB_FLAG_INTRA)) {
                const int nx = mbx + (lane == 3 ? 1 : (lane == 1 ? 0 : -1)), ny = mby - (lane == 0 ? 0 : 1);   // left, top, top-left, top-right
                const int* flag = reinterpret_cast<const int*>(pic.mb_done + ny * W + nx);
                unsigned ns = 16;
                while ((uint32_t)ld_acquire(flag) != epoch) { __nanosleep(ns); if (ns < 256) ns *= 2; }
            }
            __syncwarp();
            intra_reconstruct_mb<false>(pic, g, smem_all[warp], pre, mbx, mby, lane);
            __syncwarp();
            if (lane == 0) st_release(reinterpret_cast<int*>(pic.mb_done + addr), (int)epoch);
        }
        ticket = __shfl_sync(0xFFFFFFFFu, next, 0);
    }
}

} // namespace h264r
#endif
