// H.264 (ITU-T Rec. H.264 | ISO/IEC 14496-10) specification tables used on the host side of the
// reconstruction boundary: default scaling lists (Tables 7-3, 7-4), normAdjust (8.5.9), zig-zag scans
// (Table 8-13 / Figure 8-8), QPc (Table 8-15).  The reference holds the same spec tables at
// decoder/transform.cc:54-170, 307-341 and parser/interpret_mb.cc:777-782.
#ifndef H264_TABLES_H_
#define H264_TABLES_H_

#include <stdint.h>

namespace h264r {

// Table 7-3 (raster order)
static const int kDefault4x4Intra[16] = { 6, 13, 20, 28, 13, 20, 28, 32, 20, 28, 32, 37, 28, 32, 37, 42 };
static const int kDefault4x4Inter[16] = { 10, 14, 20, 24, 14, 20, 24, 27, 20, 24, 27, 30, 24, 27, 30, 34 };
// Table 7-4 (raster order); both are symmetric Toeplitz-like: value depends on anti-diagonal, listed per row
static const int kDefault8x8Intra[64] = {
     6, 10, 13, 16, 18, 23, 25, 27,   10, 11, 16, 18, 23, 25, 27, 29,
    13, 16, 18, 23, 25, 27, 29, 31,   16, 18, 23, 25, 27, 29, 31, 33,
    18, 23, 25, 27, 29, 31, 33, 36,   23, 25, 27, 29, 31, 33, 36, 38,
    25, 27, 29, 31, 33, 36, 38, 40,   27, 29, 31, 33, 36, 38, 40, 42 };
static const int kDefault8x8Inter[64] = {
     9, 13, 15, 17, 19, 21, 22, 24,   13, 13, 17, 19, 21, 22, 24, 25,
    15, 17, 19, 21, 22, 24, 25, 27,   17, 19, 21, 22, 24, 25, 27, 28,
    19, 21, 22, 24, 25, 27, 28, 30,   21, 22, 24, 25, 27, 28, 30, 32,
    22, 24, 25, 27, 28, 30, 32, 33,   24, 25, 27, 28, 30, 32, 33, 35 };
static const int kFlat16[64] = {
    16,16,16,16,16,16,16,16, 16,16,16,16,16,16,16,16, 16,16,16,16,16,16,16,16, 16,16,16,16,16,16,16,16,
    16,16,16,16,16,16,16,16, 16,16,16,16,16,16,16,16, 16,16,16,16,16,16,16,16, 16,16,16,16,16,16,16,16 };

// 8.5.9: normAdjust4x4(m, i, j) and normAdjust8x8(m, i, j)
inline int norm_adjust_4x4(int m, int i, int j)
{
    static const int v[6][3] = { {10,16,13}, {11,18,14}, {13,20,16}, {14,23,18}, {16,25,20}, {18,29,23} };
    if ((i & 1) == 0 && (j & 1) == 0) return v[m][0];
    if ((i & 1) == 1 && (j & 1) == 1) return v[m][1];
    return v[m][2];
}
inline int norm_adjust_8x8(int m, int i, int j)
{
    static const int v[6][6] = { {20,18,32,19,25,24}, {22,19,35,21,28,26}, {26,23,42,24,33,31},
                                 {28,25,45,26,35,33}, {32,28,51,30,40,38}, {36,32,58,34,46,43} };
    if ((i & 3) == 0 && (j & 3) == 0) return v[m][0];
    if ((i & 1) == 1 && (j & 1) == 1) return v[m][1];
    if ((i & 3) == 2 && (j & 3) == 2) return v[m][2];
    if (((i & 3) == 0 && (j & 1) == 1) || ((i & 1) == 1 && (j & 3) == 0)) return v[m][3];
    if (((i & 3) == 0 && (j & 3) == 2) || ((i & 3) == 2 && (j & 3) == 0)) return v[m][4];
    return v[m][5];
}

// Table 8-15: QPc as a function of qPi
inline int qpc_from_qpi(int qpi)
{
    static const uint8_t t[22] = { 29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39 };
    return qpi < 30 ? qpi : t[qpi - 30];
}

// Frame zig-zag scans, generated: idx -> (x, y).  4x4: Figure 8-8a; 8x8: Figure 8-8 (8x8 zig-zag).
// Field scans (field pictures, transform.cc:344-382): H.264 Table 8-13 (4x4: down the first column, then column by column)
// and Table 8-14 (8x8), one hex digit per index.
struct ZigZag {
    uint8_t x4[16], y4[16], x8[64], y8[64];
    uint8_t fx4[16], fy4[16], fx8[64], fy8[64];
    ZigZag()
    {
        gen(4, x4, y4);
        gen(8, x8, y8);
        digits("0010011122223333", fx4); digits("0102312301230123", fy4);
        digits("0001100121000123211123432222345433334565444456655556776666777777", fx8);
        digits("0120134203567410256731024567310245673102456731245673014567234567", fy8);
    }
    static void digits(const char* d, uint8_t* out) { for (int k = 0; d[k]; ++k) out[k] = (uint8_t)(d[k] - '0'); }
    const uint8_t* sx4(bool field) const { return field ? fx4 : x4; }
    const uint8_t* sy4(bool field) const { return field ? fy4 : y4; }
    const uint8_t* sx8(bool field) const { return field ? fx8 : x8; }
    const uint8_t* sy8(bool field) const { return field ? fy8 : y8; }
    static void gen(int n, uint8_t* xs, uint8_t* ys)
    {
        int x = 0, y = 0;
        for (int k = 0; k < n * n; ++k) {
            xs[k] = (uint8_t)x; ys[k] = (uint8_t)y;
            if (((x + y) & 1) == 0) {            // moving up-right
                if (x == n - 1) ++y;
                else if (y == 0) ++x;
                else { ++x; --y; }
            } else {                              // moving down-left
                if (y == n - 1) ++x;
                else if (x == 0) ++y;
                else { --x; ++y; }
            }
        }
    }
};

} // namespace h264r
#endif
