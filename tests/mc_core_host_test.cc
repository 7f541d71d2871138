// Host check of arrow-h264_b200/csrc/mc_core.cuh (the packed-arithmetic core of recon_inter_kernel) with the PTX
// primitives emulated: every quarter-pel / eighth-pel case, every window alignment, extreme sample patterns and the
// four weighting modes, against a plain per-sample restatement of the reference's formulas
// (decoder/inter_prediction.cc:53-156, 158-406 -- the same restatement oracle/port_recon.c is built from).
#define H264R_HOST_EMUL 1
#include "mc_core.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

using namespace h264r;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd()
{
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 16);
}
static int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
static int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }

// reference luma sample at patch position (x, y); W(x, y) reads the window relative to the patch origin
template <typename WF>
static int ref_luma(WF W, int x, int y, int xf, int yf)
{
    if ((xf | yf) == 0) return W(x, y);
    if (yf == 0) {
        int b = clip255((tap6(W(x - 2, y), W(x - 1, y), W(x, y), W(x + 1, y), W(x + 2, y), W(x + 3, y)) + 16) >> 5);
        return xf == 2 ? b : (W(x + (xf == 3), y) + b + 1) >> 1;
    }
    if (xf == 0) {
        int h = clip255((tap6(W(x, y - 2), W(x, y - 1), W(x, y), W(x, y + 1), W(x, y + 2), W(x, y + 3)) + 16) >> 5);
        return yf == 2 ? h : (W(x, y + (yf == 3)) + h + 1) >> 1;
    }
    if ((xf & 1) && (yf & 1)) {
        const int yy = y + (yf == 3), xx = x + (xf == 3);
        int b = clip255((tap6(W(x - 2, yy), W(x - 1, yy), W(x, yy), W(x + 1, yy), W(x + 2, yy), W(x + 3, yy)) + 16) >> 5);
        int h = clip255((tap6(W(xx, y - 2), W(xx, y - 1), W(xx, y), W(xx, y + 1), W(xx, y + 2), W(xx, y + 3)) + 16) >> 5);
        return (b + h + 1) >> 1;
    }
    int t[6];
    for (int r = 0; r < 6; ++r) {
        const int yy = y - 2 + r;
        t[r] = tap6(W(x - 2, yy), W(x - 1, yy), W(x, yy), W(x + 1, yy), W(x + 2, yy), W(x + 3, yy));
    }
    const int j = clip255((tap6(t[0], t[1], t[2], t[3], t[4], t[5]) + 512) >> 10);
    if (xf == 2 && yf != 2) return (j + clip255((t[yf == 3 ? 3 : 2] + 16) >> 5) + 1) >> 1;
    if (yf == 2 && xf != 2) {
        const int xx = x + (xf == 3);
        int h = clip255((tap6(W(xx, y - 2), W(xx, y - 1), W(xx, y), W(xx, y + 1), W(xx, y + 2), W(xx, y + 3)) + 16) >> 5);
        return (j + h + 1) >> 1;
    }
    return j;
}

static int fails = 0;

static void fill(uint8_t* p, int n, int pattern)
{
    for (int i = 0; i < n; ++i) {
        switch (pattern) {
        case 0: p[i] = (uint8_t)rnd(); break;
        case 1: p[i] = 255; break;
        case 2: p[i] = 0; break;
        case 3: p[i] = (rnd() & 1) ? 255 : 0; break;
        case 4: p[i] = (i & 1) ? 255 : 0; break;
        default: p[i] = (uint8_t)(128 + (int)(rnd() % 7) - 3); break;
        }
    }
}

static void test_luma()
{
    alignas(16) uint8_t win[16 * 16 + 16];
    for (int iter = 0; iter < 4000; ++iter) {
        const int pattern = iter % 6;
        fill(win, sizeof(win), pattern);
        {
            const int pitch = kLumaPitchWords * 4;
            const int max_off = 2 + 4 + 3;                                  // uniform quadrant, second block, x0 & 3 == 3
            for (int off = 2; off <= max_off; ++off)
                for (int xf = 0; xf < 4; ++xf)
                    for (int yf = 0; yf < 4; ++yf) {
                        uint32_t o0, o1;
                        unsigned hm, cm;
                        mc_luma_masks(xf, yf, hm, cm);
                        // own masks, then every stage on (what a warp with lanes of all sixteen positions passes)
                        mc_luma_patch_4x2(reinterpret_cast<const uint32_t*>(win), off, xf, yf, hm, cm, o0, o1);
                        uint32_t a0, a1;
                        mc_luma_patch_4x2(reinterpret_cast<const uint32_t*>(win), off, xf, yf, 0x7Fu, 0x7Fu, a0, a1);
                        if ((a0 != o0 || a1 != o1) && fails++ < 20) printf("warp masks change the result at xf=%d yf=%d off=%d\n", xf, yf, off);
                        auto W = [&](int x, int y) { return (int)win[(y + 2) * pitch + off + x]; };
                        {   // the 4x4 patch of the same position (two MBs per warp kernel)
                            unsigned hm4, cm4; uint32_t o4[4], b4[4];
                            mc_luma_masks_r<4>(xf, yf, hm4, cm4);
                            mc_luma_patch<4>(reinterpret_cast<const uint32_t*>(win), off, xf, yf, hm4, cm4, o4);
                            mc_luma_patch<4>(reinterpret_cast<const uint32_t*>(win), off, xf, yf, 0x1FFu, 0x1FFu, b4);
                            for (int y = 0; y < 4; ++y) {
                                if (o4[y] != b4[y] && fails++ < 20) printf("4x4: warp masks change the result at xf=%d yf=%d off=%d\n", xf, yf, off);
                                for (int x = 0; x < 4; ++x) {
                                    const int want = ref_luma(W, x, y, xf, yf);
                                    const int got = (int)((o4[y] >> (8 * x)) & 0xFF);
                                    if (want != got && fails++ < 20) printf("4x4 luma off=%d xf=%d yf=%d (%d,%d): want %d got %d\n", off, xf, yf, x, y, want, got);
                                }
                            }
                        }
                        for (int y = 0; y < 2; ++y)
                            for (int x = 0; x < 4; ++x) {
                                const int want = ref_luma(W, x, y, xf, yf);
                                const int got = (int)(((y ? o1 : o0) >> (8 * x)) & 0xFF);
                                if (want != got && fails++ < 20)
                                    printf("luma mismatch pattern %d pitch %d off %d frac (%d,%d) at (%d,%d): want %d got %d\n",
                                           pattern, pitch, off, xf, yf, x, y, want, got);
                            }
                    }
        }
    }
}

static void test_chroma()
{
    alignas(16) uint8_t win[8 * 6 + 8];
    for (int iter = 0; iter < 3000; ++iter) {
        fill(win, sizeof(win), iter % 6);
        for (int off = 0; off <= 5; ++off)
            for (int xf = 0; xf < 8; ++xf)
                for (int yf = 0; yf < 8; ++yf) {
                    const uint32_t got = mc_chroma_patch_2x2(reinterpret_cast<const uint32_t*>(win), off, xf, yf);
                    for (int y = 0; y < 2; ++y)
                        for (int x = 0; x < 2; ++x) {
                            auto W = [&](int xx, int yy) { return (int)win[yy * 8 + off + xx]; };
                            const int want = ((8 - xf) * (8 - yf) * W(x, y) + xf * (8 - yf) * W(x + 1, y) +
                                              (8 - xf) * yf * W(x, y + 1) + xf * yf * W(x + 1, y + 1) + 32) >> 6;
                            const int g = (int)((got >> ((y * 2 + x) * 8)) & 0xFF);
                            if (want != g && fails++ < 20)
                                printf("chroma mismatch off %d frac (%d,%d) at (%d,%d): want %d got %d\n", off, xf, yf, x, y, want, g);
                        }
                }
    }
}

static int rshift_rnd(int x, int a) { return a > 0 ? (x + (1 << (a - 1))) >> a : x; }

static void test_weight()
{
    for (int iter = 0; iter < 400000; ++iter) {
        const int mode = (int)(rnd() & 3);
        uint32_t p0 = rnd() ^ (rnd() << 16), p1 = rnd() ^ (rnd() << 16);
        if (iter % 7 == 0) p0 = 0xFFFFFFFFu;
        if (iter % 11 == 0) p1 = 0;
        const int extreme = iter % 5 == 0;
        int w0 = extreme ? ((rnd() & 1) ? 127 : -128) : (int)(rnd() % 256) - 128;
        int w1 = extreme ? ((rnd() & 1) ? 128 : -64) : (int)(rnd() % 193) - 64;     // implicit weights reach 128
        if (mode == 3 && (rnd() & 1)) w0 = 64 - w1;
        const int d = (int)(rnd() % 8);
        const int o = (int)(rnd() % 256) - 128;
        int16_t res[4];
        for (int i = 0; i < 4; ++i) res[i] = (int16_t)((int)(rnd() % 511) - 255);
        const uint32_t res01 = (uint32_t)(uint16_t)res[0] | (uint32_t)(uint16_t)res[1] << 16;
        const uint32_t res23 = (uint32_t)(uint16_t)res[2] | (uint32_t)(uint16_t)res[3] << 16;
        const uint32_t got = mc_weight_recon4(mode, p0, p1, w0, w1, d, o, res01, res23);
        for (int i = 0; i < 4; ++i) {
            const int s0 = (p0 >> (8 * i)) & 0xFF, s1 = (p1 >> (8 * i)) & 0xFF;
            int v;
            if (mode == 0) v = s0;
            else if (mode == 1) v = clip255(rshift_rnd(w0 * s0, d) + o);
            else if (mode == 2) v = (s0 + s1 + 1) >> 1;
            else v = clip255(rshift_rnd(w0 * s0 + w1 * s1, d + 1) + o);
            v = clip255(v + res[i]);
            const int g = (int)((got >> (8 * i)) & 0xFF);
            if (v != g && fails++ < 20)
                printf("weight mismatch mode %d w (%d,%d) d %d o %d s (%d,%d) res %d: want %d got %d\n", mode, w0, w1, d, o, s0, s1, res[i], v, g);
        }
    }
}

int main()
{
    test_luma();
    test_chroma();
    test_weight();
    if (fails) { printf("FAILED: %d mismatches\n", fails); return 1; }
    printf("mc_core ok\n");
    return 0;
}
