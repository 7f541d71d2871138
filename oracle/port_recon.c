/*
 * TEST INFRASTRUCTURE ONLY -- the oracle.  Nothing under arrow-h264_b200/ may include, link or call this.
 *
 * Plain-C restatement of the reference's macroblock reconstruction path (luuvish/arrow-h264,
 * src/codec/h264/decoder/) operating directly on the neutral picture description of include/h264recon.h.
 * Every function cites the reference lines it follows.  Scope: 8-bit 4:2:0 frame pictures, no MBAFF/PAFF,
 * no transform bypass, no SP/SI.
 *
 * Parity status: PINNED BY EXECUTION.  The reference ships no golden vectors for this path, so
 * this file is pinned against the reference's own Decoder compiled unmodified from /root/reference
 * (oracle/_ref/libh264ref.so, driven by oracle/ref_harness.cc): tests/test_oracle_vs_reference.py compares
 * every sample of every picture of all BASELINE configs (reduced sizes), and the tests/golden json files holds the
 * per-frame MD5 digests the reference produced, which this file must reproduce where /root/reference is
 * absent (the GPU box).
 */
#define _POSIX_C_SOURCE 199309L
#include "port_recon.h"

#include <stdlib.h>
#include <string.h>
#include <time.h>

#define MAX_FRAMES 64

typedef struct {
    uint8_t* pl[3];                /* Y, Cb, Cr; tight pitch */
    int used;
} frame_t;

struct port_dec {
    int W, H;                      /* in MBs */
    int direct8x8;
    frame_t fr[MAX_FRAMES];
};

typedef struct {
    port_dec* d;
    const h264r_pic_params* pp;
    const h264r_slice* slices;
    const h264r_mb* mbs;
    const h264r_mb_motion* motion;
    const h264r_level* levels;
    frame_t* dst;
} pic_t;

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int clip1(int v) { return clip3(0, 255, v); }
static inline int iabs(int v) { return v < 0 ? -v : v; }

/* ---- frames ------------------------------------------------------------------------------------- */

port_dec* port_open(const h264r_seq_params* sp)
{
    port_dec* d = (port_dec*)calloc(1, sizeof(*d));
    d->W = sp->width_mbs; d->H = sp->height_mbs; d->direct8x8 = sp->direct_8x8_inference_flag;
    return d;
}
void port_close(port_dec* d)
{
    for (int i = 0; i < MAX_FRAMES; ++i) if (d->fr[i].used) free(d->fr[i].pl[0]);
    free(d);
}
int port_frame_alloc(port_dec* d)
{
    for (int i = 0; i < MAX_FRAMES; ++i)
        if (!d->fr[i].used) {
            size_t ny = (size_t)d->W * 16 * d->H * 16, nc = ny / 4;
            uint8_t* p = (uint8_t*)calloc(ny + 2 * nc, 1);
            d->fr[i].pl[0] = p; d->fr[i].pl[1] = p + ny; d->fr[i].pl[2] = p + ny + nc; d->fr[i].used = 1;
            return i;
        }
    return -1;
}
void port_frame_release(port_dec* d, int id)
{
    if (id >= 0 && id < MAX_FRAMES && d->fr[id].used) { free(d->fr[id].pl[0]); d->fr[id].used = 0; }
}
void port_frame_get(port_dec* d, int id, uint8_t* y, uint8_t* cb, uint8_t* cr)
{
    size_t ny = (size_t)d->W * 16 * d->H * 16, nc = ny / 4;
    memcpy(y, d->fr[id].pl[0], ny); memcpy(cb, d->fr[id].pl[1], nc); memcpy(cr, d->fr[id].pl[2], nc);
}
void port_frame_set(port_dec* d, int id, const uint8_t* y, const uint8_t* cb, const uint8_t* cr)
{
    size_t ny = (size_t)d->W * 16 * d->H * 16, nc = ny / 4;
    memcpy(d->fr[id].pl[0], y, ny); memcpy(d->fr[id].pl[1], cb, nc); memcpy(d->fr[id].pl[2], cr, nc);
}

/* ---- residual: dequant + Hadamard + IDCT ---------------------------------------------------------- */

/* transform.cc:597-641 inverse_4x4: rows then columns, (x + 32) >> 6.  d, r: 4x4 blocks with row stride s */
static void idct4x4(const int* d, int* r, int s)
{
    int f[4][4];
    for (int i = 0; i < 4; ++i) {
        int d0 = d[i * s], d1 = d[i * s + 1], d2 = d[i * s + 2], d3 = d[i * s + 3];
        int e0 = d0 + d2, e1 = d0 - d2, e2 = (d1 >> 1) - d3, e3 = d1 + (d3 >> 1);
        f[i][0] = e0 + e3; f[i][1] = e1 + e2; f[i][2] = e1 - e2; f[i][3] = e0 - e3;
    }
    for (int j = 0; j < 4; ++j) {
        int f0 = f[0][j], f1 = f[1][j], f2 = f[2][j], f3 = f[3][j];
        int g0 = f0 + f2, g1 = f0 - f2, g2 = (f1 >> 1) - f3, g3 = f1 + (f3 >> 1);
        r[0 * s + j] = (g0 + g3 + 32) >> 6;
        r[1 * s + j] = (g1 + g2 + 32) >> 6;
        r[2 * s + j] = (g1 - g2 + 32) >> 6;
        r[3 * s + j] = (g0 - g3 + 32) >> 6;
    }
}

/* transform.cc:643-733 inverse_8x8: one 8-point butterfly, applied to rows then columns */
static void idct8_1d(const int* in, int istride, int* out, int ostride, int round_shift)
{
    int d0 = in[0], d1 = in[istride], d2 = in[2 * istride], d3 = in[3 * istride];
    int d4 = in[4 * istride], d5 = in[5 * istride], d6 = in[6 * istride], d7 = in[7 * istride];
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
    int o[8] = { f0 + f7, f2 + f5, f4 + f3, f6 + f1, f6 - f1, f4 - f3, f2 - f5, f0 - f7 };
    for (int k = 0; k < 8; ++k) out[k * ostride] = round_shift ? (o[k] + 32) >> 6 : o[k];
}
static void idct8x8(const int* d, int* r, int s)
{
    int g[64];
    for (int i = 0; i < 8; ++i) idct8_1d(d + i * s, 1, g + i * 8, 1, 0);
    for (int j = 0; j < 8; ++j) idct8_1d(g + j, 8, r + j, s, 1);
}

/* Computes the residual of one MB.  res[0] = luma 16x16 (stride 16), res[1], res[2] = chroma 8x8 (stride 8).
 * has[0] = bitmask over the four luma 8x8 blocks that carry a residual ("add" vs "copy pred",
 * transform.cc:926-934, 1018-1031, 1070-1073); has[1] = chroma residual present (transform.cc:1033-1049, 1085-1090). */
static void mb_residual(const pic_t* p, const h264r_mb* mb, int res[3][256], int has[2])
{
    const h264r_slice* sl = &p->slices[mb->slice_idx];
    const int intra = (mb->flags & H264R_MB_FLAG_INTRA) != 0;
    const int t8 = (mb->flags & H264R_MB_FLAG_T8x8) != 0;
    /* the MB's transmitted levels, placed at their raster positions (what Transform::cof holds after the
       residual parser ran, transform.cc:425-456, before dequantisation) */
    int16_t dense[H264R_COEFFS_PER_MB];
    const int16_t* lev = NULL;
    if (mb->coeff_count) {
        memset(dense, 0, sizeof(dense));
        for (int i = 0; i < mb->coeff_count; ++i) {
            h264r_level e = p->levels[mb->coeff_offset + i];
            dense[H264R_LEVEL_POS(e)] = (int16_t)H264R_LEVEL_VALUE(e);
        }
        lev = dense;
    }
    int cof[256];
    memset(res[0], 0, sizeof(int) * 256); memset(res[1], 0, sizeof(int) * 256); memset(res[2], 0, sizeof(int) * 256);
    has[0] = has[1] = 0;

    /* ---- luma ---- */
    const int qp = mb->qp_y, per = qp / 6, rem = qp % 6;
    const uint16_t* ls4 = sl->level_scale_4x4[intra ? 0 : 1][0][rem];
    if (mb->mb_type == H264R_MB_I16x16) {
        memset(cof, 0, sizeof(cof));
        if (lev) {
            /* AC: inverse_quantize at coeff_luma_ac time (transform.cc:394-422, 431-440): ((l*LS) << per + 8) >> 4 */
            for (int y = 0; y < 16; ++y)
                for (int x = 0; x < 16; ++x) {
                    if (((x | y) & 3) == 0) continue;
                    int l = lev[y * 16 + x];
                    if (l) cof[y * 16 + x] = ((l * ls4[(y & 3) * 4 + (x & 3)]) * (1 << per) + 8) >> 4;
                }
        }
        /* transform_luma_dc (transform.cc:825-856) + ihadamard_4x4 (transform.cc:515-554) */
        int c[4][4], e[4][4], f[4][4];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = lev ? lev[i * 4 * 16 + j * 4] : 0;
        for (int i = 0; i < 4; ++i) {
            int a0 = c[i][0] + c[i][2], a1 = c[i][0] - c[i][2], a2 = c[i][1] - c[i][3], a3 = c[i][1] + c[i][3];
            e[i][0] = a0 + a3; e[i][1] = a1 + a2; e[i][2] = a1 - a2; e[i][3] = a0 - a3;
        }
        for (int j = 0; j < 4; ++j) {
            int a0 = e[0][j] + e[2][j], a1 = e[0][j] - e[2][j], a2 = e[1][j] - e[3][j], a3 = e[1][j] + e[3][j];
            f[0][j] = a0 + a3; f[1][j] = a1 + a2; f[2][j] = a1 - a2; f[3][j] = a0 - a3;
        }
        int scale = sl->level_scale_4x4[0][0][rem][0];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                cof[i * 4 * 16 + j * 4] = qp >= 36 ? (f[i][j] * scale) * (1 << (per - 6))
                                                   : (f[i][j] * scale + (1 << (5 - per))) >> (6 - per);
        /* inverse_transform_16x16 (transform.cc:1018-1031): all 16 blocks, always added */
        for (int by = 0; by < 4; ++by) for (int bx = 0; bx < 4; ++bx)
            idct4x4(cof + by * 4 * 16 + bx * 4, res[0] + by * 4 * 16 + bx * 4, 16);
        has[0] = 15;
    } else if (mb->mb_type != H264R_MB_IPCM) {
        for (int b8 = 0; b8 < 4; ++b8) {
            if (!(mb->cbp_luma & (1 << b8)) || !lev) continue;     /* quirk 6: levels ignored without the CBP bit */
            int x0 = (b8 & 1) * 8, y0 = (b8 >> 1) * 8;
            memset(cof, 0, sizeof(cof));
            if (t8) {
                const uint16_t* ls8 = sl->level_scale_8x8[intra ? 0 : 1][rem];
                for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) {
                    int l = lev[(y0 + y) * 16 + x0 + x];
                    if (l) cof[(y0 + y) * 16 + x0 + x] = ((l * ls8[y * 8 + x]) * (1 << per) + 32) >> 6;
                }
                idct8x8(cof + y0 * 16 + x0, res[0] + y0 * 16 + x0, 16);
            } else {
                for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) {
                    int l = lev[(y0 + y) * 16 + x0 + x];
                    if (l) cof[(y0 + y) * 16 + x0 + x] = ((l * ls4[(y & 3) * 4 + (x & 3)]) * (1 << per) + 8) >> 4;
                }
                for (int k = 0; k < 4; ++k) {
                    int off = (y0 + (k >> 1) * 4) * 16 + x0 + (k & 1) * 4;
                    idct4x4(cof + off, res[0] + off, 16);
                }
            }
            has[0] |= 1 << b8;
        }
    }

    /* ---- chroma (transform.cc:858-910 transform_chroma_dc, 1033-1049 inverse_transform_chroma) ---- */
    if (mb->mb_type != H264R_MB_IPCM && (intra || mb->cbp_chroma)) {
        has[1] = 1;
        for (int pl = 1; pl <= 2; ++pl) {
            if (!mb->cbp_chroma || !lev) continue;                 /* all-zero cof -> zero residual */
            const int qc = mb->qp_c[pl - 1], cper = qc / 6, crem = qc % 6;
            const uint16_t* lsc = sl->level_scale_4x4[intra ? 0 : 1][pl][crem];
            const int16_t* cl = lev + 256 + (pl - 1) * 64;
            int cc[64];
            memset(cc, 0, sizeof(cc));
            for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) {
                if (((x | y) & 3) == 0) continue;
                int l = cl[y * 8 + x];
                if (l) cc[y * 8 + x] = ((l * lsc[(y & 3) * 4 + (x & 3)]) * (1 << cper) + 8) >> 4;
            }
            /* ihadamard_2x2 (transform.cc:460-481), then ((f*scale) << per) >> 5 */
            int c00 = cl[0], c01 = cl[4], c10 = cl[32], c11 = cl[36];
            int e00 = c00 + c01, e01 = c00 - c01, e10 = c10 + c11, e11 = c10 - c11;
            int f[4] = { e00 + e10, e01 + e11, e00 - e10, e01 - e11 };
            int scale = lsc[0];
            cc[0]  = ((f[0] * scale) * (1 << cper)) >> 5;
            cc[4]  = ((f[1] * scale) * (1 << cper)) >> 5;
            cc[32] = ((f[2] * scale) * (1 << cper)) >> 5;
            cc[36] = ((f[3] * scale) * (1 << cper)) >> 5;
            for (int k = 0; k < 4; ++k) {
                int off = (k >> 1) * 4 * 8 + (k & 1) * 4;
                idct4x4(cc + off, res[pl] + off, 8);
            }
        }
    }
}

/* ---- neighbour availability (parser/neighbour.cc:123-175 + slice_nr test, quirk 7) ----------------- */

static int mb_avail(const pic_t* p, int cur, int nx, int ny, int need_intra)
{
    const int W = p->d->W, H = p->d->H;
    if (nx < 0 || nx >= W || ny < 0 || ny >= H) return 0;
    int nb = ny * W + nx;
    if (nb >= cur) return 0;                                   /* not decoded yet: slice_nr == -1 */
    if (p->mbs[nb].slice_idx != p->mbs[cur].slice_idx) return 0;
    if (need_intra && !(p->mbs[nb].flags & H264R_MB_FLAG_INTRA)) return 0;
    return 1;
}

/* ---- intra prediction ----------------------------------------------------------------------------- */

/* 9 directional modes shared by Intra4x4 (intra_prediction.cc:189-346) and Intra8x8 (:449-606).
 * t[-1..2n-1] = top row incl. corner at t[-1], l[-1..n-1] = left column incl. corner at l[-1]. */
static void pred_directional(int mode, int n, const int* t, const int* l, int dcval, uint8_t* pred, int stride)
{
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x) {
            int v;
            switch (mode) {
            case 0: v = t[x]; break;
            case 1: v = l[y]; break;
            case 2: v = dcval; break;
            case 3:
                if (x == n - 1 && y == n - 1) v = (t[x + y] + 3 * t[x + y + 1] + 2) >> 2;
                else v = (t[x + y] + 2 * t[x + y + 1] + t[x + y + 2] + 2) >> 2;
                break;
            case 4:
                if (x > y) v = (t[x - y - 2] + 2 * t[x - y - 1] + t[x - y] + 2) >> 2;
                else if (x < y) v = (l[y - x - 2] + 2 * l[y - x - 1] + l[y - x] + 2) >> 2;
                else v = (t[0] + 2 * t[-1] + l[0] + 2) >> 2;
                break;
            case 5: {
                int z = 2 * x - y;
                if (z >= 0 && (z & 1) == 0) v = (t[x - (y >> 1) - 1] + t[x - (y >> 1)] + 1) >> 1;
                else if (z >= 0) v = (t[x - (y >> 1) - 2] + 2 * t[x - (y >> 1) - 1] + t[x - (y >> 1)] + 2) >> 2;
                else if (z == -1) v = (l[0] + 2 * t[-1] + t[0] + 2) >> 2;
                else v = (l[y - 2 * x - 1] + 2 * l[y - 2 * x - 2] + l[y - 2 * x - 3] + 2) >> 2;
                break; }
            case 6: {
                int z = 2 * y - x;
                if (z >= 0 && (z & 1) == 0) v = (l[y - (x >> 1) - 1] + l[y - (x >> 1)] + 1) >> 1;
                else if (z >= 0) v = (l[y - (x >> 1) - 2] + 2 * l[y - (x >> 1) - 1] + l[y - (x >> 1)] + 2) >> 2;
                else if (z == -1) v = (l[0] + 2 * t[-1] + t[0] + 2) >> 2;
                else v = (t[x - 2 * y - 1] + 2 * t[x - 2 * y - 2] + t[x - 2 * y - 3] + 2) >> 2;
                break; }
            case 7:
                if ((y & 1) == 0) v = (t[x + (y >> 1)] + t[x + (y >> 1) + 1] + 1) >> 1;
                else v = (t[x + (y >> 1)] + 2 * t[x + (y >> 1) + 1] + t[x + (y >> 1) + 2] + 2) >> 2;
                break;
            default: {
                int z = x + 2 * y, m = 2 * n - 3;
                if (z < m && (z & 1) == 0) v = (l[y + (x >> 1)] + l[y + (x >> 1) + 1] + 1) >> 1;
                else if (z < m) v = (l[y + (x >> 1)] + 2 * l[y + (x >> 1) + 1] + l[y + (x >> 1) + 2] + 2) >> 2;
                else if (z == m) v = (l[n - 2] + 3 * l[n - 1] + 2) >> 2;
                else v = l[n - 1];
                break; }
            }
            pred[y * stride + x] = (uint8_t)v;
        }
}

static int dc_value(int n, int a, int b, const int* t, const int* l)
{
    /* intra_prediction.cc:211-232, 471-492, 690-711: round/shift depend on which neighbours exist */
    int log2n = n == 4 ? 2 : (n == 8 ? 3 : 4);
    if (!a && !b) return 128;
    int sum = 0;
    if (a) for (int y = 0; y < n; ++y) sum += l[y];
    if (b) for (int x = 0; x < n; ++x) sum += t[x];
    int shift = log2n - 1 + (a ? 1 : 0) + (b ? 1 : 0);
    int round = (a ? n / 2 : 0) + (b ? n / 2 : 0);
    return (sum + round) >> shift;
}

/* availability of A, B, C, D for a luma block of size n at (xO, yO) inside MB `cur`
 * (Intra4x4 ctor intra_prediction.cc:137-168, Intra8x8 ctor :359-390) */
static void luma_block_avail(const pic_t* p, int cur, int n, int xO, int yO, int ci, int av[4])
{
    const int W = p->d->W, mbx = cur % W, mby = cur / W;
    int L = mb_avail(p, cur, mbx - 1, mby, ci), T = mb_avail(p, cur, mbx, mby - 1, ci);
    int TL = mb_avail(p, cur, mbx - 1, mby - 1, ci), TR = mb_avail(p, cur, mbx + 1, mby - 1, ci);
    av[0] = xO > 0 ? 1 : L;
    av[1] = yO > 0 ? 1 : T;
    av[3] = (xO > 0 && yO > 0) ? 1 : (xO > 0 ? T : (yO > 0 ? L : TL));
    if (yO == 0) av[2] = (xO + n < 16) ? T : TR;
    else if (xO + n >= 16) av[2] = 0;                           /* right MB: not decoded yet */
    else av[2] = 1;
    if (n == 4 && xO == 4 && (yO == 4 || yO == 12)) av[2] = 0;  /* :154 */
    if (n == 8 && xO == 8 && yO == 8) av[2] = 0;                /* :376 */
}

static void intra_luma_block(const pic_t* p, int cur, int n, int xO, int yO, int mode, uint8_t* pred /*stride 16*/)
{
    const int W = p->d->W, stride = W * 16;
    const int ci = p->slices[p->mbs[cur].slice_idx].constrained_intra_pred_flag;
    const uint8_t* img = p->dst->pl[0];
    const int px = (cur % W) * 16 + xO, py = (cur / W) * 16 + yO;
    int av[4];
    luma_block_avail(p, cur, n, xO, yO, ci, av);
    int tbuf[18], lbuf[10];                                     /* index -1 .. */
    int* t = tbuf + 1; int* l = lbuf + 1;
    for (int i = -1; i < 2 * n; ++i) t[i] = 0;
    for (int i = -1; i < n; ++i) l[i] = 0;
    if (av[3]) t[-1] = l[-1] = img[(py - 1) * stride + px - 1];
    if (av[0]) for (int y = 0; y < n; ++y) l[y] = img[(py + y) * stride + px - 1];
    if (av[1]) {
        for (int x = 0; x < n; ++x) t[x] = img[(py - 1) * stride + px + x];
        for (int x = n; x < 2 * n; ++x) t[x] = av[2] ? img[(py - 1) * stride + px + x] : t[n - 1];   /* :182-185, 404-407 */
    }
    if (n == 8) {
        /* reference sample filtering, intra_prediction.cc:413-447 */
        int ft[18], fl[10]; int* f_t = ft + 1; int* f_l = fl + 1;
        memcpy(ft, tbuf, sizeof(tbuf)); memcpy(fl, lbuf, sizeof(lbuf));
        if (av[1]) {
            f_t[0] = av[3] ? (t[-1] + 2 * t[0] + t[1] + 2) >> 2 : (3 * t[0] + t[1] + 2) >> 2;
            for (int x = 1; x < 15; ++x) f_t[x] = (t[x - 1] + 2 * t[x] + t[x + 1] + 2) >> 2;
            f_t[15] = (t[14] + 3 * t[15] + 2) >> 2;
        }
        if (av[3]) {
            int c = t[-1], v;
            if (av[0] && av[1]) v = (t[0] + 2 * c + l[0] + 2) >> 2;
            else if (av[1]) v = (3 * c + t[0] + 2) >> 2;
            else if (av[0]) v = (3 * c + l[0] + 2) >> 2;
            else v = c;
            f_t[-1] = f_l[-1] = v;
        }
        if (av[0]) {
            f_l[0] = av[3] ? (l[-1] + 2 * l[0] + l[1] + 2) >> 2 : (3 * l[0] + l[1] + 2) >> 2;
            for (int y = 1; y < 7; ++y) f_l[y] = (l[y - 1] + 2 * l[y] + l[y + 1] + 2) >> 2;
            f_l[7] = (l[6] + 3 * l[7] + 2) >> 2;
        }
        memcpy(tbuf, ft, sizeof(tbuf)); memcpy(lbuf, fl, sizeof(lbuf));
    }
    int dcv = mode == 2 ? dc_value(n, av[0], av[1], t, l) : 0;
    pred_directional(mode, n, t, l, dcv, pred, 16);
}

/* Intra16x16 (intra_prediction.cc:624-745) and Chroma (:748-904); n = 16 (luma) or 8 (chroma plane pl) */
static void intra_mb_plane(const pic_t* p, int cur, int pl, int mode, uint8_t* pred /*stride 16*/)
{
    const int n = pl ? 8 : 16, W = p->d->W, stride = W * n;
    const int ci = p->slices[p->mbs[cur].slice_idx].constrained_intra_pred_flag;
    const uint8_t* img = p->dst->pl[pl];
    const int mbx = cur % W, mby = cur / W, px = mbx * n, py = mby * n;
    int A = mb_avail(p, cur, mbx - 1, mby, ci), B = mb_avail(p, cur, mbx, mby - 1, ci);
    int D = mb_avail(p, cur, mbx - 1, mby - 1, ci);
    int tb[17], lb[17]; int* t = tb + 1; int* l = lb + 1;
    memset(tb, 0, sizeof(tb)); memset(lb, 0, sizeof(lb));
    if (D) t[-1] = l[-1] = img[(py - 1) * stride + px - 1];
    if (A) for (int y = 0; y < n; ++y) l[y] = img[(py + y) * stride + px - 1];
    if (B) for (int x = 0; x < n; ++x) t[x] = img[(py - 1) * stride + px + x];

    /* map chroma mode numbering (DC=0,H=1,V=2,Plane=3) onto luma's (V=0,H=1,DC=2,Plane=3) */
    int m = pl ? (mode == 0 ? 2 : (mode == 2 ? 0 : mode)) : mode;
    if (m == 0) { for (int y = 0; y < n; ++y) for (int x = 0; x < n; ++x) pred[y * 16 + x] = (uint8_t)t[x]; return; }
    if (m == 1) { for (int y = 0; y < n; ++y) for (int x = 0; x < n; ++x) pred[y * 16 + x] = (uint8_t)l[y]; return; }
    if (m == 3) {
        int h = n / 2, Hs = 0, Vs = 0;
        for (int x = 0; x < h; ++x) Hs += (x + 1) * (t[h + x] - t[h - 2 - x]);
        for (int y = 0; y < h; ++y) Vs += (y + 1) * (l[h + y] - l[h - 2 - y]);
        int a = 16 * (l[n - 1] + t[n - 1]);
        int b = pl ? (34 * Hs + 32) >> 6 : (5 * Hs + 32) >> 6;
        int c = pl ? (34 * Vs + 32) >> 6 : (5 * Vs + 32) >> 6;
        for (int y = 0; y < n; ++y) for (int x = 0; x < n; ++x)
            pred[y * 16 + x] = (uint8_t)clip1((a + b * (x - (h - 1)) + c * (y - (h - 1)) + 16) >> 5);
        return;
    }
    if (!pl) {
        int dcv = dc_value(16, A, B, t, l);
        for (int y = 0; y < 16; ++y) for (int x = 0; x < 16; ++x) pred[y * 16 + x] = (uint8_t)dcv;
        return;
    }
    /* chroma DC per 4x4 block (intra_prediction.cc:798-849) */
    for (int k = 0; k < 4; ++k) {
        int xO = (k & 1) * 4, yO = (k >> 1) * 4, a, b;
        if ((xO == 0 && yO == 0) || (xO > 0 && yO > 0)) { a = A; b = B; }
        else if (xO > 0) { a = B ? 0 : A; b = B; }
        else { a = A; b = A ? 0 : B; }
        int dcv = dc_value(4, a, b, t + xO, l + yO);
        for (int y = 0; y < 4; ++y) for (int x = 0; x < 4; ++x) pred[(yO + y) * 16 + xO + x] = (uint8_t)dcv;
    }
}

/* ---- inter prediction ----------------------------------------------------------------------------- */

static inline int refpx(const uint8_t* img, int w, int h, int x, int y)
{
    return img[clip3(0, h - 1, y) * w + clip3(0, w - 1, x)];
}
static inline int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }

/* one luma sample at integer position (x, y) + fraction (xf, yf): get_block_luma, inter_prediction.cc:158-340;
 * the reference's block pre-clamp + padded planes equal clamping every tap (SURVEY.md §8a derived facts) */
static int luma_sample(const uint8_t* img, int w, int h, int x, int y, int xf, int yf)
{
#define P(dx, dy) refpx(img, w, h, x + (dx), y + (dy))
#define B1(dy) tap6(P(-2, dy), P(-1, dy), P(0, dy), P(1, dy), P(2, dy), P(3, dy))       /* horizontal, unrounded */
#define H1(dx) tap6(P(dx, -2), P(dx, -1), P(dx, 0), P(dx, 1), P(dx, 2), P(dx, 3))       /* vertical, unrounded   */
    if (xf == 0 && yf == 0) return P(0, 0);
    if (yf == 0) {
        int b = clip1((B1(0) + 16) >> 5);
        return xf == 2 ? b : (P(xf == 1 ? 0 : 1, 0) + b + 1) >> 1;
    }
    if (xf == 0) {
        int hh = clip1((H1(0) + 16) >> 5);
        return yf == 2 ? hh : (P(0, yf == 1 ? 0 : 1) + hh + 1) >> 1;
    }
    if ((xf & 1) && (yf & 1)) {
        int b = clip1((B1(yf == 3 ? 1 : 0) + 16) >> 5);
        int hh = clip1((H1(xf == 3 ? 1 : 0) + 16) >> 5);
        return (b + hh + 1) >> 1;
    }
    int j1 = tap6(B1(-2), B1(-1), B1(0), B1(1), B1(2), B1(3));
    int j = clip1((j1 + 512) >> 10);
    if (xf == 2 && yf == 2) return j;
    if (xf == 2) { int q = clip1((B1(yf == 3 ? 1 : 0) + 16) >> 5); return (j + q + 1) >> 1; }
    { int q = clip1((H1(xf == 3 ? 1 : 0) + 16) >> 5); return (j + q + 1) >> 1; }
#undef P
#undef B1
#undef H1
}

/* get_block_chroma, inter_prediction.cc:342-406 */
static int chroma_sample(const uint8_t* img, int w, int h, int x, int y, int xf, int yf)
{
    int A = refpx(img, w, h, x, y), B = refpx(img, w, h, x + 1, y);
    int C = refpx(img, w, h, x, y + 1), D = refpx(img, w, h, x + 1, y + 1);
    return ((8 - xf) * (8 - yf) * A + xf * (8 - yf) * B + (8 - xf) * yf * C + xf * yf * D + 32) >> 6;
}

static inline int rshift_rnd(int x, int a) { return a > 0 ? (x + (1 << (a - 1))) >> a : x; }

static const int BLOCK_STEP[8][2] = { {0,0}, {4,4}, {4,2}, {2,4}, {2,2}, {2,1}, {1,2}, {1,1} };

/* Decoder::mb_pred_inter partition walk (decoder.cc:217-262): for every 4x4 block, the 4x4 block whose
 * motion entry the reference reads (partition origin) and the prediction direction. */
static void partition_map(const pic_t* p, int cur, int origin[16], int dir[16])
{
    const h264r_mb* mb = &p->mbs[cur];
    const h264r_slice* sl = &p->slices[mb->slice_idx];
    const h264r_mb_motion* m = &p->motion[cur];
    const int is_b = sl->slice_type == H264R_B_SLICE;
    int sh0 = BLOCK_STEP[mb->mb_type][0], sv0 = BLOCK_STEP[mb->mb_type][1];
    if (mb->mb_type == 0) sh0 = sv0 = is_b ? 2 : 4;
    for (int j0 = 0; j0 < 4; j0 += sv0)
        for (int i0 = 0; i0 < 4; i0 += sh0) {
            int b8 = 2 * (j0 >> 1) + (i0 >> 1);
            int mode = mb->u.inter.sub_mb_type[b8], pd = mb->u.inter.sub_mb_pred_mode[b8];
            int sh4 = BLOCK_STEP[mode][0], sv4 = BLOCK_STEP[mode][1];
            if (mode == 0) sh4 = sv4 = p->d->direct8x8 ? 2 : 1;
            if (is_b && mb->mb_type == H264R_MB_8x8 && sl->direct_spatial_mv_pred_flag) {
                int b = j0 * 4 + i0;
                pd = m->ref_idx[1][b] < 0 ? 0 : (m->ref_idx[0][b] < 0 ? 1 : 2);
            }
            for (int j = j0; j < j0 + sv0; j += sv4)
                for (int i = i0; i < i0 + sh0; i += sh4)
                    for (int y = j; y < j + sv4; ++y)
                        for (int x = i; x < i + sh4; ++x) { origin[y * 4 + x] = j * 4 + i; dir[y * 4 + x] = pd; }
        }
}

/* inter prediction of the whole MB into pred[3] (stride 16), per 4x4 block (inter_pred :448-536 +
 * mc_prediction/bi_prediction :53-156) */
static void inter_mb(const pic_t* p, int cur, uint8_t pred[3][256])
{
    const h264r_mb* mb = &p->mbs[cur];
    const h264r_slice* sl = &p->slices[mb->slice_idx];
    const h264r_mb_motion* m = &p->motion[cur];
    const int W = p->d->W, H = p->d->H, mbx = cur % W, mby = cur / W;
    const int wY = W * 16, hY = H * 16, wC = W * 8, hC = H * 8;
    int origin[16], dir[16];
    partition_map(p, cur, origin, dir);
    const int is_b = sl->slice_type == H264R_B_SLICE;
    const int uni_weighted = (sl->weighted_pred_flag && !is_b) || (sl->weighted_bipred_idc == 1 && is_b);

    for (int blk = 0; blk < 16; ++blk) {
        const int bx = blk & 3, by = blk >> 2, o = origin[blk], pd = dir[blk];
        for (int pl = 0; pl < 3; ++pl) {
            const int n = pl ? 2 : 4;                              /* block size in this plane */
            int smp[2][16];
            int refidx[2] = { 0, 0 };
            for (int k = 0; k < 2; ++k) {                          /* k = 0: first (or only) list, 1: second */
                if (k == 1 && pd != 2) break;
                int list = pd == 2 ? k : pd;
                refidx[k] = m->ref_idx[list][o];
                const int slot = sl->ref_pic_list[list][refidx[k]];
                const frame_t* rf = &p->d->fr[p->pp->ref_frames[slot]];
                /* absolute quarter-pel position of this 4x4 block's origin (:477-478) */
                int vx = (mbx * 4 + bx) * 16 + m->mv[list][o][0];
                int vy = (mby * 4 + by) * 16 + m->mv[list][o][1];
                /* field pictures: a reference field of the other parity lies half a frame line off in chroma
                 * (get_block_chroma, inter_prediction.cc:352-354) */
                if (pl && p->pp->structure != H264R_FRAME && p->pp->ref_structure[slot] != p->pp->structure)
                    vy += p->pp->structure == H264R_BOTTOM_FIELD ? 2 : -2;
                for (int y = 0; y < n; ++y)
                    for (int x = 0; x < n; ++x)
                        smp[k][y * n + x] = pl == 0
                            ? luma_sample(rf->pl[0], wY, hY, (vx >> 2) + x, (vy >> 2) + y, vx & 3, vy & 3)
                            : chroma_sample(rf->pl[pl], wC, hC, (vx >> 3) + x, (vy >> 3) + y, vx & 7, vy & 7);
            }
            uint8_t* out = pred[pl] + (by * n) * 16 + bx * n;
            const int denom = pl ? sl->chroma_log2_weight_denom : sl->luma_log2_weight_denom;
            for (int i = 0; i < n * n; ++i) {
                int v;
                if (pd != 2) {
                    if (uni_weighted)
                        v = clip1(rshift_rnd(sl->wp_weight[pd][pl][refidx[0]] * smp[0][i], denom) + sl->wp_offset[pd][pl][refidx[0]]);
                    else v = smp[0][i];
                } else if (sl->weighted_bipred_idc == 0) {
                    v = (smp[0][i] + smp[1][i] + 1) >> 1;
                } else {
                    int w0, w1, o0 = 0, o1 = 0;
                    if (sl->weighted_bipred_idc == 1) {
                        w0 = sl->wp_weight[0][pl][refidx[0]]; w1 = sl->wp_weight[1][pl][refidx[1]];
                        o0 = sl->wp_offset[0][pl][refidx[0]]; o1 = sl->wp_offset[1][pl][refidx[1]];
                    } else { w1 = sl->implicit_w1[refidx[0]][refidx[1]]; w0 = 64 - w1; }
                    v = clip1(rshift_rnd(w0 * smp[0][i] + w1 * smp[1][i], denom + 1) + ((o0 + o1 + 1) >> 1));
                }
                out[(i / n) * 16 + (i % n)] = (uint8_t)v;
            }
        }
    }
}

/* ---- macroblock reconstruction (Decoder::decode, decoder.cc:65-262) ------------------------------ */

static void store_block(uint8_t* img, int stride, int px, int py, int w, int h,
                        const uint8_t* pred /*stride 16*/, const int* res /*stride rs*/, int rs, int add)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            img[(py + y) * stride + px + x] = add ? (uint8_t)clip1(pred[y * 16 + x] + res[y * rs + x]) : pred[y * 16 + x];
}

static void reconstruct_mb(const pic_t* p, int cur)
{
    const h264r_mb* mb = &p->mbs[cur];
    const int W = p->d->W, mbx = cur % W, mby = cur / W;
    uint8_t* Y = p->dst->pl[0];
    const int sY = W * 16, sC = W * 8;

    if (mb->mb_type == H264R_MB_IPCM) {                             /* mb_pred_ipcm, decoder.cc:149-168 */
        int16_t c[H264R_COEFFS_PER_MB];
        memset(c, 0, sizeof(c));
        for (int i = 0; i < mb->coeff_count; ++i) {
            h264r_level e = p->levels[mb->coeff_offset + i];
            c[H264R_LEVEL_POS(e)] = (int16_t)H264R_LEVEL_VALUE(e);
        }
        for (int y = 0; y < 16; ++y) for (int x = 0; x < 16; ++x) Y[(mby * 16 + y) * sY + mbx * 16 + x] = (uint8_t)c[y * 16 + x];
        for (int pl = 1; pl <= 2; ++pl)
            for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x)
                p->dst->pl[pl][(mby * 8 + y) * sC + mbx * 8 + x] = (uint8_t)c[256 + (pl - 1) * 64 + y * 8 + x];
        return;
    }

    int res[3][256], has[2];
    mb_residual(p, mb, res, has);
    uint8_t pred[3][256];

    if (mb->flags & H264R_MB_FLAG_INTRA) {                          /* mb_pred_intra, decoder.cc:170-208 */
        if (mb->mb_type == H264R_MB_I16x16) {
            intra_mb_plane(p, cur, 0, mb->intra16_mode, pred[0]);
            store_block(Y, sY, mbx * 16, mby * 16, 16, 16, pred[0], res[0], 16, 1);
        } else {
            const int n = mb->mb_type == H264R_MB_I8x8 ? 8 : 4, nblk = n == 8 ? 4 : 16;
            for (int k = 0; k < nblk; ++k) {                        /* coding (Z) order: each block predicts from the previous ones */
                int xO, yO, mode = (mb->u.intra_modes[k >> 1] >> ((k & 1) * 4)) & 15;
                if (n == 8) { xO = (k & 1) * 8; yO = (k >> 1) * 8; }
                else { xO = ((k >> 2) & 1) * 8 + (k & 1) * 4; yO = (k >> 3) * 8 + ((k >> 1) & 1) * 4; }
                intra_luma_block(p, cur, n, xO, yO, mode, pred[0] + yO * 16 + xO);
                int b8 = (yO >> 3) * 2 + (xO >> 3);
                store_block(Y, sY, mbx * 16 + xO, mby * 16 + yO, n, n, pred[0] + yO * 16 + xO,
                            res[0] + yO * 16 + xO, 16, (has[0] >> b8) & 1);
            }
        }
        for (int pl = 1; pl <= 2; ++pl) {
            intra_mb_plane(p, cur, pl, mb->chroma_mode, pred[pl]);
            store_block(p->dst->pl[pl], sC, mbx * 8, mby * 8, 8, 8, pred[pl], res[pl], 8, 1);
        }
        return;
    }

    inter_mb(p, cur, pred);                                         /* mb_pred_inter + inverse_transform_inter (transform.cc:1051-1095) */
    for (int b8 = 0; b8 < 4; ++b8) {
        int xO = (b8 & 1) * 8, yO = (b8 >> 1) * 8;
        store_block(Y, sY, mbx * 16 + xO, mby * 16 + yO, 8, 8, pred[0] + yO * 16 + xO, res[0] + yO * 16 + xO, 16, (has[0] >> b8) & 1);
    }
    for (int pl = 1; pl <= 2; ++pl)
        store_block(p->dst->pl[pl], sC, mbx * 8, mby * 8, 8, 8, pred[pl], res[pl], 8, has[1]);
}

/* ---- deblocking (decoder/deblock.cc) ------------------------------------------------------------- */

static const uint8_t TAB_ALPHA[52] = {   /* Table 8-16 */
    0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,4,4,5,6,7,8,9,10,12,13,15,17,20,22,25,28,32,36,40,45,50,56,63,71,80,90,101,113,127,144,162,182,203,226,255,255 };
static const uint8_t TAB_BETA[52] = {
    0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,2,2,3,3,3,3,4,4,4,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13,14,14,15,15,16,16,17,17,18,18 };
static const uint8_t TAB_TC0[52][3] = {  /* Table 8-17 */
    {0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},
    {0,0,1},{0,0,1},{0,0,1},{0,0,1},{0,1,1},{0,1,1},{1,1,1},{1,1,1},{1,1,1},{1,1,1},{1,1,2},{1,1,2},{1,1,2},{1,1,2},{1,2,3},{1,2,3},{2,2,3},
    {2,2,4},{2,3,4},{2,3,4},{3,3,5},{3,4,6},{3,4,6},{4,5,7},{4,5,8},{4,6,9},{5,7,10},{6,8,11},{6,8,13},{7,10,14},{8,11,16},{9,12,18},
    {10,13,20},{11,15,23},{13,17,25} };

/* bs_compare_mvs, deblock.cc:35-75.  (mp, bp) / (mq, bq): motion record + 4x4 block index of either side */
/* mvlimit (deblock.cc:86, 164): 4 quarter samples, 2 vertically in field pictures */
static int mv_differs(int mvlimit, const h264r_mb_motion* a, int ba, int la, const h264r_mb_motion* b, int bb, int lb)
{
    return (iabs(a->mv[la][ba][0] - b->mv[lb][bb][0]) >= 4) | (iabs(a->mv[la][ba][1] - b->mv[lb][bb][1]) >= mvlimit);
}
static int bs_compare(int mvlimit, const h264r_mb_motion* mp, int bp, const h264r_mb_motion* mq, int bq)
{
    int p0 = mp->ref_pic[0][bp], p1 = mp->ref_pic[1][bp], q0 = mq->ref_pic[0][bq], q1 = mq->ref_pic[1][bq];
    if (!((p0 == q0 && p1 == q1) || (p0 == q1 && p1 == q0))) return 1;
    if (p0 != p1) {
        if (p0 == q0) return mv_differs(mvlimit, mp, bp, 0, mq, bq, 0) | mv_differs(mvlimit, mp, bp, 1, mq, bq, 1);
        return mv_differs(mvlimit, mp, bp, 0, mq, bq, 1) | mv_differs(mvlimit, mp, bp, 1, mq, bq, 0);
    }
    return (mv_differs(mvlimit, mp, bp, 0, mq, bq, 0) | mv_differs(mvlimit, mp, bp, 1, mq, bq, 1)) &
           (mv_differs(mvlimit, mp, bp, 0, mq, bq, 1) | mv_differs(mvlimit, mp, bp, 1, mq, bq, 0));
}

/* strength_vertical / strength_horizontal (deblock.cc:78-228) for frame pictures without SP/SI.
 * dir 0: vertical edge `edge` (0..3), bS[k] for the 16 rows; dir 1: horizontal edge, bS[k] for the 16 columns */
static void edge_strength(const pic_t* p, int q, int dir, int edge, uint8_t bS[16])
{
    const int W = p->d->W;
    const h264r_mb* Q = &p->mbs[q];
    const int pidx = edge ? q : (dir == 0 ? q - 1 : q - W);
    const h264r_mb* P = &p->mbs[pidx];
    const h264r_slice* sl = &p->slices[Q->slice_idx];
    if (edge > 0 && sl->slice_type == H264R_P_SLICE && Q->mb_type == 0) { memset(bS, 0, 16); return; }
    const int intra = ((P->flags | Q->flags) & H264R_MB_FLAG_INTRA) != 0;
    /* bS 4 on MB edges of frame pictures; in field pictures only on vertical MB edges (cond_bS4, deblock.cc:106-107, 188-189) */
    if (intra) { memset(bS, edge == 0 && (p->pp->structure == H264R_FRAME || dir == 0) ? 4 : 3, 16); return; }
    for (int k4 = 0; k4 < 4; ++k4) {
        int blkQ = dir == 0 ? k4 * 4 + edge : edge * 4 + k4;
        int blkP = dir == 0 ? k4 * 4 + (edge ? edge - 1 : 3) : (edge ? edge - 1 : 3) * 4 + k4;
        int s;
        if (((Q->cbp_blks >> blkQ) & 1) || ((P->cbp_blks >> blkP) & 1)) s = 2;
        else if (edge > 0 && (Q->mb_type == 1 || Q->mb_type == (dir == 0 ? 2 : 3))) s = 0;
        else s = bs_compare(p->pp->structure != H264R_FRAME ? 2 : 4, &p->motion[pidx], blkP, &p->motion[q], blkQ);
        memset(bS + k4 * 4, s, 4);
    }
}

/* filter_strong / filter_normal (deblock.cc:327-415) on the 8 samples across an edge; pix points at q0, step to q1 */
static void filter_samples(uint8_t* pix, int step, int bS, int alpha, int beta, int tc0, int chroma)
{
    int p0 = pix[-step], p1 = pix[-2 * step], q0 = pix[0], q1 = pix[step];
    if (!(iabs(p0 - q0) < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta)) return;
    int p2 = chroma ? 0 : pix[-3 * step], q2 = chroma ? 0 : pix[2 * step];
    if (bS == 4) {
        int ap = iabs(p2 - p0), aq = iabs(q2 - q0), small = iabs(p0 - q0) < (alpha >> 2) + 2;
        if (!chroma && ap < beta && small) {
            int p3 = pix[-4 * step];
            pix[-step]     = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            pix[-2 * step] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
            pix[-3 * step] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else pix[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        if (!chroma && aq < beta && small) {
            int q3 = pix[3 * step];
            pix[0]        = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            pix[step]     = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
            pix[2 * step] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
        } else pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
        return;
    }
    int ap = iabs(p2 - p0), aq = iabs(q2 - q0);
    int tc = chroma ? tc0 + 1 : tc0 + (ap < beta) + (aq < beta);
    int delta = clip3(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
    pix[-step] = (uint8_t)clip1(p0 + delta);
    pix[0]     = (uint8_t)clip1(q0 - delta);
    if (!chroma && ap < beta) pix[-2 * step] = (uint8_t)(p1 + clip3(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - (p1 * 2)) >> 1));
    if (!chroma && aq < beta) pix[step]      = (uint8_t)(q1 + clip3(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - (q1 * 2)) >> 1));
}

/* filter_edge, deblock.cc:418-486.  pl: 0 Y, 1 Cb, 2 Cr.  `edge` in samples of that plane (0,4,8,12 / 0,4) */
static void filter_edge(const pic_t* p, int q, int pl, int dir, int edge, const uint8_t bS[16])
{
    const int W = p->d->W, n = pl ? 8 : 16, stride = W * n;
    int any = 0;
    for (int k = 0; k < 16; ++k) any |= bS[k];
    if (!any) return;
    const h264r_mb* Q = &p->mbs[q];
    const h264r_mb* P = edge ? Q : &p->mbs[dir == 0 ? q - 1 : q - W];
    const h264r_slice* sl = &p->slices[Q->slice_idx];
    int qPp = pl ? P->qp_c[pl - 1] : P->qp_y, qPq = pl ? Q->qp_c[pl - 1] : Q->qp_y;
    int qPav = (qPp + qPq + 1) >> 1;
    int indexA = clip3(0, 51, qPav + sl->filter_offset_a), indexB = clip3(0, 51, qPav + sl->filter_offset_b);
    int alpha = TAB_ALPHA[indexA], beta = TAB_BETA[indexB];
    uint8_t* img = p->dst->pl[pl];
    int px = (q % W) * n, py = (q / W) * n;
    for (int pel = 0; pel < n; ++pel) {
        int s = bS[pl ? pel << 1 : pel];
        if (!s) continue;
        uint8_t* pix = dir == 0 ? img + (py + pel) * stride + px + edge : img + (py + edge) * stride + px + pel;
        filter_samples(pix, dir == 0 ? 1 : stride, s, alpha, beta, s < 4 ? TAB_TC0[indexA][s - 1] : 0, pl != 0);
    }
}

/* Deblock::strength + filter_vertical + filter_horizontal for one MB (deblock.cc:230-289, 488-535) */
static void deblock_mb(const pic_t* p, int q)
{
    const int W = p->d->W, mbx = q % W, mby = q / W;
    const h264r_mb* Q = &p->mbs[q];
    const h264r_slice* sl = &p->slices[Q->slice_idx];
    const int idc = sl->disable_deblocking_filter_idc;
    if (idc == 1) return;
    int left = mbx > 0, top = mby > 0;
    if (idc == 2) {
        left = left && p->mbs[q - 1].slice_idx == Q->slice_idx;
        top  = top  && p->mbs[q - W].slice_idx == Q->slice_idx;
    }
    const int t8 = (Q->flags & H264R_MB_FLAG_T8x8) != 0;
    uint8_t bS[4][16];
    for (int dir = 0; dir < 2; ++dir) {
        int mbedge = dir == 0 ? left : top;
        for (int e = 0; e < 4; ++e) {
            int luma_on = e == 0 ? mbedge : !(t8 && (e & 1));
            if (luma_on) edge_strength(p, q, dir, e, bS[e]);
        }
        for (int e = 0; e < 4; ++e) {
            int luma_on = e == 0 ? mbedge : !(t8 && (e & 1));
            int chroma_on = e == 0 ? mbedge : e == 1;              /* chroma edges 2, 3 off for 4:2:0 */
            if (luma_on) filter_edge(p, q, 0, dir, e * 4, bS[e]);
            if (chroma_on) {                                        /* chroma edge e uses luma edge 2e's strengths */
                filter_edge(p, q, 1, dir, e * 4, bS[e * 2]);
                filter_edge(p, q, 2, dir, e * 4, bS[e * 2]);
            }
        }
    }
}

static double now_sec(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int port_reconstruct(port_dec* d, int dst, const h264r_pic_params* pp, int used_for_reference,
                     const h264r_slice* slices, const h264r_mb* mbs, const h264r_mb_motion* motion,
                     const h264r_level* levels, double* sec_decode, double* sec_deblock)
{
    pic_t p = { d, pp, slices, mbs, motion, levels, &d->fr[dst] };
    const int nmb = d->W * d->H;
    (void)used_for_reference;
    double t0 = now_sec();
    for (int a = 0; a < nmb; ++a) reconstruct_mb(&p, a);
    double t1 = now_sec();
    if (pp->run_deblock)                                            /* deblock.cc:631-643 */
        for (int a = 0; a < nmb; ++a) deblock_mb(&p, a);            /* deblock_pic pass 2, raster order, V then H */
    double t2 = now_sec();
    if (sec_decode) *sec_decode = t1 - t0;
    if (sec_deblock) *sec_deblock = t2 - t1;
    return 0;
}
