// Deterministic synthetic macroblock-data generator.  See include/h264synth.h and SURVEY.md §8d.
//
// What is generated is what the reference's parser would leave behind at the Decoder boundary:
//   * mb_t fields after Parser::Macroblock::parse (parser/interpret_mb.cc:178-316) -- mb_type, SubMbType and
//     SubMbPredMode exactly as mb_type_{i,p,b}_slice / sub_mb_type fill them (interpret_mb.cc:346-403, 472-496),
//   * pic_motion_params after mb_pred_inter / skip_macroblock / direct derivation
//     (interpret_mb.cc:573-624, interpret_mv.cc:159-190, 300-434): unused lists hold ref_idx=-1, ref_pic=null,
//     mv=0, except P_Skip whose list-1 entry keeps the calloc'd zeros,
//   * raw coefficient levels at their inverse-scanned raster positions plus the cbp_blks bits
//     Transform::coeff_luma_ac sets (decoder/transform.cc:431-440).
#include "h264synth.h"
#include "h264_tables.h"

#include <string.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

namespace {

struct Rng {
    uint64_t s;
    uint64_t next()
    {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    uint32_t u32() { return (uint32_t)(next() >> 32); }
    int below(int n) { return (int)(((uint64_t)u32() * (uint64_t)n) >> 32); }   // uniform [0, n)
    int range(int lo, int hi) { return lo + below(hi - lo + 1); }               // uniform [lo, hi]
    bool chance(int percent) { return below(100) < percent; }
};

inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

struct PicPlan {
    int type;              // H264R_*_SLICE
    int poc;
    int is_ref;
    int l0[2], n_l0;       // decode-order indexes
    int l1[1], n_l1;
    int all_intra;
};

} // namespace

struct h264s_stream {
    int config, stream_idx;
    int W, H, num_frames;
    Rng rng;
    std::vector<PicPlan> plan;
    std::vector<int> last_use;        // decode index of the last picture referencing picture i
    int next_pic;
    // stream-level properties
    int transform8x8;                 // pps.transform_8x8_mode_flag
    int default_matrices;             // scaling lists: 0 flat, 1 Default Tables 7-3/7-4
    int chroma_qp_offset[2];
    int qp_lo, qp_hi;
    int allow_b;
    int direct8x8;                    // sps.direct_8x8_inference_flag: 0 on every third stream of the B-picture configs
    h264r::ZigZag zz;
};

namespace {

void build_plan(h264s_stream* s)
{
    const int n = s->num_frames;
    s->plan.clear();
    if (!s->allow_b) {                               // I P P P ...
        for (int i = 0; i < n; ++i) {
            PicPlan p; memset(&p, 0, sizeof(p));
            p.type = i == 0 ? H264R_I_SLICE : H264R_P_SLICE;
            p.poc = 2 * i; p.is_ref = 1;
            if (i > 0) { p.l0[0] = i - 1; p.n_l0 = 1; }
            s->plan.push_back(p);
        }
    } else {                                         // display I B B P B B P ... ; decode I P B B P B B ...
        // anchors at display 0, 3, 6, ...; if the tail has fewer than 3 pictures they become P anchors.
        std::vector<int> anchors_dec;                // decode indexes of anchors, in order
        int disp = 0, prev_anchor_disp = -1;
        while ((int)s->plan.size() < n) {
            int remaining = n - (int)s->plan.size();
            PicPlan a; memset(&a, 0, sizeof(a));
            int nb = 0;
            if (prev_anchor_disp < 0) { a.type = H264R_I_SLICE; disp = 0; }
            else {
                nb = remaining >= 3 ? 2 : remaining - 1;
                disp = prev_anchor_disp + nb + 1;
                a.type = H264R_P_SLICE;
                if (s->config == H264S_CFG_4K_HIGH && (int)s->plan.size() == 4) a.type = H264R_I_SLICE;
            }
            a.poc = 2 * disp; a.is_ref = 1;
            if (a.type == H264R_P_SLICE) {
                int na = (int)anchors_dec.size();
                a.l0[0] = anchors_dec[na - 1]; a.n_l0 = 1;
                if (na >= 2) { a.l0[1] = anchors_dec[na - 2]; a.n_l0 = 2; }
            }
            int a_dec = (int)s->plan.size();
            s->plan.push_back(a);
            for (int b = 0; b < nb; ++b) {
                PicPlan p; memset(&p, 0, sizeof(p));
                p.type = H264R_B_SLICE; p.poc = 2 * (prev_anchor_disp + 1 + b); p.is_ref = 0;
                int na = (int)anchors_dec.size();
                p.l0[0] = anchors_dec[na - 1]; p.n_l0 = 1;
                if (na >= 2) { p.l0[1] = anchors_dec[na - 2]; p.n_l0 = 2; }
                p.l1[0] = a_dec; p.n_l1 = 1;
                s->plan.push_back(p);
            }
            anchors_dec.push_back(a_dec);
            prev_anchor_disp = disp;
        }
    }
    if (s->config == H264S_CFG_4K_HIGH) { s->plan[0].all_intra = 1; }
    s->last_use.assign(n, -1);
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < s->plan[i].n_l0; ++k) s->last_use[s->plan[i].l0[k]] = i;
        for (int k = 0; k < s->plan[i].n_l1; ++k) s->last_use[s->plan[i].l1[k]] = i;
    }
}

// ---- per-picture generation context -----------------------------------------------------------------
struct Gen {
    h264s_stream* s;
    Rng* rng;
    int W, H;
    const PicPlan* plan;
    h264r_mb* mbs;
    h264r_mb_motion* motion;
    h264r_slice* slices;
    int16_t dense[H264R_COEFFS_PER_MB];   // the current MB's levels at raster positions (scratch)
    bool     dense_used;
    h264r_level* levels;
    uint32_t n_levels, level_capacity;
    bool     overflow;
    int num_slices;
    int slice_first_mb[4];
    int qp;                           // running QpY

    bool mb_available(int cur, int nb_x, int nb_y, bool need_intra) const
    {
        if (nb_x < 0 || nb_x >= W || nb_y < 0 || nb_y >= H) return false;
        int nb = nb_y * W + nb_x;
        if (nb >= cur) return false;                              // not decoded yet (slice_nr == -1)
        if (mbs[nb].slice_idx != mbs[cur].slice_idx) return false;
        if (need_intra && !(mbs[nb].flags & H264R_MB_FLAG_INTRA)) return false;
        return true;
    }
};

// level magnitude: mostly small, rare large (bounded so that dequantisation stays far inside int32)
int draw_level(Rng& r, int qp)
{
    int lim = std::max(1, 2048 >> (qp / 6));
    int t = r.below(100), m;
    if (t < 60) m = 1;
    else if (t < 85) m = r.range(2, 3);
    else if (t < 96) m = r.range(4, 8);
    else if (t < 99) m = r.range(9, 40);
    else m = r.range(41, 2048);
    m = std::min(m, lim);
    return r.chance(50) ? -m : m;
}

// number of non-zero levels in one block of `maxc` scan positions: 1 + Geom(0.4), clipped
int draw_nnz(Rng& r, int maxc)
{
    int n = 1;
    while (n < maxc && r.below(100) < 60) ++n;
    return n;
}

// place `nnz` levels at low scan positions [start, start+span) of a 4x4 block whose top-left raster sample
// is dst (row stride `stride`); returns whether anything non-zero was written
bool fill_block4x4(Gen& g, int16_t* dst, int stride, int start, int qp)
{
    Rng& r = *g.rng;
    int nnz = draw_nnz(r, 16 - start);
    int pos = start;
    bool any = false;
    for (int k = 0; k < nnz && pos < 16; ++k) {
        pos += (k == 0) ? r.below(2) : r.below(3);
        if (pos >= 16) break;
        int x = g.s->zz.x4[pos], y = g.s->zz.y4[pos];
        dst[y * stride + x] = (int16_t)draw_level(r, qp);
        any = true;
        ++pos;
    }
    return any;
}

bool fill_block8x8(Gen& g, int16_t* dst, int stride, int qp)
{
    Rng& r = *g.rng;
    int nnz = draw_nnz(r, 64) + r.below(6);
    int pos = 0;
    bool any = false;
    for (int k = 0; k < nnz && pos < 64; ++k) {
        pos += (k == 0) ? r.below(2) : r.below(4);
        if (pos >= 64) break;
        int x = g.s->zz.x8[pos], y = g.s->zz.y8[pos];
        dst[y * stride + x] = (int16_t)draw_level(r, qp);
        any = true;
        ++pos;
    }
    return any;
}

int16_t* take_slot(Gen& g, h264r_mb&)
{
    if (!g.dense_used) { memset(g.dense, 0, sizeof(g.dense)); g.dense_used = true; }
    return g.dense;
}

// what the residual parser hands over: one entry per non-zero level (every sample for I_PCM)
void emit_levels(Gen& g, h264r_mb& mb)
{
    mb.coeff_offset = g.n_levels; mb.coeff_count = 0;
    if (!g.dense_used) return;
    const bool pcm = mb.mb_type == H264R_MB_IPCM;
    int n = 0;
    for (int p = 0; p < H264R_COEFFS_PER_MB; ++p) {
        if (!pcm && g.dense[p] == 0) continue;
        if (g.n_levels + n >= g.level_capacity) { g.overflow = true; break; }
        g.levels[g.n_levels + n] = H264R_LEVEL(p, g.dense[p]);
        ++n;
    }
    mb.coeff_count = (uint16_t)n;
    g.n_levels += n;
}

// residual for a non-I16x16, non-PCM MB.  Sets cbp_luma/cbp_chroma/cbp_blks and writes levels.
void gen_residual(Gen& g, h264r_mb& mb, int coded_percent)
{
    Rng& r = *g.rng;
    const bool t8 = (mb.flags & H264R_MB_FLAG_T8x8) != 0;
    int cbp = 0;
    for (int b = 0; b < 4; ++b) if (r.chance(coded_percent)) cbp |= 1 << b;
    if (t8 && cbp == 0 && !(mb.flags & H264R_MB_FLAG_INTRA)) cbp = 1 << r.below(4);  // t8x8 on inter needs cbp_luma > 0
    mb.cbp_luma = (uint8_t)cbp;
    mb.cbp_chroma = (uint8_t)r.below(3);
    if (g.s->config == H264S_CFG_CIF_BASELINE && r.chance(30)) mb.cbp_chroma = 0;
    unsigned blks = 0;
    for (int b = 0; b < 4; ++b) {
        if (!(cbp & (1 << b))) continue;
        int16_t* slot = take_slot(g, mb);
        int bx0 = (b & 1) * 2, by0 = (b >> 1) * 2;                 // in 4x4 units
        if (t8) {
            if (r.chance(90) && fill_block8x8(g, slot + by0 * 4 * 16 + bx0 * 4, 16, mb.qp_y))
                blks |= 0x33u << (by0 * 4 + bx0);
        } else {
            for (int k = 0; k < 4; ++k) {
                int bx = bx0 + (k & 1), by = by0 + (k >> 1);
                if (r.chance(70) && fill_block4x4(g, slot + by * 4 * 16 + bx * 4, 16, 0, mb.qp_y))
                    blks |= 1u << (by * 4 + bx);
            }
        }
    }
    mb.cbp_blks = (uint16_t)blks;
    if (mb.cbp_chroma) {
        int16_t* slot = take_slot(g, mb);
        for (int pl = 0; pl < 2; ++pl) {
            int16_t* c = slot + 256 + pl * 64;
            for (int k = 0; k < 4; ++k)                            // chroma DC at (4*(k%2), 4*(k/2))
                if (r.chance(60)) c[(k >> 1) * 4 * 8 + (k & 1) * 4] = (int16_t)draw_level(r, mb.qp_c[pl]);
            if (mb.cbp_chroma == 2)
                for (int k = 0; k < 4; ++k)
                    if (r.chance(60)) fill_block4x4(g, c + (k >> 1) * 4 * 8 + (k & 1) * 4, 8, 1, mb.qp_c[pl]);
        }
    }
}

void set_nibble(uint8_t* a, int i, int v) { a[i >> 1] = (uint8_t)((a[i >> 1] & (i & 1 ? 0x0F : 0xF0)) | (v << ((i & 1) * 4))); }

int pick_mode(Rng& r, bool a, bool b, bool d, int nmodes_9)
{
    // legal set per SURVEY §8a quirk 11
    int legal[9], n = 0;
    if (nmodes_9) {
        legal[n++] = 2;                                   // DC
        if (b) { legal[n++] = 0; legal[n++] = 3; legal[n++] = 7; }
        if (a) { legal[n++] = 1; legal[n++] = 8; }
        if (a && b && d) { legal[n++] = 4; legal[n++] = 5; legal[n++] = 6; }
    }
    return legal[r.below(n)];
}

void gen_intra_mb(Gen& g, int addr, int kind /*8,9,10,12*/)
{
    Rng& r = *g.rng;
    h264r_mb& mb = g.mbs[addr];
    const h264r_slice& sl = g.slices[mb.slice_idx];
    const int mbx = addr % g.W, mby = addr / g.W;
    const bool ci = sl.constrained_intra_pred_flag != 0;
    mb.mb_type = (uint8_t)kind;
    mb.flags = H264R_MB_FLAG_INTRA | (kind == H264R_MB_I8x8 ? H264R_MB_FLAG_T8x8 : 0);
    const bool availL  = g.mb_available(addr, mbx - 1, mby, ci);
    const bool availT  = g.mb_available(addr, mbx, mby - 1, ci);
    const bool availTL = g.mb_available(addr, mbx - 1, mby - 1, ci);

    if (kind == H264R_MB_IPCM) {
        mb.qp_y = 0;
        for (int i = 0; i < 2; ++i)
            mb.qp_c[i] = (int8_t)h264r::qpc_from_qpi(clip3(0, 51, 0 + g.s->chroma_qp_offset[i]));
        mb.cbp_blks = 0xFFFF;
        int16_t* slot = take_slot(g, mb);
        int base = r.range(16, 235), amp = r.range(0, 20);
        for (int i = 0; i < 384; ++i) slot[i] = (int16_t)clip3(0, 255, base + r.range(-amp, amp));
        return;
    }

    // chroma mode: DC always, H needs A, V needs B, plane needs A, B, D
    {
        int legal[4], n = 0;
        legal[n++] = 0;
        if (availL) legal[n++] = 1;
        if (availT) legal[n++] = 2;
        if (availL && availT && availTL) legal[n++] = 3;
        mb.chroma_mode = (uint8_t)legal[r.below(n)];
    }

    if (kind == H264R_MB_I16x16) {
        int legal[4], n = 0;
        legal[n++] = 2;
        if (availT) legal[n++] = 0;
        if (availL) legal[n++] = 1;
        if (availL && availT && availTL) legal[n++] = 3;
        mb.intra16_mode = (uint8_t)legal[r.below(n)];
        int16_t* slot = take_slot(g, mb);                          // I16x16 always owns a slot (luma DC)
        for (int k = 0; k < 16; ++k)
            if (r.chance(45)) slot[(k >> 2) * 4 * 16 + (k & 3) * 4] = (int16_t)draw_level(r, mb.qp_y);
        mb.cbp_luma = r.chance(50) ? 15 : 0;
        unsigned blks = 0;
        if (mb.cbp_luma)
            for (int k = 0; k < 16; ++k)
                if (r.chance(60) && fill_block4x4(g, slot + (k >> 2) * 4 * 16 + (k & 3) * 4, 16, 1, mb.qp_y))
                    blks |= 1u << k;
        mb.cbp_blks = (uint16_t)blks;
        mb.cbp_chroma = (uint8_t)r.below(3);
        if (mb.cbp_chroma) {
            for (int pl = 0; pl < 2; ++pl) {
                int16_t* c = slot + 256 + pl * 64;
                for (int k = 0; k < 4; ++k)
                    if (r.chance(60)) c[(k >> 1) * 4 * 8 + (k & 1) * 4] = (int16_t)draw_level(r, mb.qp_c[pl]);
                if (mb.cbp_chroma == 2)
                    for (int k = 0; k < 4; ++k)
                        if (r.chance(60)) fill_block4x4(g, c + (k >> 1) * 4 * 8 + (k & 1) * 4, 8, 1, mb.qp_c[pl]);
            }
        }
        return;
    }

    const bool availTR = g.mb_available(addr, mbx + 1, mby - 1, ci);
    (void)availTR;                                                 // C only substitutes samples; never gates a mode
    if (kind == H264R_MB_I4x4) {
        for (int blk = 0; blk < 16; ++blk) {                       // luma4x4BlkIdx (Z order)
            int bx = ((blk >> 2) & 1) * 2 + (blk & 1), by = (blk >> 3) * 2 + ((blk >> 1) & 1);
            bool a = bx > 0 ? true : availL;
            bool b = by > 0 ? true : availT;
            bool d = (bx > 0 && by > 0) ? true : (bx > 0 ? availT : (by > 0 ? availL : availTL));
            set_nibble(mb.u.intra_modes, blk, pick_mode(r, a, b, d, 1));
        }
    } else {
        for (int blk = 0; blk < 4; ++blk) {
            int bx = blk & 1, by = blk >> 1;
            bool a = bx > 0 ? true : availL;
            bool b = by > 0 ? true : availT;
            bool d = (bx > 0 && by > 0) ? true : (bx > 0 ? availT : (by > 0 ? availL : availTL));
            set_nibble(mb.u.intra_modes, blk, pick_mode(r, a, b, d, 1));
        }
    }
    gen_residual(g, mb, g.s->config == H264S_CFG_4K_HIGH ? 75 : 55);
}

void clear_motion(h264r_mb_motion& m)
{
    memset(m.mv, 0, sizeof(m.mv));
    memset(m.ref_idx, -1, sizeof(m.ref_idx));
    memset(m.ref_pic, -1, sizeof(m.ref_pic));
}

void draw_mv(Gen& g, int mbx, int mby, int16_t mv[2])
{
    Rng& r = *g.rng;
    const bool edge = mbx == 0 || mby == 0 || mbx == g.W - 1 || mby == g.H - 1;
    if (edge && r.chance(25)) {            // vectors reaching (far) outside the picture
        int rx = r.chance(30) ? 900 : 200, ry = r.chance(30) ? 700 : 200;
        mv[0] = (int16_t)r.range(-rx, rx); mv[1] = (int16_t)r.range(-ry, ry);
    } else if (r.chance(2)) {
        mv[0] = (int16_t)r.range(-200, 200); mv[1] = (int16_t)r.range(-200, 200);
    } else if (r.chance(15)) {
        mv[0] = (int16_t)(4 * r.range(-8, 8)); mv[1] = (int16_t)(4 * r.range(-8, 8));   // integer-pel
    } else {
        mv[0] = (int16_t)r.range(-64, 64); mv[1] = (int16_t)r.range(-64, 64);
    }
}

// fill the 4x4 blocks [bx0,bx0+w) x [by0,by0+h) of one partition
void set_part(Gen& g, h264r_mb_motion& m, const h264r_slice& sl, int mbx, int mby,
              int bx0, int by0, int w, int h, int pred_dir, int ref0, int ref1)
{
    int16_t mv0[2] = {0, 0}, mv1[2] = {0, 0};
    if (pred_dir != H264R_PRED_L1) draw_mv(g, mbx, mby, mv0);
    if (pred_dir != H264R_PRED_L0) draw_mv(g, mbx, mby, mv1);
    for (int y = by0; y < by0 + h; ++y)
        for (int x = bx0; x < bx0 + w; ++x) {
            int b = y * 4 + x;
            if (pred_dir != H264R_PRED_L1) {
                m.mv[0][b][0] = mv0[0]; m.mv[0][b][1] = mv0[1];
                m.ref_idx[0][b] = (int8_t)ref0; m.ref_pic[0][b] = sl.ref_pic_list[0][ref0];
            }
            if (pred_dir != H264R_PRED_L0) {
                m.mv[1][b][0] = mv1[0]; m.mv[1][b][1] = mv1[1];
                m.ref_idx[1][b] = (int8_t)ref1; m.ref_pic[1][b] = sl.ref_pic_list[1][ref1];
            }
        }
}

// One direct-predicted 8x8 quadrant (B_Skip / B_Direct_16x16, or a direct sub-macroblock of B_8x8) after direct-mode
// resolution (parser/interpret_mv.cc:116-434).  Spatial: the direction and (refIdxL0, refIdxL1) are the MB's; the vector of
// a list is the predicted vector or zero (colZeroFlag), decided per 8x8 block -- per 4x4 block without
// direct_8x8_inference_flag.  Temporal: bi-predictive; refIdxL0 and both vectors follow the co-located block, again per
// 8x8 or per 4x4.  mv0 == nullptr: the quadrant draws its own predicted vectors.
void gen_direct_quadrant(Gen& g, h264r_mb& mb, h264r_mb_motion& m, const h264r_slice& sl, int mbx, int mby, int q,
                         int dir, int ref0, int ref1, const int16_t* mv0, const int16_t* mv1)
{
    Rng& r = *g.rng;
    const int bx0 = (q & 1) * 2, by0 = (q >> 1) * 2;
    const bool fine = !g.s->direct8x8;                               // motion per 4x4 block
    if (!sl.direct_spatial_mv_pred_flag) {
        mb.u.inter.sub_mb_pred_mode[q] = H264R_PRED_BI;
        if (!fine) { set_part(g, m, sl, mbx, mby, bx0, by0, 2, 2, H264R_PRED_BI, r.below(sl.num_ref[0]), 0); return; }
        for (int k = 0; k < 4; ++k)
            set_part(g, m, sl, mbx, mby, bx0 + (k & 1), by0 + (k >> 1), 1, 1, H264R_PRED_BI, r.below(sl.num_ref[0]), 0);
        return;
    }
    mb.u.inter.sub_mb_pred_mode[q] = (uint8_t)dir;
    if (!mv0 && !fine) { set_part(g, m, sl, mbx, mby, bx0, by0, 2, 2, dir, ref0, ref1); return; }
    int16_t own0[2] = {0, 0}, own1[2] = {0, 0};
    if (!mv0) {
        if (dir != H264R_PRED_L1) draw_mv(g, mbx, mby, own0);
        if (dir != H264R_PRED_L0) draw_mv(g, mbx, mby, own1);
        mv0 = own0; mv1 = own1;
    }
    bool z0 = false, z1 = false;
    if (!fine) { z0 = r.chance(20); z1 = r.chance(20); }
    for (int k = 0; k < 4; ++k) {
        const int b = (by0 + (k >> 1)) * 4 + bx0 + (k & 1);
        if (fine) { z0 = r.chance(30); z1 = r.chance(30); }
        if (dir != H264R_PRED_L1) {
            m.ref_idx[0][b] = (int8_t)ref0; m.ref_pic[0][b] = sl.ref_pic_list[0][ref0];
            m.mv[0][b][0] = z0 ? 0 : mv0[0]; m.mv[0][b][1] = z0 ? 0 : mv0[1];
        }
        if (dir != H264R_PRED_L0) {
            m.ref_idx[1][b] = (int8_t)ref1; m.ref_pic[1][b] = sl.ref_pic_list[1][ref1];
            m.mv[1][b][0] = z1 ? 0 : mv1[0]; m.mv[1][b][1] = z1 ? 0 : mv1[1];
        }
    }
}

static const int kBlockStep[8][2] = { {0,0}, {4,4}, {4,2}, {2,4}, {2,2}, {2,1}, {1,2}, {1,1} };

void gen_inter_mb(Gen& g, int addr)
{
    Rng& r = *g.rng;
    h264r_mb& mb = g.mbs[addr];
    h264r_mb_motion& m = g.motion[addr];
    const h264r_slice& sl = g.slices[mb.slice_idx];
    const int mbx = addr % g.W, mby = addr / g.W;
    const bool is_b = sl.slice_type == H264R_B_SLICE;
    const int cfg = g.s->config;
    mb.flags = 0;
    clear_motion(m);

    auto pick_dir = [&]() -> int {
        if (!is_b) return H264R_PRED_L0;
        int t = r.below(100);
        return t < 40 ? H264R_PRED_BI : (t < 70 ? H264R_PRED_L0 : H264R_PRED_L1);
    };
    auto pick_ref = [&](int list) -> int { return r.below(sl.num_ref[list]); };

    int t = r.below(100);
    const int skip_pct = is_b ? 20 : 25;
    if (t < skip_pct) {
        // P_Skip / B_Skip / B_Direct_16x16
        mb.mb_type = H264R_MB_SKIP_DIRECT;
        if (!is_b) {
            // skip_macroblock (interpret_mv.cc:159-190): list 0 only, refIdx 0; list 1 keeps calloc zeros
            int16_t mv[2]; draw_mv(g, mbx, mby, mv);
            if (r.chance(40)) { mv[0] = mv[1] = 0; }
            for (int b = 0; b < 16; ++b) {
                m.mv[0][b][0] = mv[0]; m.mv[0][b][1] = mv[1];
                m.ref_idx[0][b] = 0; m.ref_pic[0][b] = sl.ref_pic_list[0][0];
                m.ref_idx[1][b] = 0; m.ref_pic[1][b] = -1;          // quirk 2
            }
            memset(mb.u.inter.sub_mb_type, 0, 4); memset(mb.u.inter.sub_mb_pred_mode, 0, 4);
            mb.cbp_luma = mb.cbp_chroma = 0; mb.cbp_blks = 0;
            return;
        }
        // B direct: resolved direction per MB (spatial: one (refIdxL0, refIdxL1) pair for the whole MB,
        // mv per 8x8 -- per 4x4 without direct_8x8_inference -- may be zeroed by colZeroFlag; temporal: bi, ref/mv per
        // 8x8 / 4x4 block)
        memset(mb.u.inter.sub_mb_type, 0, 4);
        int ddir = H264R_PRED_BI, dref0 = 0, dref1 = 0;
        int16_t dmv0[2] = {0, 0}, dmv1[2] = {0, 0};
        if (sl.direct_spatial_mv_pred_flag) {
            ddir = pick_dir(); dref0 = pick_ref(0); dref1 = pick_ref(1);
            draw_mv(g, mbx, mby, dmv0); draw_mv(g, mbx, mby, dmv1);
        }
        for (int q = 0; q < 4; ++q) gen_direct_quadrant(g, mb, m, sl, mbx, mby, q, ddir, dref0, dref1, dmv0, dmv1);
        if (r.chance(50)) { mb.cbp_luma = mb.cbp_chroma = 0; mb.cbp_blks = 0; }     // B_Skip
        else {                                                                      // B_Direct_16x16
            if (g.s->transform8x8 && r.chance(50) && g.s->direct8x8) mb.flags |= H264R_MB_FLAG_T8x8;   // 8x8 transform on direct MBs needs the inference flag
            gen_residual(g, mb, 50);
            if (mb.cbp_luma == 0) mb.flags &= (uint8_t)~H264R_MB_FLAG_T8x8;
        }
        return;
    }

    int u = r.below(100);
    bool all_ge_8x8 = true;
    if (u < 40) {
        mb.mb_type = H264R_MB_16x16;
        int dir = pick_dir();
        memset(mb.u.inter.sub_mb_type, H264R_MB_16x16, 4);
        memset(mb.u.inter.sub_mb_pred_mode, dir, 4);
        set_part(g, m, sl, mbx, mby, 0, 0, 4, 4, dir, pick_ref(0), is_b ? pick_ref(1) : 0);
    } else if (u < 60) {
        bool horiz = r.chance(50);                                   // 16x8 : 8x16
        mb.mb_type = horiz ? H264R_MB_16x8 : H264R_MB_8x16;
        memset(mb.u.inter.sub_mb_type, mb.mb_type, 4);
        int d0 = pick_dir(), d1 = pick_dir();
        for (int i = 0; i < 4; ++i)                                  // interpret_mb.cc:397-401
            mb.u.inter.sub_mb_pred_mode[i] = (uint8_t)(horiz ? (i / 2 ? d1 : d0) : (i % 2 ? d1 : d0));
        if (horiz) {
            set_part(g, m, sl, mbx, mby, 0, 0, 4, 2, d0, pick_ref(0), is_b ? pick_ref(1) : 0);
            set_part(g, m, sl, mbx, mby, 0, 2, 4, 2, d1, pick_ref(0), is_b ? pick_ref(1) : 0);
        } else {
            set_part(g, m, sl, mbx, mby, 0, 0, 2, 4, d0, pick_ref(0), is_b ? pick_ref(1) : 0);
            set_part(g, m, sl, mbx, mby, 2, 0, 2, 4, d1, pick_ref(0), is_b ? pick_ref(1) : 0);
        }
    } else {
        mb.mb_type = H264R_MB_8x8;
        // spatial direct sub-blocks of one MB share (refIdxL0, refIdxL1) and hence the direction
        const int ddir = pick_dir(), dref0 = pick_ref(0), dref1 = is_b ? pick_ref(1) : 0;
        const int16_t* const dmv0 = nullptr; const int16_t* const dmv1 = nullptr;
        for (int q = 0; q < 4; ++q) {
            int bx0 = (q & 1) * 2, by0 = (q >> 1) * 2;
            int st = r.range(is_b ? 3 : 4, 7);                       // 3 stands for "direct" in B slices
            if (st == 3) {
                mb.u.inter.sub_mb_type[q] = 0;
                gen_direct_quadrant(g, mb, m, sl, mbx, mby, q, ddir, dref0, dref1, dmv0, dmv1);
                if (!g.s->direct8x8) all_ge_8x8 = false;             // 4x4-granular direct motion excludes the 8x8 transform
                continue;
            }
            int dir = pick_dir();
            mb.u.inter.sub_mb_type[q] = (uint8_t)st;
            mb.u.inter.sub_mb_pred_mode[q] = (uint8_t)dir;
            if (st != H264R_MB_8x8) all_ge_8x8 = false;
            int ref0 = pick_ref(0), ref1 = is_b ? pick_ref(1) : 0;   // one ref per 8x8, mv per sub-partition
            int sw = kBlockStep[st][0], sh = kBlockStep[st][1];
            for (int y = 0; y < 2; y += sh)
                for (int x = 0; x < 2; x += sw)
                    set_part(g, m, sl, mbx, mby, bx0 + x, by0 + y, sw, sh, dir, ref0, ref1);
        }
    }
    if (g.s->transform8x8 && all_ge_8x8 && r.chance(50)) mb.flags |= H264R_MB_FLAG_T8x8;
    gen_residual(g, mb, cfg == H264S_CFG_4K_HIGH ? 65 : 50);
    if (mb.cbp_luma == 0) mb.flags &= (uint8_t)~H264R_MB_FLAG_T8x8;
}

void fill_slice_tables(h264s_stream* s, Gen& g, h264r_slice& sl, const PicPlan& plan, int pic_idx, int slice_no,
                       const int* ref_slot_l0, const int* ref_slot_l1)
{
    Rng& r = *g.rng;
    memset(&sl, 0, sizeof(sl));
    memset(sl.ref_pic_list, -1, sizeof(sl.ref_pic_list));
    sl.slice_type = (uint8_t)plan.type;
    const int cfg = s->config;
    // deblocking
    sl.disable_deblocking_filter_idc = 0;
    if (g.num_slices > 1 && r.chance(50)) sl.disable_deblocking_filter_idc = 2;
    if (cfg == H264S_CFG_720P_MAIN && pic_idx % 11 == 10 && slice_no == 0) sl.disable_deblocking_filter_idc = 1;
    if (cfg == H264S_CFG_4K_HIGH) { sl.filter_offset_a = (int8_t)(r.chance(50) ? 6 : -6); sl.filter_offset_b = (int8_t)(r.chance(50) ? 6 : -6); }
    else if (cfg == H264S_CFG_CIF_BASELINE) { if (pic_idx % 4 == 3) sl.filter_offset_a = sl.filter_offset_b = -2; }
    else { sl.filter_offset_a = (int8_t)(2 * r.range(-3, 3)); sl.filter_offset_b = (int8_t)(2 * r.range(-3, 3)); }

    sl.constrained_intra_pred_flag = (uint8_t)((cfg == H264S_CFG_1080P_HIGH || cfg == H264S_CFG_MULTI_1080P || cfg == H264S_CFG_1080I_FIELDS || cfg == H264S_CFG_4K_HIGH)
                                               && plan.type == H264R_P_SLICE && (pic_idx % 3) == 1);
    sl.direct_spatial_mv_pred_flag = (uint8_t)r.chance(50);
    sl.num_ref[0] = (uint8_t)plan.n_l0; sl.num_ref[1] = (uint8_t)plan.n_l1;
    for (int i = 0; i < plan.n_l0; ++i) sl.ref_pic_list[0][i] = (int8_t)ref_slot_l0[i];
    for (int i = 0; i < plan.n_l1; ++i) sl.ref_pic_list[1][i] = (int8_t)ref_slot_l1[i];

    // weighted prediction
    sl.luma_log2_weight_denom = sl.chroma_log2_weight_denom = 0;
    if (cfg != H264S_CFG_CIF_BASELINE) {
        if (plan.type == H264R_P_SLICE) sl.weighted_pred_flag = (uint8_t)(cfg == H264S_CFG_720P_MAIN ? 1 : r.chance(50));
        if (plan.type == H264R_B_SLICE) sl.weighted_bipred_idc = (uint8_t)(cfg == H264S_CFG_720P_MAIN ? (1 + pic_idx) % 3 : r.below(3));
    }
    const bool explicit_wp = (plan.type == H264R_P_SLICE && sl.weighted_pred_flag) ||
                             (plan.type == H264R_B_SLICE && sl.weighted_bipred_idc == 1);
    if (explicit_wp) {
        sl.luma_log2_weight_denom   = (uint8_t)(cfg == H264S_CFG_720P_MAIN && pic_idx < 8 ? 5 : r.range(0, 6));
        sl.chroma_log2_weight_denom = (uint8_t)r.range(0, 6);
        for (int list = 0; list < 2; ++list)
            for (int pl = 0; pl < 3; ++pl) {
                int d = pl ? sl.chroma_log2_weight_denom : sl.luma_log2_weight_denom;
                for (int i = 0; i < sl.num_ref[list]; ++i) {
                    int one = 1 << d;
                    if (r.chance(25)) { sl.wp_weight[list][pl][i] = (int8_t)one; sl.wp_offset[list][pl][i] = 0; }   // weight_flag = 0
                    else {
                        sl.wp_weight[list][pl][i] = (int8_t)r.range(one / 2, one + one / 2);
                        sl.wp_offset[list][pl][i] = (int8_t)r.range(-8, 8);
                        if (r.chance(5)) sl.wp_weight[list][pl][i] = (int8_t)(-sl.wp_weight[list][pl][i]);
                    }
                }
            }
    }
    if (plan.type == H264R_B_SLICE && sl.weighted_bipred_idc == 2) {
        sl.luma_log2_weight_denom = sl.chroma_log2_weight_denom = 5;     // interpret_rbsp.cc:721-724
        for (int i = 0; i < plan.n_l0; ++i)
            for (int j = 0; j < plan.n_l1; ++j) {
                int w0, w1;
                h264r_implicit_weights(plan.poc, s->plan[plan.l0[i]].poc, s->plan[plan.l1[j]].poc, 0, 0, &w0, &w1);
                sl.implicit_w1[i][j] = (int16_t)w1;
            }
    }
    // scaling lists -> InvLevelScale tables (Transform::init + set_quant)
    const int* q4[6]; const int* q8[2];
    if (s->default_matrices) {
        q4[0] = q4[1] = q4[2] = h264r::kDefault4x4Intra; q4[3] = q4[4] = q4[5] = h264r::kDefault4x4Inter;
        q8[0] = h264r::kDefault8x8Intra; q8[1] = h264r::kDefault8x8Inter;
    } else {
        for (int i = 0; i < 6; ++i) q4[i] = h264r::kFlat16;
        q8[0] = q8[1] = h264r::kFlat16;
    }
    h264r_build_level_scale(&sl, q4, q8);
}

} // namespace

extern "C" {

h264s_stream* h264s_open(int config, int stream_idx, int width_mbs, int height_mbs, int num_frames)
{
    if (config < H264S_CFG_CIF_BASELINE || config > H264S_CFG_1080I_FIELDS) return nullptr;
    h264s_stream* s = new h264s_stream();
    s->config = config; s->stream_idx = stream_idx;
    static const int dims[7][3] = { {0,0,0}, {22,18,30}, {80,45,24}, {120,68,16}, {240,135,8}, {120,68,16}, {120,34,16} };
    s->W = width_mbs > 0 ? width_mbs : dims[config][0];
    s->H = height_mbs > 0 ? height_mbs : dims[config][1];
    s->num_frames = num_frames > 0 ? num_frames : dims[config][2];
    s->rng.s = 0x4832363400000000ull + ((uint64_t)config << 16) + ((uint64_t)stream_idx << 8);
    s->allow_b = config != H264S_CFG_CIF_BASELINE;
    // direct_8x8_inference_flag = 0 (direct motion per 4x4 block, decoder.cc:239-242) on every third stream with B pictures
    s->direct8x8 = !(s->allow_b && stream_idx % 3 == 2);
    s->transform8x8 = config >= H264S_CFG_1080P_HIGH;
    s->default_matrices = (config == H264S_CFG_1080P_HIGH || config == H264S_CFG_MULTI_1080P || config == H264S_CFG_1080I_FIELDS) ? (stream_idx & 1)
                        : (config == H264S_CFG_4K_HIGH ? 1 : 0);
    s->chroma_qp_offset[0] = (config == H264S_CFG_CIF_BASELINE) ? ((stream_idx & 1) ? -2 : 0) : s->rng.range(-3, 3);
    s->chroma_qp_offset[1] = s->transform8x8 ? s->rng.range(-3, 3) : s->chroma_qp_offset[0];
    s->qp_lo = config == H264S_CFG_4K_HIGH ? 10 : 18;
    s->qp_hi = config == H264S_CFG_4K_HIGH ? 30 : 40;
    s->next_pic = 0;
    build_plan(s);
    return s;
}

void h264s_close(h264s_stream* s) { delete s; }

void h264s_get_seq(const h264s_stream* s, h264r_seq_params* sp, int* num_frames)
{
    memset(sp, 0, sizeof(*sp));
    sp->width_mbs = s->W; sp->height_mbs = s->H;
    sp->direct_8x8_inference_flag = s->direct8x8;
    sp->max_frames = 8; sp->max_pictures_in_flight = 4; sp->max_slices_per_picture = 4; sp->max_levels_per_picture = 0;
    if (num_frames) *num_frames = s->num_frames;
}

int h264s_next(h264s_stream* s, h264s_pic_info* info, h264r_pic_params* pp, h264r_mb* mbs,
               h264r_mb_motion* motion, h264r_slice* slices, h264r_level* levels, uint32_t level_capacity)
{
    if (s->next_pic >= s->num_frames) return 0;
    const int pic_idx = s->next_pic++;
    const PicPlan& plan = s->plan[pic_idx];
    const int W = s->W, H = s->H, nmb = W * H;
    Rng& r = s->rng;

    Gen g;
    g.s = s; g.rng = &r; g.W = W; g.H = H; g.plan = &plan;
    g.mbs = mbs; g.motion = motion; g.slices = slices;
    g.levels = levels; g.n_levels = 0; g.level_capacity = level_capacity; g.overflow = false; g.dense_used = false;
    memset(mbs, 0, sizeof(h264r_mb) * (size_t)nmb);

    // distinct reference pictures of this picture -> slots
    memset(info, 0, sizeof(*info));
    memset(pp, 0, sizeof(*pp));
    for (int i = 0; i < H264R_MAX_REFS; ++i) pp->ref_frames[i] = -1;
    int slot_l0[2] = {-1, -1}, slot_l1[1] = {-1};
    int nref = 0;
    auto slot_of = [&](int dec_idx) -> int {
        for (int i = 0; i < nref; ++i) if (info->ref_pic_index[i] == dec_idx) return i;
        info->ref_pic_index[nref] = dec_idx;
        pp->ref_poc[nref] = s->plan[dec_idx].poc;
        info->last_use_of_ref[nref] = s->last_use[dec_idx] == pic_idx;
        return nref++;
    };
    for (int i = 0; i < plan.n_l0; ++i) slot_l0[i] = slot_of(plan.l0[i]);
    for (int i = 0; i < plan.n_l1; ++i) slot_l1[i] = slot_of(plan.l1[i]);
    info->pic_index = pic_idx; info->pic_type = plan.type; info->used_for_reference = plan.is_ref;
    info->poc = plan.poc; info->num_refs = nref;
    pp->num_ref_frames = nref; pp->poc = plan.poc;
    pp->direct_8x8_inference_flag = s->direct8x8;
    // field pictures (config 6): parities alternate in display order, so the references of a picture are fields of both
    // parities (chroma vector offset, inter_prediction.cc:352-354)
    if (s->config == H264S_CFG_1080I_FIELDS) {
        pp->structure = ((plan.poc >> 1) & 1) ? H264R_BOTTOM_FIELD : H264R_TOP_FIELD;
        for (int i = 0; i < nref; ++i) pp->ref_structure[i] = (uint8_t)(((pp->ref_poc[i] >> 1) & 1) ? H264R_BOTTOM_FIELD : H264R_TOP_FIELD);
    }

    // slices: one, or two on odd pictures of the multi-slice configs (split not row aligned)
    g.num_slices = 1;
    if (s->config != H264S_CFG_CIF_BASELINE && (pic_idx & 1)) g.num_slices = 2;
    g.slice_first_mb[0] = 0;
    g.slice_first_mb[1] = g.num_slices > 1 ? r.range(nmb / 4, 3 * nmb / 4) : nmb;
    pp->num_slices = g.num_slices;
    int any_deblock = 0;
    for (int k = 0; k < g.num_slices; ++k) {
        fill_slice_tables(s, g, slices[k], plan, pic_idx, k, slot_l0, slot_l1);
        if (slices[k].disable_deblocking_filter_idc != 1) any_deblock = 1;
    }
    pp->run_deblock = any_deblock;

    g.qp = r.range(s->qp_lo, s->qp_hi);
    for (int addr = 0; addr < nmb; ++addr) {
        h264r_mb& mb = mbs[addr];
        mb.slice_idx = (uint16_t)((g.num_slices > 1 && addr >= g.slice_first_mb[1]) ? 1 : 0);
        g.dense_used = false;
        g.qp = clip3(s->qp_lo, s->qp_hi, g.qp + r.range(-2, 2));
        mb.qp_y = (int8_t)g.qp;
        for (int i = 0; i < 2; ++i)
            mb.qp_c[i] = (int8_t)h264r::qpc_from_qpi(clip3(0, 51, g.qp + s->chroma_qp_offset[i]));

        bool intra = plan.type == H264R_I_SLICE || plan.all_intra;
        if (!intra) intra = r.chance(plan.type == H264R_B_SLICE ? 5 : 10);
        if (intra) {
            int kind, t = r.below(100);
            if (s->config == H264S_CFG_CIF_BASELINE || s->config == H264S_CFG_720P_MAIN)
                kind = t < 59 ? H264R_MB_I4x4 : (t < 99 ? H264R_MB_I16x16 : H264R_MB_IPCM);
            else if (s->config == H264S_CFG_4K_HIGH)
                kind = t < 50 ? H264R_MB_I4x4 : (t < 75 ? H264R_MB_I8x8 : (t < 99 ? H264R_MB_I16x16 : H264R_MB_IPCM));
            else
                kind = t < 40 ? H264R_MB_I8x8 : (t < 70 ? H264R_MB_I4x4 : (t < 99 ? H264R_MB_I16x16 : H264R_MB_IPCM));
            clear_motion(motion[addr]);
            gen_intra_mb(g, addr, kind);
        } else {
            gen_inter_mb(g, addr);
        }
        emit_levels(g, mb);
    }
    info->num_levels = g.n_levels;
    return g.overflow ? -1 : 1;
}

void h264s_account(const h264r_mb* mbs, const h264r_slice* slices, int nmb, int run_deblock, uint64_t out[8])
{
    for (int i = 0; i < 8; ++i) out[i] = 0;
    for (int a = 0; a < nmb; ++a) {
        const h264r_mb& m = mbs[a];
        const bool intra = (m.flags & H264R_MB_FLAG_INTRA) != 0;
        uint64_t b = 32 + 384;
        if (m.coeff_count) { b += 4 * (uint64_t)m.coeff_count; out[6] += 1; }
        if (!intra) {
            b += 192;
            for (int q = 0; q < 4; ++q) b += (m.u.inter.sub_mb_pred_mode[q] == H264R_PRED_BI ? 2 : 1) * 96;
            out[1] += b; out[4] += 1;
        } else { out[2] += b; out[5] += 1; }
        out[0] += b;
        if (run_deblock && slices[m.slice_idx].disable_deblocking_filter_idc != 1) {
            out[3] += 32 + 384 + 384 + (intra ? 0 : 192);
            out[7] += 1;
        }
    }
}

} // extern "C"
