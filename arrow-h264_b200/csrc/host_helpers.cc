// Host-side helpers of the reconstruction boundary that restate what the reference computes per slice header
// (not per sample): implicit bi-prediction weights and the InvLevelScale tables.
#include "h264recon.h"
#include "h264_tables.h"

#include <stdlib.h>
#include <string.h>

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

extern "C" {

const uint8_t h264r_motion_entries_of_code[6] = { 0, 1, 2, 2, 4, 16 };

int h264r_pack_motion(const h264r_mb_motion* m, h264r_motion_entry out[16])
{
    h264r_motion_entry e[16];
    for (int b = 0; b < 16; ++b) {
        memcpy(e[b].mv[0], m->mv[0][b], 4); memcpy(e[b].mv[1], m->mv[1][b], 4);
        e[b].ref_idx[0] = m->ref_idx[0][b]; e[b].ref_idx[1] = m->ref_idx[1][b];
        e[b].ref_pic[0] = m->ref_pic[0][b]; e[b].ref_pic[1] = m->ref_pic[1][b];
    }
    auto same = [&](int a, int b) { return memcmp(&e[a], &e[b], sizeof(h264r_motion_entry)) == 0; };
    bool quad = true;
    for (int q = 0; q < 4 && quad; ++q) {
        const int b0 = (q >> 1) * 8 + (q & 1) * 2;
        quad = same(b0, b0 + 1) && same(b0, b0 + 4) && same(b0, b0 + 5);
    }
    if (!quad) { memcpy(out, e, sizeof(e)); return 5; }
    const bool top = same(0, 2), bottom = same(8, 10), left = same(0, 8), right = same(2, 10);
    out[0] = e[0];
    if (top && bottom && left) return 1;
    if (top && bottom) { out[1] = e[8]; return 2; }
    if (left && right) { out[1] = e[2]; return 3; }
    out[1] = e[2]; out[2] = e[8]; out[3] = e[10];
    return 4;
}

void h264r_unpack_motion(const uint32_t* stream, uint32_t motion, h264r_mb_motion* out)
{
    const int code = (int)(motion & 15);
    const h264r_motion_entry* e = reinterpret_cast<const h264r_motion_entry*>(stream + (motion >> 4));
    for (int b = 0; b < 16; ++b) {
        const int row2 = b >> 3, col2 = (b >> 1) & 1;
        const int k = code == 5 ? b : ((code == 2 || code == 4) ? row2 << (code == 4) : 0) + ((code == 3 || code == 4) ? col2 : 0);
        for (int l = 0; l < 2; ++l) {
            out->mv[l][b][0] = e[k].mv[l][0]; out->mv[l][b][1] = e[k].mv[l][1];
            out->ref_idx[l][b] = e[k].ref_idx[l]; out->ref_pic[l][b] = e[k].ref_pic[l];
        }
    }
}

int64_t h264r_pack_picture(int num_mbs, const h264r_mb* mbs, const h264r_mb_motion* motion, const h264r_level* levels,
                           uint32_t num_levels, h264r_mb* out_mbs, uint32_t* stream, uint32_t stream_capacity)
{
    if (num_mbs <= 0 || !mbs || !motion || !out_mbs || !stream || (num_levels && !levels)) return H264R_ERR_INVALID;
    if (num_levels > stream_capacity) return H264R_ERR_NOMEM;
    // the levels keep their indexes (h264r_mb::coeff_offset stays valid); the motion entries follow them
    if (num_levels) memcpy(stream, levels, sizeof(h264r_level) * (size_t)num_levels);
    if (out_mbs != mbs) memcpy(out_mbs, mbs, sizeof(h264r_mb) * (size_t)num_mbs);
    uint32_t words = num_levels;
    for (int i = 0; i < num_mbs; ++i) {
        h264r_mb& m = out_mbs[i];
        if (m.flags & H264R_MB_FLAG_INTRA) { m.motion = 0; continue; }
        h264r_motion_entry e[16];
        const int code = h264r_pack_motion(&motion[i], e);
        const uint32_t n = 3u * h264r_motion_entries_of_code[code];
        if (words + n > stream_capacity) return H264R_ERR_NOMEM;
        memcpy(stream + words, e, sizeof(uint32_t) * n);
        m.motion = words << 4 | (uint32_t)code;
        words += n;
    }
    return (int64_t)words;
}

void h264r_implicit_weights(int cur_poc, int poc0, int poc1, int long_term0, int long_term1, int* w0, int* w1)
{
    // restates inter_prediction.cc:112-139 for frame pictures
    int td = clip3(-128, 127, poc1 - poc0);
    if (td == 0 || long_term0 || long_term1) { *w0 = 32; *w1 = 32; return; }
    int tb = clip3(-128, 127, cur_poc - poc0);
    int tx = (16384 + abs(td / 2)) / td;
    int dsf = clip3(-1024, 1023, (tx * tb + 32) >> 6);
    *w1 = dsf >> 2;
    *w0 = 64 - *w1;
    if (*w1 < -64 || *w1 > 128) { *w0 = 32; *w1 = 32; }
}

void h264r_build_level_scale(h264r_slice* s, const int* const q4[6], const int* const q8[2])
{
    // restates Transform::set_quant (transform.cc:265-302): LevelScale = normAdjust * weightScale
    for (int inter = 0; inter < 2; ++inter)
        for (int pl = 0; pl < 3; ++pl)
            for (int k = 0; k < 6; ++k)
                for (int j = 0; j < 4; ++j)
                    for (int i = 0; i < 4; ++i)
                        s->level_scale_4x4[inter][pl][k][j * 4 + i] =
                            (uint16_t)(h264r::norm_adjust_4x4(k, i, j) * q4[inter * 3 + pl][j * 4 + i]);
    for (int inter = 0; inter < 2; ++inter)
        for (int k = 0; k < 6; ++k)
            for (int j = 0; j < 8; ++j)
                for (int i = 0; i < 8; ++i)
                    s->level_scale_8x8[inter][k][j * 8 + i] =
                        (uint16_t)(h264r::norm_adjust_8x8(k, i, j) * q8[inter][j * 8 + i]);
}

} // extern "C"
