// Drives the host-side Decoder facade with exactly the calls the reference's parser makes (the same derivation as
// oracle/ref_harness.cc uses for the real reference Decoder) and checks that the buffers it fills are equal to the
// generator's picture description: headers, motion (unpacked again from the distinct entries the facade writes into the
// stream), cbp_blks and the per-MB multiset of levels.
#include "decoder_facade.h"
#include "h264synth.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

using namespace h264r;

int main(int argc, char** argv)
{
    const int config = argc > 1 ? atoi(argv[1]) : 3, W = argc > 2 ? atoi(argv[2]) : 12, H = argc > 3 ? atoi(argv[3]) : 8, N = argc > 4 ? atoi(argv[4]) : 6;
    h264s_stream* st = h264s_open(config, 0, W, H, N);
    const int nmb = W * H;
    std::vector<h264r_mb> mbs(nmb), fmbs(nmb);
    std::vector<h264r_mb_motion> motion(nmb);
    std::vector<h264r_slice> slices(4), fslices(4);
    std::vector<h264r_level> levels((size_t)nmb * 384), flevels((size_t)nmb * (384 + 48));
    ZigZag zz;
    h264s_pic_info info; h264r_pic_params pp;
    int pics = 0;
    while (h264s_next(st, &info, &pp, mbs.data(), motion.data(), slices.data(), levels.data(), (uint32_t)levels.size()) == 1) {
        h264r_pic_buffers bufs = { fmbs.data(), fslices.data(), flevels.data(), (uint32_t)flevels.size(), 0 };
        Decoder dec;
        const bool field = pp.structure != H264R_FRAME;       // field pictures: the parser's scan indexes are field-scan indexes
        const uint8_t* const X4 = zz.sx4(field); const uint8_t* const Y4 = zz.sy4(field);
        const uint8_t* const X8 = zz.sx8(field); const uint8_t* const Y8 = zz.sy8(field);
        dec.init(bufs, W, H, field);
        for (int addr = 0; addr < nmb; ++addr) {
            const h264r_mb& hm = mbs[addr];
            FacadeMb mb; memset(&mb, 0, sizeof(mb));
            mb.mbAddrX = addr; mb.is_intra_block = hm.flags & H264R_MB_FLAG_INTRA; mb.slice_nr = (short)hm.slice_idx;
            mb.mb_type = hm.mb_type; mb.transform_size_8x8_flag = hm.flags & H264R_MB_FLAG_T8x8;
            mb.intra_chroma_pred_mode = hm.chroma_mode; mb.Intra16x16PredMode = hm.intra16_mode;
            mb.CodedBlockPatternLuma = hm.cbp_luma; mb.CodedBlockPatternChroma = hm.cbp_chroma;
            mb.QpY = hm.qp_y; mb.QpC[0] = hm.qp_c[0]; mb.QpC[1] = hm.qp_c[1];
            if (mb.is_intra_block) {
                for (int i = 0; i < 16; ++i) mb.Intra4x4PredMode[i] = (hm.u.intra_modes[i >> 1] >> ((i & 1) * 4)) & 15;
                for (int i = 0; i < 4; ++i) mb.Intra8x8PredMode[i] = (hm.u.intra_modes[i >> 1] >> ((i & 1) * 4)) & 15;
            } else for (int i = 0; i < 4; ++i) { mb.SubMbType[i] = hm.u.inter.sub_mb_type[i]; mb.SubMbPredMode[i] = hm.u.inter.sub_mb_pred_mode[i]; }
            int16_t c[384]; memset(c, 0, sizeof(c));
            for (int i = 0; i < hm.coeff_count; ++i) { h264r_level e = levels[hm.coeff_offset + i]; c[H264R_LEVEL_POS(e)] = (int16_t)H264R_LEVEL_VALUE(e); }
            if (hm.mb_type == H264R_MB_IPCM) {
                for (int y = 0; y < 16; ++y) for (int x = 0; x < 16; ++x) dec.pcm_sample(&mb, PLANE_Y, x, y, c[y * 16 + x]);
                for (int pl = 1; pl <= 2; ++pl) for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) dec.pcm_sample(&mb, (ColorPlane)pl, x, y, c[256 + (pl - 1) * 64 + y * 8 + x]);
                mb.cbp_blks[0] = 0xFFFF;
            } else if (hm.coeff_count) {
                const bool i16 = hm.mb_type == H264R_MB_I16x16;
                if (i16) {
                    for (int k = 0; k < 16; ++k) { int lev = c[Y4[k] * 64 + X4[k] * 4]; if (lev) dec.coeff_luma_dc(&mb, PLANE_Y, 0, 0, k, lev); }
                    dec.transform_luma_dc(&mb, PLANE_Y);
                }
                for (int i8 = 0; i8 < 4; ++i8) {
                    if (!(hm.cbp_luma & (1 << i8))) continue;
                    int bx0 = (i8 & 1) * 2, by0 = (i8 >> 1) * 2;
                    if (mb.transform_size_8x8_flag) {
                        for (int k = 0; k < 64; ++k) { int lev = c[(by0 * 4 + Y8[k]) * 16 + bx0 * 4 + X8[k]]; if (lev) dec.coeff_luma_ac(&mb, PLANE_Y, bx0, by0, k, lev); }
                    } else for (int i4 = 0; i4 < 4; ++i4) {
                        int bx = bx0 + (i4 & 1), by = by0 + (i4 >> 1);
                        for (int k = i16 ? 1 : 0; k < 16; ++k) { int lev = c[(by * 4 + Y4[k]) * 16 + bx * 4 + X4[k]]; if (lev) dec.coeff_luma_ac(&mb, PLANE_Y, bx, by, k, lev); }
                    }
                }
                if (hm.cbp_chroma & 3) {
                    for (int pl = 1; pl <= 2; ++pl) {
                        const int16_t* cc = c + 256 + (pl - 1) * 64;
                        for (int k = 0; k < 4; ++k) { int lev = cc[(k >> 1) * 32 + (k & 1) * 4]; if (lev) dec.coeff_chroma_dc(&mb, (ColorPlane)pl, 0, 0, k, lev); }
                        dec.transform_chroma_dc(&mb, (ColorPlane)pl);
                    }
                    if (hm.cbp_chroma & 2)
                        for (int pl = 1; pl <= 2; ++pl) {
                            const int16_t* cc = c + 256 + (pl - 1) * 64;
                            for (int i4 = 0; i4 < 4; ++i4) for (int k = 1; k < 16; ++k) {
                                int lev = cc[((i4 >> 1) * 4 + Y4[k]) * 8 + (i4 & 1) * 4 + X4[k]];
                                if (lev) dec.coeff_chroma_ac(&mb, (ColorPlane)pl, i4 & 1, i4 >> 1, k, lev);
                            }
                        }
                }
            }
            FacadeMotion fm[16];
            for (int b = 0; b < 16; ++b) for (int l = 0; l < 2; ++l) {
                fm[b].ref_pic[l] = motion[addr].ref_pic[l][b]; fm[b].ref_idx[l] = motion[addr].ref_idx[l][b];
                fm[b].mv[l][0] = motion[addr].mv[l][b][0]; fm[b].mv[l][1] = motion[addr].mv[l][b][1];
            }
            dec.decode(mb, fm);
        }
        // compare
        for (int addr = 0; addr < nmb; ++addr) {
            h264r_mb a = mbs[addr], b = fmbs[addr];
            std::vector<h264r_level> la(levels.begin() + a.coeff_offset, levels.begin() + a.coeff_offset + a.coeff_count);
            std::vector<h264r_level> lb(flevels.begin() + b.coeff_offset, flevels.begin() + b.coeff_offset + b.coeff_count);
            std::sort(la.begin(), la.end()); std::sort(lb.begin(), lb.end());
            if (la != lb) { printf("picture %d MB %d: level lists differ (%zu vs %zu)\n", pics, addr, la.size(), lb.size()); return 1; }
            const uint32_t packed = b.motion;
            a.coeff_offset = b.coeff_offset = 0; a.motion = b.motion = 0;
            if (memcmp(&a, &b, sizeof(a))) { printf("picture %d MB %d: headers differ (type %d)\n", pics, addr, a.mb_type); return 1; }
            if (!(a.flags & H264R_MB_FLAG_INTRA)) {
                h264r_mb_motion fm;
                h264r_unpack_motion(flevels.data(), packed, &fm);
                if (memcmp(&motion[addr], &fm, sizeof(h264r_mb_motion))) { printf("picture %d MB %d: motion differs\n", pics, addr); return 1; }
            }
        }
        ++pics;
    }
    h264s_close(st);
    printf("facade round trip ok: %d pictures\n", pics);
    return pics > 0 ? 0 : 1;
}
