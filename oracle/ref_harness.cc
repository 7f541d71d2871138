// TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// Drives the reference's own `vio::h264::Decoder` (compiled unmodified from /root/reference by oracle/Makefile)
// with the neutral picture description of include/h264recon.h, following the recipe of SURVEY.md §8c:
// construct VideoParameters / sps_t / pps_t / slice_t / storable_picture / mb_t[] by hand (no parser),
// then per MB feed the levels through Decoder::coeff_* in the parser's order
// (parser/interpret_residual.cc:420-509), call Decoder::decode (core/slice_data.cc:646), and per picture
// Decoder::deblock_filter (framebuf/picture.cc:253) + pad_buf (framebuf/picture.cc:182-237).
// Only this file is ours; every line of arithmetic that produces a sample is the reference's.
#include "global.h"
#include "input_parameters.h"
#include "h264decoder.h"
#include "slice.h"
#include "sets.h"
#include "memalloc.h"
#include "macroblock.h"
#include "neighbour.h"
#include "decoder.h"
#include "dpb.h"

#include "h264recon.h"
#include "h264_tables.h"

#include <chrono>
#include <vector>

using vio::h264::mb_t;

extern void pad_buf(px_t* pImgBuf, int iWidth, int iHeight, int iStride, int iPadX, int iPadY);

struct ref_dec {
    VideoParameters* vid;
    sps_t sps;
    int W, H;
    std::vector<storable_picture*> frames;
    std::vector<mb_t> mb_data;
    h264r::ZigZag zz;
    uint8_t inv4[4][4], inv8[8][8];         // raster (y, x) -> scan index
};

extern "C" {

ref_dec* ref_open(const h264r_seq_params* sp)
{
    ref_dec* d = new ref_dec();
    d->W = sp->width_mbs; d->H = sp->height_mbs;
    d->vid = new VideoParameters;
    sps_t& sps = d->sps;
    sps = sps_t();
    sps.Valid = true;
    sps.profile_idc = 100; sps.level_idc = 51;
    sps.chroma_format_idc = 1; sps.separate_colour_plane_flag = 0;
    sps.bit_depth_luma_minus8 = 0; sps.bit_depth_chroma_minus8 = 0;
    sps.qpprime_y_zero_transform_bypass_flag = 0;
    sps.seq_scaling_matrix_present_flag = 0;
    sps.frame_mbs_only_flag = 1; sps.mb_adaptive_frame_field_flag = 0;
    sps.direct_8x8_inference_flag = sp->direct_8x8_inference_flag != 0;
    sps.pic_width_in_mbs_minus1 = d->W - 1; sps.pic_height_in_map_units_minus1 = d->H - 1;
    sps.ChromaArrayType = 1; sps.SubWidthC = 2; sps.SubHeightC = 2; sps.MbWidthC = 8; sps.MbHeightC = 8;
    sps.BitDepthY = 8; sps.BitDepthC = 8; sps.QpBdOffsetY = 0; sps.QpBdOffsetC = 0;
    sps.PicWidthInMbs = d->W; sps.PicWidthInSamplesL = d->W * 16; sps.PicWidthInSamplesC = d->W * 8;
    sps.PicHeightInMapUnits = d->H; sps.PicSizeInMapUnits = d->W * d->H; sps.FrameHeightInMbs = d->H;
    sps.max_num_ref_frames = 4;
    d->vid->active_sps = &d->sps;
    d->vid->structure = FRAME;
    d->mb_data.resize((size_t)d->W * d->H);
    d->vid->mb_data = d->mb_data.data();
    for (int k = 0; k < 16; ++k) d->inv4[d->zz.y4[k]][d->zz.x4[k]] = (uint8_t)k;
    for (int k = 0; k < 64; ++k) d->inv8[d->zz.y8[k]][d->zz.x8[k]] = (uint8_t)k;
    return d;
}

void ref_close(ref_dec* d)
{
    for (storable_picture* p : d->frames) delete p;
    // VideoParameters is leaked on purpose: its destructor walks decoder-global state we never set up.
    delete d;
}

int ref_frame_alloc(ref_dec* d)
{
    for (size_t i = 0; i < d->frames.size(); ++i)
        if (!d->frames[i]) {
            d->frames[i] = new storable_picture(d->vid, FRAME, d->W * 16, d->H * 16, d->W * 8, d->H * 8, 1);
            return (int)i;
        }
    d->frames.push_back(new storable_picture(d->vid, FRAME, d->W * 16, d->H * 16, d->W * 8, d->H * 8, 1));
    return (int)d->frames.size() - 1;
}

void ref_frame_release(ref_dec* d, int id)
{
    if (id >= 0 && id < (int)d->frames.size() && d->frames[id]) { delete d->frames[id]; d->frames[id] = nullptr; }
}

static void pad_frame(storable_picture* p)
{
    pad_buf(*p->imgY, p->size_x, p->size_y, p->iLumaStride, MCBUF_LUMA_PAD_X, MCBUF_LUMA_PAD_Y);
    pad_buf(*p->imgUV[0], p->size_x_cr, p->size_y_cr, p->iChromaStride, MCBUF_CHROMA_PAD_X, MCBUF_CHROMA_PAD_Y);
    pad_buf(*p->imgUV[1], p->size_x_cr, p->size_y_cr, p->iChromaStride, MCBUF_CHROMA_PAD_X, MCBUF_CHROMA_PAD_Y);
}

void ref_frame_get(ref_dec* d, int id, uint8_t* y, uint8_t* cb, uint8_t* cr)
{
    storable_picture* p = d->frames[id];
    for (int j = 0; j < p->size_y; ++j)
        for (int i = 0; i < p->size_x; ++i) y[(size_t)j * p->size_x + i] = (uint8_t)p->imgY[j][i];
    for (int j = 0; j < p->size_y_cr; ++j)
        for (int i = 0; i < p->size_x_cr; ++i) {
            cb[(size_t)j * p->size_x_cr + i] = (uint8_t)p->imgUV[0][j][i];
            cr[(size_t)j * p->size_x_cr + i] = (uint8_t)p->imgUV[1][j][i];
        }
}

void ref_frame_set(ref_dec* d, int id, const uint8_t* y, const uint8_t* cb, const uint8_t* cr)
{
    storable_picture* p = d->frames[id];
    for (int j = 0; j < p->size_y; ++j)
        for (int i = 0; i < p->size_x; ++i) p->imgY[j][i] = y[(size_t)j * p->size_x + i];
    for (int j = 0; j < p->size_y_cr; ++j)
        for (int i = 0; i < p->size_x_cr; ++i) {
            p->imgUV[0][j][i] = cb[(size_t)j * p->size_x_cr + i];
            p->imgUV[1][j][i] = cr[(size_t)j * p->size_x_cr + i];
        }
    pad_frame(p);
}

// returns 0 on success, >0 = number of MBs whose cbp_blks (set by the reference's coeff_luma_ac) differ from the
// description's -- a consistency check of the generator/facade, not of the reconstruction.
int ref_reconstruct(ref_dec* d, int dst, const h264r_pic_params* pp, int used_for_reference,
                    const h264r_slice* slices, const h264r_mb* mbs, const h264r_mb_motion* motion,
                    const h264r_level* levels, double* sec_decode, double* sec_deblock)
{
    VideoParameters* vid = d->vid;
    const int W = d->W, H = d->H, nmb = W * H;
    storable_picture* pic = d->frames[dst];
    // A field picture (field_pic_flag = 1) is a picture of its own of half the frame height: d->H is the height of the
    // field, every storable_picture of the pool holds one field (what the reference's own field pictures are).
    const bool field = pp->structure != H264R_FRAME;
    const PictureStructure structure = pp->structure == H264R_TOP_FIELD ? TOP_FIELD : (pp->structure == H264R_BOTTOM_FIELD ? BOTTOM_FIELD : FRAME);
    d->sps.frame_mbs_only_flag = !field;
    vid->structure = structure;
    pic->slice.structure = structure;
    const uint8_t* const X4 = d->zz.sx4(field); const uint8_t* const Y4 = d->zz.sy4(field);     // transform.cc:344-382
    const uint8_t* const X8 = d->zz.sx8(field); const uint8_t* const Y8 = d->zz.sy8(field);
    pic->slice_headers.clear();
    pic->sps = &d->sps;
    pic->poc = pic->frame_poc = pic->top_poc = pic->bottom_poc = pp->poc;
    pic->used_for_reference = used_for_reference;
    pic->is_long_term = 0;
    vid->dec_picture = pic;
    // calloc state of a fresh storable_picture (framebuf/memalloc.cc:38-49), since frames are recycled here
    memset(&pic->mv_info[0][0], 0, sizeof(pic_motion_params) * (size_t)(W * 4) * (H * 4));

    for (int i = 0; i < pp->num_ref_frames; ++i) {
        storable_picture* r = d->frames[pp->ref_frames[i]];
        r->poc = pp->ref_poc[i];
        r->is_long_term = pp->ref_long_term[i];
    }

    std::vector<slice_t*> sl(pp->num_slices);
    std::vector<pps_t> ppss(pp->num_slices);
    for (int k = 0; k < pp->num_slices; ++k) {
        const h264r_slice& hs = slices[k];
        pps_t& pps = ppss[k];
        pps = pps_t();
        pps.Valid = true;
        pps.weighted_pred_flag = hs.weighted_pred_flag != 0;
        pps.weighted_bipred_idc = hs.weighted_bipred_idc;
        pps.constrained_intra_pred_flag = hs.constrained_intra_pred_flag != 0;
        pps.transform_8x8_mode_flag = 1;
        // weightScale lists recovered from LevelScale = normAdjust * weightScale (exact division), so that the
        // reference rebuilds InvLevelScale itself in Transform::init/set_quant (transform.cc:173-302)
        pps.pic_scaling_matrix_present_flag = 1;
        for (int i = 0; i < 8; ++i) pps.pic_scaling_list_present_flag[i] = 1;
        for (int i = 0; i < 6; ++i) {
            pps.UseDefaultScalingMatrix4x4Flag[i] = 0;
            for (int j = 0; j < 4; ++j)
                for (int x = 0; x < 4; ++x)
                    pps.ScalingList4x4[i][j * 4 + x] =
                        hs.level_scale_4x4[i / 3][i % 3][0][j * 4 + x] / h264r::norm_adjust_4x4(0, x, j);
        }
        for (int i = 0; i < 2; ++i) {
            pps.UseDefaultScalingMatrix8x8Flag[i] = 0;
            for (int j = 0; j < 8; ++j)
                for (int x = 0; x < 8; ++x)
                    pps.ScalingList8x8[i][j * 8 + x] =
                        hs.level_scale_8x8[i][0][j * 8 + x] / h264r::norm_adjust_8x8(0, x, j);
        }

        slice_t* s = new slice_t;
        sl[k] = s;
        s->p_Vid = vid; s->p_Dpb = nullptr;
        s->active_sps = &d->sps; s->active_pps = &pps;
        s->dec_picture = pic;
        s->neighbour.mb_data = d->mb_data.data();
        s->current_slice_nr = (short)k;
        s->layer_id = 0; s->view_id = 0;
        shr_t& shr = s->header;
        shr.slice_type = hs.slice_type;
        shr.field_pic_flag = field; shr.bottom_field_flag = structure == BOTTOM_FIELD; shr.MbaffFrameFlag = 0; shr.structure = structure;
        shr.colour_plane_id = 0;
        shr.direct_spatial_mv_pred_flag = hs.direct_spatial_mv_pred_flag != 0;
        shr.disable_deblocking_filter_idc = hs.disable_deblocking_filter_idc;
        shr.FilterOffsetA = hs.filter_offset_a; shr.FilterOffsetB = hs.filter_offset_b;
        shr.luma_log2_weight_denom = hs.luma_log2_weight_denom;
        shr.chroma_log2_weight_denom = hs.chroma_log2_weight_denom;
        shr.PicHeightInMbs = H; shr.PicHeightInSamplesL = H * 16; shr.PicHeightInSamplesC = H * 8;
        shr.PicSizeInMbs = nmb;
        shr.PicOrderCnt = shr.TopFieldOrderCnt = shr.BottomFieldOrderCnt = pp->poc;
        shr.num_ref_idx_l0_active_minus1 = hs.num_ref[0] ? hs.num_ref[0] - 1 : 0;
        shr.num_ref_idx_l1_active_minus1 = hs.num_ref[1] ? hs.num_ref[1] - 1 : 0;
        for (int list = 0; list < 2; ++list) {
            s->RefPicSize[list] = (char)hs.num_ref[list];
            for (int i = 0; i < H264R_MAX_REFS; ++i)
                s->RefPicList[list][i] = hs.ref_pic_list[list][i] >= 0 ? d->frames[pp->ref_frames[hs.ref_pic_list[list][i]]] : nullptr;
            for (int pl = 0; pl < 3; ++pl) {
                shr.pred_weight_l[list][pl].resize(H264R_MAX_REFS);
                for (int i = 0; i < H264R_MAX_REFS; ++i) {
                    shr.pred_weight_l[list][pl][i].weight_flag = true;
                    shr.pred_weight_l[list][pl][i].weight = hs.wp_weight[list][pl][i];
                    shr.pred_weight_l[list][pl][i].offset = hs.wp_offset[list][pl][i];
                }
            }
        }
        vid->active_pps = &pps;
        s->decoder.init(*s);
        s->decoder.assign_quant_params(*s);
        pic->slice_headers.push_back(s);
    }
    pic->pps = sl[0]->active_pps;

    // init_picture: every MB starts unavailable (core/slice_data.cc:53-58, 261-271)
    for (int i = 0; i < nmb; ++i) { d->mb_data[i].slice_nr = -1; d->mb_data[i].ei_flag = 1; d->mb_data[i].dpl_flag = 0; }

    int cbp_mismatch = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int addr = 0; addr < nmb; ++addr) {
        const h264r_mb& hm = mbs[addr];
        mb_t& mb = d->mb_data[addr];
        slice_t& s = *sl[hm.slice_idx];
        vio::h264::Decoder& dec = s.decoder;
        mb.p_Slice = &s; mb.mbAddrX = addr; mb.mb.x = addr % W; mb.mb.y = addr / W;
        mb.slice_nr = (short)hm.slice_idx; mb.ei_flag = 0; mb.dpl_flag = 0;
        mb.is_intra_block = (hm.flags & H264R_MB_FLAG_INTRA) != 0;
        mb.mb_skip_flag = 0; mb.mb_field_decoding_flag = 0;
        mb.mb_type = hm.mb_type;
        mb.transform_size_8x8_flag = (hm.flags & H264R_MB_FLAG_T8x8) != 0;
        mb.intra_chroma_pred_mode = hm.chroma_mode;
        mb.Intra16x16PredMode = hm.intra16_mode;
        mb.CodedBlockPatternLuma = hm.cbp_luma; mb.CodedBlockPatternChroma = hm.cbp_chroma;
        mb.QpY = hm.qp_y; mb.QpC[0] = hm.qp_c[0]; mb.QpC[1] = hm.qp_c[1];
        mb.qp_scaled[0] = hm.qp_y; mb.qp_scaled[1] = hm.qp_c[0]; mb.qp_scaled[2] = hm.qp_c[1];
        mb.TransformBypassModeFlag = 0;
        memset(mb.cbp_blks, 0, sizeof(mb.cbp_blks));
        if (mb.is_intra_block) {
            for (int i = 0; i < 16; ++i) mb.Intra4x4PredMode[i] = (hm.u.intra_modes[i >> 1] >> ((i & 1) * 4)) & 15;
            for (int i = 0; i < 4; ++i)  mb.Intra8x8PredMode[i] = (hm.u.intra_modes[i >> 1] >> ((i & 1) * 4)) & 15;
        } else {
            for (int i = 0; i < 4; ++i) { mb.SubMbType[i] = hm.u.inter.sub_mb_type[i]; mb.SubMbPredMode[i] = hm.u.inter.sub_mb_pred_mode[i]; }
        }
        // motion (written by the parser before Decoder::decode)
        const h264r_mb_motion& mm = motion[addr];
        for (int b = 0; b < 16; ++b) {
            pic_motion_params& mv = pic->mv_info[mb.mb.y * 4 + (b >> 2)][mb.mb.x * 4 + (b & 3)];
            mv.slice_no = (uint8_t)hm.slice_idx;
            for (int list = 0; list < 2; ++list) {
                if (mb.is_intra_block) { mv.ref_pic[list] = nullptr; mv.ref_idx[list] = -1; mv.mv[list] = {0, 0}; continue; }
                mv.ref_pic[list] = mm.ref_pic[list][b] >= 0 ? d->frames[pp->ref_frames[mm.ref_pic[list][b]]] : nullptr;
                mv.ref_idx[list] = mm.ref_idx[list][b];
                mv.mv[list] = { mm.mv[list][b][0], mm.mv[list][b][1] };
            }
        }
        // coefficients: mb_t::init zeroes Transform::cof (core/slice_data.cc:496-503)
        memset(dec.transform->cof, 0, sizeof(dec.transform->cof));
        if (hm.coeff_count) {
            int16_t c[H264R_COEFFS_PER_MB];
            memset(c, 0, sizeof(c));
            for (int i = 0; i < hm.coeff_count; ++i) {
                const h264r_level e = levels[hm.coeff_offset + i];
                c[H264R_LEVEL_POS(e)] = (int16_t)H264R_LEVEL_VALUE(e);
            }
            if (hm.mb_type == H264R_MB_IPCM) {                          // parse_i_pcm, interpret_mb.cc:405-470
                for (int y = 0; y < 16; ++y) for (int x = 0; x < 16; ++x) dec.transform->cof[0][y][x] = c[y * 16 + x];
                for (int pl = 0; pl < 2; ++pl)
                    for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) dec.transform->cof[pl + 1][y][x] = c[256 + pl * 64 + y * 8 + x];
                mb.cbp_blks[0] = 0xFFFF;
            } else {
                const bool i16 = hm.mb_type == H264R_MB_I16x16;
                if (i16) {                                              // residual_luma, interpret_residual.cc:424-431
                    for (int k = 0; k < 16; ++k) {
                        int bx = X4[k], by = Y4[k];
                        int lev = c[by * 4 * 16 + bx * 4];
                        if (lev) dec.coeff_luma_dc(&mb, PLANE_Y, 0, 0, k, lev);
                    }
                    dec.transform_luma_dc(&mb, PLANE_Y);
                }
                for (int i8 = 0; i8 < 4; ++i8) {
                    if (!(hm.cbp_luma & (1 << i8))) continue;
                    int bx0 = (i8 & 1) * 2, by0 = (i8 >> 1) * 2;
                    if (mb.transform_size_8x8_flag) {
                        for (int k = 0; k < 64; ++k) {
                            int lev = c[(by0 * 4 + Y8[k]) * 16 + bx0 * 4 + X8[k]];
                            if (lev) dec.coeff_luma_ac(&mb, PLANE_Y, bx0, by0, k, lev);
                        }
                    } else {
                        for (int i4 = 0; i4 < 4; ++i4) {
                            int bx = bx0 + (i4 & 1), by = by0 + (i4 >> 1);
                            for (int k = i16 ? 1 : 0; k < 16; ++k) {
                                int lev = c[(by * 4 + Y4[k]) * 16 + bx * 4 + X4[k]];
                                if (lev) dec.coeff_luma_ac(&mb, PLANE_Y, bx, by, k, lev);
                            }
                        }
                    }
                }
                if (hm.cbp_chroma & 3) {                                 // residual_chroma, interpret_residual.cc:468-494
                    for (int pl = 0; pl < 2; ++pl) {
                        const int16_t* cc = c + 256 + pl * 64;
                        for (int k = 0; k < 4; ++k) {
                            int lev = cc[(k >> 1) * 4 * 8 + (k & 1) * 4];
                            if (lev) dec.coeff_chroma_dc(&mb, (ColorPlane)(pl + 1), 0, 0, k, lev);
                        }
                        dec.transform_chroma_dc(&mb, (ColorPlane)(pl + 1));
                    }
                    if (hm.cbp_chroma & 2)
                        for (int pl = 0; pl < 2; ++pl) {
                            const int16_t* cc = c + 256 + pl * 64;
                            for (int i4 = 0; i4 < 4; ++i4) {
                                int bx = i4 & 1, by = i4 >> 1;
                                for (int k = 1; k < 16; ++k) {
                                    int lev = cc[(by * 4 + Y4[k]) * 8 + bx * 4 + X4[k]];
                                    if (lev) dec.coeff_chroma_ac(&mb, (ColorPlane)(pl + 1), bx, by, k, lev);
                                }
                            }
                        }
                }
            }
        } else if (hm.mb_type == H264R_MB_I16x16) {
            dec.transform_luma_dc(&mb, PLANE_Y);
        }
        if ((uint16_t)mb.cbp_blks[0] != hm.cbp_blks) ++cbp_mismatch;
        dec.decode(mb);
    }
    auto t1 = std::chrono::steady_clock::now();
    sl[0]->decoder.deblock_filter(*sl[0]);
    auto t2 = std::chrono::steady_clock::now();
    if (used_for_reference) pad_frame(pic);
    if (sec_decode)  *sec_decode  = std::chrono::duration<double>(t1 - t0).count();
    if (sec_deblock) *sec_deblock = std::chrono::duration<double>(t2 - t1).count();

    pic->slice_headers.clear();
    for (slice_t* s : sl) delete s;
    vid->dec_picture = nullptr;
    return cbp_mismatch;
}

} // extern "C"
