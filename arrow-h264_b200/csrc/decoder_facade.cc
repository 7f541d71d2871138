#include "decoder_facade.h"

#include <string.h>

namespace h264r {

void Decoder::init(const h264r_pic_buffers& bufs, int width_mbs, int height_mbs, bool field_pic_flag)
{
    bufs_ = bufs; width_mbs_ = width_mbs; height_mbs_ = height_mbs; field_ = field_pic_flag;
    n_levels_ = 0; cur_mb_ = -1; cur_first_ = 0; overflow_ = false;
}

void Decoder::assign_quant_params(int slice_nr, const int* const q4[6], const int* const q8[2])
{
    h264r_build_level_scale(&bufs_.slices[slice_nr], q4, q8);
}

void Decoder::append(FacadeMb* mb, int pos, int level)
{
    if (mb->mbAddrX != cur_mb_) { cur_mb_ = mb->mbAddrX; cur_first_ = n_levels_; }
    if (n_levels_ >= bufs_.stream_capacity) { overflow_ = true; return; }
    bufs_.stream[n_levels_++] = H264R_LEVEL(pos, level);
}

// Transform::coeff_luma_dc, transform.cc:425-429: cof[pos.y*4][pos.x*4] = level
void Decoder::coeff_luma_dc(FacadeMb* mb, ColorPlane, int, int, int runarr, int levarr)
{
    append(mb, (zz_.sy4(field_)[runarr] * 4) * 16 + zz_.sx4(field_)[runarr] * 4, levarr);
}

// Transform::coeff_luma_ac, transform.cc:431-440
void Decoder::coeff_luma_ac(FacadeMb* mb, ColorPlane pl, int x0, int y0, int runarr, int levarr)
{
    if (!mb->transform_size_8x8_flag) {
        mb->cbp_blks[pl] |= (uint64_t)0x01 << (y0 * 4 + x0);
        append(mb, (y0 * 4 + zz_.sy4(field_)[runarr]) * 16 + x0 * 4 + zz_.sx4(field_)[runarr], levarr);
    } else {
        mb->cbp_blks[pl] |= (uint64_t)0x33 << (y0 * 4 + x0);
        append(mb, (y0 * 4 + zz_.sy8(field_)[runarr]) * 16 + x0 * 4 + zz_.sx8(field_)[runarr], levarr);
    }
}

// Transform::coeff_chroma_dc, transform.cc:442-446 with inverse_scan_chroma_dc for ChromaArrayType 1 (:369-371)
void Decoder::coeff_chroma_dc(FacadeMb* mb, ColorPlane pl, int, int, int runarr, int levarr)
{
    append(mb, 256 + (pl - 1) * 64 + ((runarr / 2) * 4) * 8 + (runarr % 2) * 4, levarr);
}

// Transform::coeff_chroma_ac, transform.cc:448-456
void Decoder::coeff_chroma_ac(FacadeMb* mb, ColorPlane pl, int x0, int y0, int runarr, int levarr)
{
    append(mb, 256 + (pl - 1) * 64 + (y0 * 4 + zz_.sy4(field_)[runarr]) * 8 + x0 * 4 + zz_.sx4(field_)[runarr], levarr);
}

void Decoder::pcm_sample(FacadeMb* mb, ColorPlane pl, int x, int y, int value)
{
    append(mb, pl == PLANE_Y ? y * 16 + x : 256 + (pl - 1) * 64 + y * 8 + x, value);
}

void Decoder::decode(FacadeMb& mb, const FacadeMotion motion[16])
{
    h264r_mb& h = bufs_.mbs[mb.mbAddrX];
    memset(&h, 0, sizeof(h));
    h.mb_type = mb.mb_type;
    h.flags = (uint8_t)((mb.is_intra_block ? H264R_MB_FLAG_INTRA : 0) | (mb.transform_size_8x8_flag ? H264R_MB_FLAG_T8x8 : 0));
    h.slice_idx = (uint16_t)mb.slice_nr;
    h.cbp_luma = mb.CodedBlockPatternLuma; h.cbp_chroma = mb.CodedBlockPatternChroma;
    h.qp_y = mb.QpY; h.qp_c[0] = mb.QpC[0]; h.qp_c[1] = mb.QpC[1];
    h.intra16_mode = mb.Intra16x16PredMode; h.chroma_mode = mb.intra_chroma_pred_mode;
    h.cbp_blks = (uint16_t)mb.cbp_blks[0];
    const bool has = cur_mb_ == mb.mbAddrX;
    h.coeff_offset = has ? cur_first_ : n_levels_;
    h.coeff_count = has ? (uint16_t)(n_levels_ - cur_first_) : 0;
    if (mb.is_intra_block) {
        const uint8_t* modes = mb.mb_type == H264R_MB_I8x8 ? mb.Intra8x8PredMode : mb.Intra4x4PredMode;
        const int n = mb.mb_type == H264R_MB_I8x8 ? 4 : (mb.mb_type == H264R_MB_I4x4 ? 16 : 0);
        for (int i = 0; i < n; ++i) h.u.intra_modes[i >> 1] |= (uint8_t)((modes[i] & 15) << ((i & 1) * 4));
    } else {
        for (int i = 0; i < 4; ++i) { h.u.inter.sub_mb_type[i] = mb.SubMbType[i]; h.u.inter.sub_mb_pred_mode[i] = mb.SubMbPredMode[i]; }
    }
    if (mb.is_intra_block) return;
    // the MB's distinct motion entries go into the stream right behind its levels (the next MB's levels follow them)
    h264r_mb_motion m;
    for (int b = 0; b < 16; ++b)
        for (int list = 0; list < 2; ++list) {
            m.mv[list][b][0] = motion[b].mv[list][0]; m.mv[list][b][1] = motion[b].mv[list][1];
            m.ref_idx[list][b] = motion[b].ref_idx[list];
            m.ref_pic[list][b] = (int8_t)motion[b].ref_pic[list];
        }
    h264r_motion_entry e[16];
    const int code = h264r_pack_motion(&m, e);
    const uint32_t n = 3u * h264r_motion_entries_of_code[code];
    if (n_levels_ + n > bufs_.stream_capacity) { overflow_ = true; return; }
    memcpy(bufs_.stream + n_levels_, e, sizeof(uint32_t) * n);
    h.motion = n_levels_ << 4 | (uint32_t)code;
    n_levels_ += n;
    cur_mb_ = -1;                                   // levels of a later MB start a new run even if the address repeats
}

} // namespace h264r
