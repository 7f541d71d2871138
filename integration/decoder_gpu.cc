// vio::h264::Decoder implemented on libh264recon.so -- the binding INTEGRATION.md describes, as compilable code.
//
// This file is compiled against the reference's own headers, UNCHANGED, where they lie under /root/reference
// (integration/Makefile; nothing of the reference is copied into this repository) and linked with the reference's
// own core/, parser/ and framebuf/ objects in place of its decoder/{decoder,transform,inter_prediction,
// intra_prediction,deblock}.cc.  The result is the reference's `ldecod` with the reconstruction on the GPU:
// NAL/slice parsing, CAVLC/CABAC, motion-vector prediction, the DPB and the YUV writer are the reference's, every
// sample is produced by the CUDA engine.  tests/test_integration_gpu.py decodes a bitstream with it and with the
// unmodified reference decoder and requires byte-identical output.
//
// Stage 2 of SURVEY.md 8f: the DPB is device resident.  Every storable_picture owns one engine frame; nothing is
// copied into storable_picture::imgY/imgUV (the reference still allocates and pads them, nobody reads them: motion
// compensation runs on the GPU, output goes through output_gpu.cc, which replaces framebuf/output.cc and downloads the
// cropped display rectangle when the DPB releases a picture).  deblock_filter() returns as soon as the picture is
// queued, so parsing picture N+1 overlaps the reconstruction of picture N.  Host-side readers of imgY that remain in
// the reference (error concealment, PSNR against a reference YUV) are outside the supported subset.
#include "global.h"
#include "dpb.h"
#include "slice.h"
#include "macroblock.h"
#include "decoder.h"

#include "h264recon.h"
#include "decoder_facade.h"
#include "h264_tables.h"

#include <algorithm>
#include <unordered_map>
#include <vector>

namespace vio  {
namespace h264 {

namespace {

const int Flat16[64] = {
    16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,
    16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16,16 };

void check(int rc, const char* what)
{
    if (rc != H264R_OK) error(500, "h264recon: %s: %s", what, h264r_strerror(rc));
}

// the twelve weightScale lists Transform::init picks (transform.cc:173-262), kept per slice until Decoder::init
struct QuantLists { int q4[6][16]; int q8[2][64]; };

// Engine frame of a decoded picture.  The map below is keyed by the storable_picture's ADDRESS, which is only an
// identity: entries are never dereferenced (the DPB frees pictures behind our back), what the engine wants to know about a
// reference (its POC) is cached here when the picture is decoded.
struct GpuFrameEntry { h264r_frame frame; int poc; bool live; uint8_t structure; };

struct GpuState {
    // [0]: the context of the stream's frame pictures, [1]: of its field pictures (half the height; a field is a picture of
    // its own, like the reference's field storable_pictures).  A PAFF stream uses both; references that exist in the other
    // form only are converted on the device when a picture needs them (convert_references).
    h264r_ctx* ctx[2] = { nullptr, nullptr };
    int kind = 0;                          // context of the picture being parsed
    int width_mbs = 0, height_mbs = 0;     // FrameHeightInMbs
    std::unordered_map<const storable_picture*, GpuFrameEntry> frames;   // engine frame of every decoded picture
    std::unordered_map<const Decoder*, QuantLists> quant;                 // assign_quant_params precedes init
    // picture being parsed
    const storable_picture* cur = nullptr;
    h264r_pic_params pp;
    h264r_pic_buffers bufs;
    h264r::Decoder facade;
    bool any_deblock = false;
} g;

const int kMaxGpuFrames = 40;        // > 16 DPB frames + the current picture + pictures waiting for output

h264r_ctx* ctx_of(int structure) { return g.ctx[structure != H264R_FRAME]; }

void open_engine(const sps_t& sps, const shr_t& shr)
{
    const int W = (int)sps.PicWidthInMbs, H = (int)sps.FrameHeightInMbs;
    if (W != g.width_mbs || H != g.height_mbs) {
        for (h264r_ctx*& c : g.ctx) if (c) { h264r_destroy(c); c = nullptr; }
        g.frames.clear();
        g.width_mbs = W; g.height_mbs = H;
    }
    g.kind = shr.field_pic_flag ? 1 : 0;
    if (g.ctx[g.kind]) return;
    if (sps.chroma_format_idc != 1 || sps.bit_depth_luma_minus8 != 0 || sps.bit_depth_chroma_minus8 != 0 || sps.mb_adaptive_frame_field_flag)
        error(500, "h264recon: %s", h264r_strerror(H264R_ERR_UNSUPPORTED));
    h264r_seq_params sp;
    memset(&sp, 0, sizeof(sp));
    sp.width_mbs = W; sp.height_mbs = g.kind ? H / 2 : H;
    sp.direct_8x8_inference_flag = sps.direct_8x8_inference_flag;
    sp.max_frames = kMaxGpuFrames; sp.max_pictures_in_flight = 4; sp.max_slices_per_picture = 64;   // 4 staging slots: up to three pictures reconstruct while the next is parsed
    sp.max_levels_per_picture = 0;
    check(h264r_create(&g.ctx[g.kind], 0, &sp), "h264r_create");
}

// engine frame of a decoded picture
h264r_frame frame_of(const storable_picture* p)
{
    auto it = g.frames.find(p);
    if (it == g.frames.end()) error(500, "h264recon: reference picture was not reconstructed by the GPU path");
    if ((it->second.structure != H264R_FRAME) != (g.kind != 0)) error(500, "h264recon: reference picture is held in the other form (frame / field)");
    return it->second.frame;
}

// Liveness follows the DPB, not a guess about use: a picture keeps its engine frame while the decoded picture buffer holds
// it (as a reference or waiting for output -- every frame store of every layer) and is released at the first picture start
// after the DPB dropped it.  Pictures that never enter the DPB (direct output, framebuf/dpb.cc:383-385) are released by
// gpu_picture_freed() right after they were written.  Only the map keys are compared, nothing is dereferenced but the
// DPB's own live frame stores.
void sweep_dead_frames(VideoParameters* p_Vid, const storable_picture* current)
{
    for (auto& kv : g.frames) kv.second.live = kv.first == current;
    for (int layer = 0; layer < MAX_NUM_DPB_LAYERS; ++layer) {
        const decoded_picture_buffer_t* dpb = p_Vid->p_Dpb_layer[layer];
        if (!dpb || !dpb->fs) continue;
        for (unsigned i = 0; i < dpb->used_size; ++i) {
            const pic_t* fs = dpb->fs[i];
            if (!fs) continue;
            for (const storable_picture* p : { (const storable_picture*)fs->frame, (const storable_picture*)fs->top_field, (const storable_picture*)fs->bottom_field }) {
                auto it = g.frames.find(p);
                if (it != g.frames.end()) it->second.live = true;
            }
        }
    }
    if (p_Vid->out_buffer)                                  // fields waiting for their other half in direct_output (output_gpu.cc)
        for (const storable_picture* p : { (const storable_picture*)p_Vid->out_buffer->top_field, (const storable_picture*)p_Vid->out_buffer->bottom_field }) {
            auto it = g.frames.find(p);
            if (it != g.frames.end()) it->second.live = true;
        }
    for (auto it = g.frames.begin(); it != g.frames.end(); ) {
        if (it->second.live) { ++it; continue; }
        h264r_frame_release(ctx_of(it->second.structure), it->second.frame);
        it = g.frames.erase(it);
    }
}

// a new picture: its storable_picture may reuse the address of a freed one
h264r_frame new_frame(const storable_picture* p, int poc, int structure)
{
    auto it = g.frames.find(p);
    if (it != g.frames.end()) { h264r_frame_release(ctx_of(it->second.structure), it->second.frame); g.frames.erase(it); }
    int same_form = 0;
    for (auto& kv : g.frames) same_form += (kv.second.structure != H264R_FRAME) == (structure != H264R_FRAME);
    if (same_form >= H264R_MAX_REFS) error(500, "h264recon: more than %d pictures of one form alive in the decoded picture buffer", H264R_MAX_REFS);
    h264r_frame f;
    check(h264r_frame_alloc(ctx_of(structure), &f), "h264r_frame_alloc");
    g.frames[p] = GpuFrameEntry{ f, poc, true, (uint8_t)structure };
    return f;
}

int slot_of(h264r_frame f)
{
    for (int i = 0; i < g.pp.num_ref_frames; ++i) if (g.pp.ref_frames[i] == f) return i;
    error(500, "h264recon: reference picture missing from the picture's reference table");
    return -1;
}

// PAFF: the DPB keeps every complete frame in both forms (picture_t::insert_picture -> dpb_split_field / dpb_combine_field,
// framebuf/picture.cc:408-745, on host arrays the GPU binding never fills).  On the device a picture exists in the form it was
// decoded in; when the picture about to be decoded is of the other kind, the reference frames of the DPB are converted:
// h264r_field_copy, device to device, asynchronous.
void convert_references(VideoParameters* p_Vid)
{
    auto has = [](const storable_picture* p) { return p && g.frames.find(p) != g.frames.end(); };
    for (int layer = 0; layer < MAX_NUM_DPB_LAYERS; ++layer) {
        const decoded_picture_buffer_t* dpb = p_Vid->p_Dpb_layer[layer];
        if (!dpb || !dpb->fs) continue;
        for (unsigned i = 0; i < dpb->used_size; ++i) {
            const pic_t* fs = dpb->fs[i];
            if (!fs || fs->is_used != 3 || !fs->is_reference) continue;
            if (g.kind == 1 && has(fs->frame)) {
                const h264r_frame src = g.frames[fs->frame].frame;
                for (int parity = 0; parity < 2; ++parity) {
                    const storable_picture* fld = parity ? fs->bottom_field : fs->top_field;
                    if (!fld || has(fld)) continue;
                    h264r_frame f;
                    check(h264r_frame_alloc(g.ctx[1], &f), "h264r_frame_alloc");
                    check(h264r_field_copy(g.ctx[0], src, g.ctx[1], f, parity, 1), "h264r_field_copy");
                    g.frames[fld] = GpuFrameEntry{ f, fld->poc, true, (uint8_t)(parity ? H264R_BOTTOM_FIELD : H264R_TOP_FIELD) };
                }
            } else if (g.kind == 0 && fs->frame && !has(fs->frame) && has(fs->top_field) && has(fs->bottom_field)) {
                h264r_frame f;
                check(h264r_frame_alloc(g.ctx[0], &f), "h264r_frame_alloc");
                check(h264r_field_copy(g.ctx[0], f, g.ctx[1], g.frames[fs->top_field].frame, 0, 0), "h264r_field_copy");
                check(h264r_field_copy(g.ctx[0], f, g.ctx[1], g.frames[fs->bottom_field].frame, 1, 0), "h264r_field_copy");
                g.frames[fs->frame] = GpuFrameEntry{ f, fs->frame->poc, true, (uint8_t)H264R_FRAME };
            }
        }
    }
}

// first slice of a picture (init_picture has run: core/slice_data.cc:149-313)
void begin_picture(slice_t& slice)
{
    open_engine(*slice.active_sps, slice.header);
    const storable_picture* pic = slice.dec_picture;
    sweep_dead_frames(slice.p_Vid, nullptr);
    convert_references(slice.p_Vid);
    const int structure = slice.header.structure == TOP_FIELD ? H264R_TOP_FIELD : (slice.header.structure == BOTTOM_FIELD ? H264R_BOTTOM_FIELD : H264R_FRAME);
    const h264r_frame dst = new_frame(pic, slice.header.PicOrderCnt, structure);
    memset(&g.pp, 0, sizeof(g.pp));
    g.pp.structure = structure;
    // reference table of the picture: every picture the DPB holds (later slices may list other references than the first
    // one).  POC and long-term state are informational for the engine (implicit weights are precomputed in fill_slice
    // from the pictures the slice lists, which are alive).
    for (auto& kv : g.frames) {
        if (kv.first == pic || (kv.second.structure != H264R_FRAME) != (g.kind != 0)) continue;      // pictures of this picture's form
        const int i = g.pp.num_ref_frames++;
        g.pp.ref_frames[i] = kv.second.frame;
        g.pp.ref_poc[i] = kv.second.poc;
        g.pp.ref_long_term[i] = 0;
        g.pp.ref_structure[i] = kv.second.structure;
    }
    g.pp.num_slices = 1;                     // grows with every slice; final value set before submit
    g.pp.poc = slice.header.PicOrderCnt;
    g.pp.run_deblock = 1;
    g.pp.direct_8x8_inference_flag = slice.active_sps->direct_8x8_inference_flag;
    h264r_pic_params tmp = g.pp;
    tmp.num_slices = 64;                     // staging capacity check only; see end_picture
    check(h264r_picture_begin(g.ctx[g.kind], dst, &tmp, &g.bufs), "h264r_picture_begin");
    g.facade.init(g.bufs, g.width_mbs, g.kind ? g.height_mbs / 2 : g.height_mbs, slice.header.field_pic_flag != 0);      // field scans: transform.cc:339-386
    g.cur = pic;
    g.any_deblock = false;
}

// shr_t / pps_t fields of one slice + the tables the reference derives per slice header
void fill_slice(slice_t& slice, const QuantLists& q)
{
    const shr_t& shr = slice.header;
    const pps_t& pps = *slice.active_pps;
    const int nr = slice.current_slice_nr;
    if (nr < 0 || nr >= 64) error(500, "h264recon: more than 64 slices in a picture");
    if (shr.MbaffFrameFlag || slice.active_sps->separate_colour_plane_flag ||
        (shr.slice_type != P_slice && shr.slice_type != B_slice && shr.slice_type != I_slice))
        error(500, "h264recon: %s", h264r_strerror(H264R_ERR_UNSUPPORTED));
    h264r_slice& s = g.bufs.slices[nr];
    memset(&s, 0, sizeof(s));
    s.slice_type = shr.slice_type == P_slice ? H264R_P_SLICE : (shr.slice_type == B_slice ? H264R_B_SLICE : H264R_I_SLICE);
    s.disable_deblocking_filter_idc = (uint8_t)shr.disable_deblocking_filter_idc;
    s.filter_offset_a = (int8_t)shr.FilterOffsetA; s.filter_offset_b = (int8_t)shr.FilterOffsetB;
    s.luma_log2_weight_denom = shr.luma_log2_weight_denom; s.chroma_log2_weight_denom = shr.chroma_log2_weight_denom;
    s.weighted_pred_flag = pps.weighted_pred_flag; s.weighted_bipred_idc = (uint8_t)pps.weighted_bipred_idc;
    s.constrained_intra_pred_flag = pps.constrained_intra_pred_flag;
    s.direct_spatial_mv_pred_flag = shr.direct_spatial_mv_pred_flag;
    if (shr.disable_deblocking_filter_idc != 1) g.any_deblock = true;
    for (int list = 0; list < 2; ++list) {
        s.num_ref[list] = (uint8_t)slice.RefPicSize[list];
        for (int i = 0; i < H264R_MAX_REFS; ++i) {
            const storable_picture* r = (s.slice_type != H264R_I_SLICE && i < slice.RefPicSize[list]) ? slice.RefPicList[list][i] : nullptr;
            s.ref_pic_list[list][i] = r ? (int8_t)slot_of(frame_of(r)) : (int8_t)-1;
            for (int pl = 0; pl < 3; ++pl)
                if (i < (int)shr.pred_weight_l[list][pl].size()) {
                    s.wp_weight[list][pl][i] = shr.pred_weight_l[list][pl][i].weight;
                    s.wp_offset[list][pl][i] = shr.pred_weight_l[list][pl][i].offset;
                }
        }
    }
    if (s.slice_type == H264R_B_SLICE && pps.weighted_bipred_idc == 2)      // inter_prediction.cc:112-139
        for (int i = 0; i < slice.RefPicSize[0] && i < H264R_MAX_REFS; ++i)
            for (int j = 0; j < slice.RefPicSize[1] && j < H264R_MAX_REFS; ++j) {
                const storable_picture* r0 = slice.RefPicList[0][i];
                const storable_picture* r1 = slice.RefPicList[1][j];
                int w0 = 32, w1 = 32;
                if (r0 && r1) h264r_implicit_weights(shr.PicOrderCnt, r0->poc, r1->poc, r0->is_long_term, r1->is_long_term, &w0, &w1);
                s.implicit_w1[i][j] = (int16_t)w1;
            }
    const int* q4[6]; const int* q8[2];
    for (int i = 0; i < 6; ++i) q4[i] = q.q4[i];
    for (int i = 0; i < 2; ++i) q8[i] = q.q8[i];
    g.facade.assign_quant_params(nr, q4, q8);
    g.pp.num_slices = std::max(g.pp.num_slices, nr + 1);
}

h264r::FacadeMb to_facade(const mb_t& mb)
{
    h264r::FacadeMb f;
    memset(&f, 0, sizeof(f));
    f.mbAddrX = mb.mbAddrX;
    f.is_intra_block = mb.is_intra_block;
    f.slice_nr = mb.slice_nr;
    f.mb_type = mb.mb_type;
    f.transform_size_8x8_flag = mb.transform_size_8x8_flag;
    f.intra_chroma_pred_mode = mb.intra_chroma_pred_mode;
    for (int i = 0; i < 4; ++i) { f.SubMbType[i] = mb.SubMbType[i]; f.SubMbPredMode[i] = mb.SubMbPredMode[i]; f.Intra8x8PredMode[i] = mb.Intra8x8PredMode[i]; }
    for (int i = 0; i < 16; ++i) f.Intra4x4PredMode[i] = mb.Intra4x4PredMode[i];
    f.Intra16x16PredMode = mb.Intra16x16PredMode;
    f.CodedBlockPatternLuma = mb.CodedBlockPatternLuma; f.CodedBlockPatternChroma = mb.CodedBlockPatternChroma;
    f.QpY = mb.QpY; f.QpC[0] = mb.QpC[0]; f.QpC[1] = mb.QpC[1];
    for (int i = 0; i < 3; ++i) f.cbp_blks[i] = mb.cbp_blks[i];
    return f;
}

// one transmitted level: the facade appends it and returns the cbp_blks bits Transform::coeff_luma_ac would have set
template <typename F>
void forward_level(mb_t* mb, F call)
{
    h264r::FacadeMb f;
    memset(&f, 0, sizeof(f));
    f.mbAddrX = mb->mbAddrX;
    f.transform_size_8x8_flag = mb->transform_size_8x8_flag;
    call(&f);
    for (int i = 0; i < 3; ++i) mb->cbp_blks[i] |= f.cbp_blks[i];
    if (g.facade.overflowed()) error(500, "h264recon: level list overflow");
}

} // namespace

// The parser pokes Transform::cof directly for I_PCM (parser/interpret_mb.cc:444-470) and zeroes it per MB
// (core/slice_data.cc:496-503); only that member is ever touched, so the unchanged class declaration serves as scratch.
Decoder::Decoder() :
    intra_prediction { nullptr },
    inter_prediction { nullptr },
    transform        { new Transform },
    deblock          { nullptr }
{
    memset(transform->cof, 0, sizeof(transform->cof));
}

Decoder::~Decoder()
{
    g.quant.erase(this);
    delete this->transform;
}

void Decoder::init(slice_t& slice)
{
    if (slice.dec_picture != g.cur) begin_picture(slice);
    auto it = g.quant.find(this);
    if (it == g.quant.end()) error(500, "h264recon: Decoder::init before assign_quant_params");
    fill_slice(slice, it->second);
}

// Transform::init (transform.cc:173-262): Flat / Default (Tables 7-3, 7-4) / SPS / PPS lists with fall-back rules A and B
void Decoder::assign_quant_params(slice_t& slice)
{
    const sps_t& sps = *slice.active_sps;
    const pps_t& pps = *slice.active_pps;
    const int* qm[8];
    for (int i = 0; i < 8; ++i) qm[i] = Flat16;
    if (pps.pic_scaling_matrix_present_flag || sps.seq_scaling_matrix_present_flag) {
        if (sps.seq_scaling_matrix_present_flag)
            for (int i = 0; i < 8; ++i) {
                if (i < 6) {
                    if (!sps.seq_scaling_list_present_flag[i])
                        qm[i] = i == 0 ? h264r::kDefault4x4Intra : (i == 3 ? h264r::kDefault4x4Inter : qm[i - 1]);
                    else if (sps.UseDefaultScalingMatrix4x4Flag[i]) qm[i] = i < 3 ? h264r::kDefault4x4Intra : h264r::kDefault4x4Inter;
                    else qm[i] = sps.ScalingList4x4[i];
                } else {
                    if (!sps.seq_scaling_list_present_flag[i]) qm[i] = i == 6 ? h264r::kDefault8x8Intra : h264r::kDefault8x8Inter;
                    else if (sps.UseDefaultScalingMatrix8x8Flag[i - 6]) qm[i] = i == 6 ? h264r::kDefault8x8Intra : h264r::kDefault8x8Inter;
                    else qm[i] = sps.ScalingList8x8[i - 6];
                }
            }
        if (pps.pic_scaling_matrix_present_flag)
            for (int i = 0; i < 8; ++i) {
                if (i < 6) {
                    if (!pps.pic_scaling_list_present_flag[i]) {
                        if (i == 0) { if (!sps.seq_scaling_matrix_present_flag) qm[i] = h264r::kDefault4x4Intra; }
                        else if (i == 3) { if (!sps.seq_scaling_matrix_present_flag) qm[i] = h264r::kDefault4x4Inter; }
                        else qm[i] = qm[i - 1];
                    } else if (pps.UseDefaultScalingMatrix4x4Flag[i]) qm[i] = i < 3 ? h264r::kDefault4x4Intra : h264r::kDefault4x4Inter;
                    else qm[i] = pps.ScalingList4x4[i];
                } else {
                    if (!pps.pic_scaling_list_present_flag[i]) {
                        if (i == 6) { if (!sps.seq_scaling_matrix_present_flag) qm[i] = h264r::kDefault8x8Intra; }
                        else if (!sps.seq_scaling_matrix_present_flag) qm[i] = h264r::kDefault8x8Inter;
                    } else if (pps.UseDefaultScalingMatrix8x8Flag[i - 6]) qm[i] = i == 6 ? h264r::kDefault8x8Intra : h264r::kDefault8x8Inter;
                    else qm[i] = pps.ScalingList8x8[i - 6];
                }
            }
    }
    QuantLists& q = g.quant[this];
    for (int i = 0; i < 6; ++i) for (int k = 0; k < 16; ++k) q.q4[i][k] = qm[i][k];
    for (int i = 0; i < 2; ++i) for (int k = 0; k < 64; ++k) q.q8[i][k] = qm[6 + i][k];
}

void Decoder::coeff_luma_dc(mb_t* mb, ColorPlane pl, int x0, int y0, int runarr, int levarr)
{
    forward_level(mb, [&](h264r::FacadeMb* f) { g.facade.coeff_luma_dc(f, (h264r::ColorPlane)pl, x0, y0, runarr, levarr); });
}
void Decoder::coeff_luma_ac(mb_t* mb, ColorPlane pl, int x0, int y0, int runarr, int levarr)
{
    forward_level(mb, [&](h264r::FacadeMb* f) { g.facade.coeff_luma_ac(f, (h264r::ColorPlane)pl, x0, y0, runarr, levarr); });
}
void Decoder::coeff_chroma_dc(mb_t* mb, ColorPlane pl, int x0, int y0, int runarr, int levarr)
{
    forward_level(mb, [&](h264r::FacadeMb* f) { g.facade.coeff_chroma_dc(f, (h264r::ColorPlane)pl, x0, y0, runarr, levarr); });
}
void Decoder::coeff_chroma_ac(mb_t* mb, ColorPlane pl, int x0, int y0, int runarr, int levarr)
{
    forward_level(mb, [&](h264r::FacadeMb* f) { g.facade.coeff_chroma_ac(f, (h264r::ColorPlane)pl, x0, y0, runarr, levarr); });
}

// the DC Hadamards run inside residual_kernel
void Decoder::transform_luma_dc(mb_t*, ColorPlane) {}
void Decoder::transform_chroma_dc(mb_t*, ColorPlane) {}

// the MB is completely parsed (core/slice_data.cc:646): snapshot it
void Decoder::decode(mb_t& mb)
{
    slice_t& slice = *mb.p_Slice;
    if (mb.mb_field_decoding_flag || mb.TransformBypassModeFlag) error(500, "h264recon: %s", h264r_strerror(H264R_ERR_UNSUPPORTED));
    h264r::FacadeMb f = to_facade(mb);
    if (mb.mb_type == I_PCM) {
        for (int pl = 0; pl < 3; ++pl)
            for (int y = 0; y < (pl ? 8 : 16); ++y)
                for (int x = 0; x < (pl ? 8 : 16); ++x)
                    g.facade.pcm_sample(&f, (h264r::ColorPlane)pl, x, y, this->transform->cof[pl][y][x]);
        if (g.facade.overflowed()) error(500, "h264recon: level list overflow");
    }
    h264r::FacadeMotion m[16];
    const storable_picture* pic = slice.dec_picture;
    for (int b = 0; b < 16; ++b) {
        const pic_motion_params& p = pic->mv_info[mb.mb.y * 4 + (b >> 2)][mb.mb.x * 4 + (b & 3)];
        for (int l = 0; l < 2; ++l) {
            m[b].ref_pic[l] = (!mb.is_intra_block && p.ref_pic[l]) ? slot_of(frame_of(p.ref_pic[l])) : -1;
            m[b].mv[l][0] = p.mv[l].mv_x; m[b].mv[l][1] = p.mv[l].mv_y;
            m[b].ref_idx[l] = p.ref_idx[l];
        }
    }
    g.facade.decode(f, m);
}

// exit_picture (framebuf/picture.cc:253): the picture is complete on the host side -- reconstruct it
void Decoder::deblock_filter(slice_t& slice)
{
    storable_picture* pic = slice.dec_picture;
    if (pic != g.cur) error(500, "h264recon: deblock_filter for a picture that was not begun");
    // Deblock::deblock (deblock.cc:622-656): runs unless every slice has disable_deblocking_filter_idc == 1
    g.pp.run_deblock = (g.any_deblock && (0x03 & (1 << pic->used_for_reference))) ? 1 : 0;
    check(h264r_picture_update(g.ctx[g.kind], g.bufs.picture, &g.pp), "h264r_picture_update");
    check(h264r_picture_submit(g.ctx[g.kind], g.bufs.picture, g.facade.stream_words()), "h264r_picture_submit");
    check(h264r_flush(g.ctx[g.kind]), "h264r_flush");
    // Nothing is copied back here: the picture stays in HBM (motion compensation of later pictures reads it there) and
    // reaches the host when the DPB outputs it (output_gpu.cc).  The call returns while the kernels run, so the parsing
    // of the next picture overlaps this one's reconstruction.
    g.cur = nullptr;
}

// for output_gpu.cc
h264r_ctx* gpu_engine_of(const storable_picture* p)
{
    auto it = g.frames.find(p);
    if (it == g.frames.end()) error(500, "h264recon: picture was not reconstructed by the GPU path");
    return ctx_of(it->second.structure);
}
h264r_frame gpu_frame_of_picture(const storable_picture* p)
{
    auto it = g.frames.find(p);
    if (it == g.frames.end()) error(500, "h264recon: picture was not reconstructed by the GPU path");
    return it->second.frame;
}
bool gpu_has_picture(const storable_picture* p) { return g.frames.find(p) != g.frames.end(); }
// a picture that never entered the DPB is about to be deleted (direct_output): its engine frame goes back to the pool
void gpu_picture_freed(const storable_picture* p)
{
    auto it = g.frames.find(p);
    if (it == g.frames.end()) return;
    h264r_frame_release(ctx_of(it->second.structure), it->second.frame);
    g.frames.erase(it);
}

// test hook (tests/quant_select_test.cc): the weightScale lists assign_quant_params selected for `d`, [0..5] 4x4, [6..7] 8x8
const int* gpu_quant_list(const Decoder* d, int i)
{
    auto it = g.quant.find(d);
    if (it == g.quant.end()) return nullptr;
    return i < 6 ? it->second.q4[i] : it->second.q8[i - 6];
}

void Decoder::get_block_luma(storable_picture*, int, int, int, int, px_t[16][16], int, mb_t&)
{
    error(500, "h264recon: error concealment is not provided by the GPU path");
}

}
}
