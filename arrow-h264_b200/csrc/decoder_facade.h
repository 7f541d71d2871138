// Host-side mirror of the reference's reconstruction interface, `class vio::h264::Decoder`
// (reference src/codec/h264/decoder/decoder.h:301-338): same entry points, same argument meaning, same call order
// as the parser uses them (parser/interpret_residual.cc:159-170, 407-414, 427-431, 471-477; core/slice_data.cc:646;
// framebuf/picture.cc:253).  Instead of reconstructing on the CPU it records what the GPU needs into the flat
// per-picture buffers of include/h264recon.h -- levels and packed motion entries appended to the picture's stream in
// parse order, nothing left for submit to do.  Plain C++11, no CUDA: the buffers normally are the pinned staging
// handed out by h264r_picture_begin, but any memory works (tests use malloc).
#ifndef H264R_DECODER_FACADE_H_
#define H264R_DECODER_FACADE_H_

#include "h264recon.h"
#include "h264_tables.h"

namespace h264r {

enum ColorPlane { PLANE_Y = 0, PLANE_U = 1, PLANE_V = 2 };

// The mb_t members the reconstruction path reads (parser/macroblock.h:78-135), names unchanged.
struct FacadeMb {
    int      mbAddrX;
    bool     is_intra_block;
    short    slice_nr;
    uint8_t  mb_type;
    bool     transform_size_8x8_flag;
    uint8_t  intra_chroma_pred_mode;
    uint8_t  SubMbType[4];
    uint8_t  SubMbPredMode[4];
    uint8_t  Intra4x4PredMode[16];
    uint8_t  Intra8x8PredMode[4];
    uint8_t  Intra16x16PredMode;
    uint8_t  CodedBlockPatternLuma;
    uint8_t  CodedBlockPatternChroma;
    int8_t   QpY;
    int8_t   QpC[2];
    uint64_t cbp_blks[3];
};

// pic_motion_params (framebuf/picture.h:66-71) with the picture pointer replaced by its reference slot
struct FacadeMotion {
    int     ref_pic[2];       // index into h264r_pic_params::ref_frames, -1 == nullptr
    int16_t mv[2][2];
    int8_t  ref_idx[2];
};

class Decoder {
public:
    // Decoder::init (decoder.cc:52-57): bind to the buffers of the picture being parsed.  field_pic_flag (shr.field_pic_flag)
    // selects the field scans for the coefficient positions (Transform::inverse_scan_*, transform.cc:339-386)
    void init(const h264r_pic_buffers& bufs, int width_mbs, int height_mbs, bool field_pic_flag = false);

    // Decoder::assign_quant_params (decoder.cc:59-62 -> Transform::init/set_quant): weightScale lists of slice
    // `slice_nr` in raster order, [0..5] 4x4 (Intra Y,Cb,Cr, Inter Y,Cb,Cr), [0..1] 8x8 (Intra Y, Inter Y)
    void assign_quant_params(int slice_nr, const int* const qmatrix4x4[6], const int* const qmatrix8x8[2]);

    // Decoder::coeff_* (decoder.cc:81-96 -> transform.cc:425-456): one transmitted level.  x0, y0 in 4x4-block
    // units, runarr = scan index, levarr = level.  The level is appended RAW; dequantisation happens on the GPU.
    void coeff_luma_dc  (FacadeMb* mb, ColorPlane pl, int x0, int y0, int runarr, int levarr);
    void coeff_luma_ac  (FacadeMb* mb, ColorPlane pl, int x0, int y0, int runarr, int levarr);
    void coeff_chroma_dc(FacadeMb* mb, ColorPlane pl, int x0, int y0, int runarr, int levarr);
    void coeff_chroma_ac(FacadeMb* mb, ColorPlane pl, int x0, int y0, int runarr, int levarr);

    // Decoder::transform_luma_dc / transform_chroma_dc (decoder.cc:98-105): the DC Hadamards run on the GPU
    void transform_luma_dc  (FacadeMb*, ColorPlane) {}
    void transform_chroma_dc(FacadeMb*, ColorPlane) {}

    // replaces the parser's direct pokes of Transform::cof for I_PCM (parser/interpret_mb.cc:444-470)
    void pcm_sample(FacadeMb* mb, ColorPlane pl, int x, int y, int value);

    // Decoder::decode (decoder.cc:65-79): the MB is completely parsed; snapshot its header and, packed to their distinct
    // entries (h264r_pack_motion), the 16 pic_motion_params of dec_picture->mv_info (raster order) into the stream
    void decode(FacadeMb& mb, const FacadeMotion motion[16]);

    // Decoder::deblock_filter (decoder.cc:107-110) is the flush point: the caller submits the picture
    // (h264r_picture_submit(ctx, picture, stream_words()) + h264r_flush).
    uint32_t stream_words() const { return n_levels_; }
    bool     overflowed() const { return overflow_; }

private:
    void append(FacadeMb* mb, int pos, int level);
    h264r_pic_buffers bufs_;
    int width_mbs_ = 0, height_mbs_ = 0;
    uint32_t n_levels_ = 0;          // words of the stream in use (levels and motion entries)
    int cur_mb_ = -1;
    uint32_t cur_first_ = 0;
    bool overflow_ = false;
    bool field_ = false;
    ZigZag zz_;
};

} // namespace h264r
#endif
