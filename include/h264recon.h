/*
 * h264recon.h -- C ABI of the B200-native H.264 macroblock-reconstruction engine.
 *
 * This is the drop-in boundary for the reconstruction path of luuvish/arrow-h264: everything that sits
 * behind `class vio::h264::Decoder` (reference src/codec/h264/decoder/decoder.h:301-338).  The serial
 * entropy decoder / slice parser stays on the host and fills, per picture, the flat buffers declared
 * here (coefficient levels, macroblock headers, motion, slice tables); the GPU reconstructs the whole
 * picture (dequant + IDCT, motion compensation, intra prediction, deblocking) on `h264r_picture_submit`.
 *
 * Plain C, plain pointers and sizes, `int` status codes, no exceptions, no exit().  One context per GPU.
 * Threading: any number of pictures can be in the filling state at once -- the reference embeds one Decoder in every
 * slice_t (parser/slice.h:173), so fill state belongs to the picture, not to the process.  h264r_picture_begin /
 * _update / _fill / _submit may be called from any thread (one parser thread per stream); the buffers of a picture
 * are written by the thread that began it, without locks.  h264r_flush, h264r_wait, the frame pool and the download
 * calls belong to ONE thread per context (the thread that owns the GPU), like the reference's decoder loop
 * (core/slice_data.cc:636-661).
 *
 * Supported stream subset (anything else returns H264R_ERR_UNSUPPORTED, there is NO CPU fallback):
 * 8-bit 4:2:0 frame pictures, or the field pictures of a stream coded in fields throughout (h264r_pic_params::structure; a
 * context holds pictures of one size; a stream that switches between the two uses a context of each kind and
 * h264r_field_copy); no MBAFF, no transform bypass, no SP/SI slices, no FMO/ASO.
 *
 * Every struct below is also the HBM layout: the host fills pinned staging memory laid out exactly as
 * the device reads it, so submit is a handful of large cudaMemcpyAsync calls.
 */
#ifndef H264RECON_H_
#define H264RECON_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------ */
/* constants (reference numbering, parser/macroblock.h:36-76, parser/slice.h:25-31)                  */

enum {
    H264R_P_SLICE = 0, H264R_B_SLICE = 1, H264R_I_SLICE = 2
};
enum {                               /* mb_t::mb_type                                                */
    H264R_MB_SKIP_DIRECT = 0,        /* P_Skip / B_Skip / B_Direct_16x16                              */
    H264R_MB_16x16 = 1, H264R_MB_16x8 = 2, H264R_MB_8x16 = 3, H264R_MB_8x8 = 4,
    H264R_SUB_8x4 = 5, H264R_SUB_4x8 = 6, H264R_SUB_4x4 = 7,      /* SubMbType values 4..7, 0=direct */
    H264R_MB_I4x4 = 8, H264R_MB_I8x8 = 9, H264R_MB_I16x16 = 10, H264R_MB_IPCM = 12
};
enum { H264R_PRED_L0 = 0, H264R_PRED_L1 = 1, H264R_PRED_BI = 2 };

#define H264R_MB_FLAG_INTRA   0x01u  /* mb_t::is_intra_block                                          */
#define H264R_MB_FLAG_T8x8    0x02u  /* mb_t::transform_size_8x8_flag                                 */

#define H264R_MAX_REFS        32     /* entries of one RefPicList (frame decoding uses <= 16)         */
#define H264R_COEFFS_PER_MB   384    /* 256 Y + 64 Cb + 64 Cr                                         */

/* One transmitted coefficient level: what one Decoder::coeff_* call carries (decoder.h:312-315).
 * pos = raster position inside the MB's 384 samples (Y 16x16: y*16+x; Cb: 256 + y*8+x; Cr: 320 + y*8+x) after
 * the inverse scan; level = the raw (not yet dequantised) level.  I_PCM: 384 entries, level = the sample. */
typedef uint32_t h264r_level;
#define H264R_LEVEL(pos, level)   ((uint32_t)(uint16_t)(pos) | ((uint32_t)(uint16_t)(int16_t)(level) << 16))
#define H264R_LEVEL_POS(e)        ((int)((e) & 0xFFFFu))
#define H264R_LEVEL_VALUE(e)      ((int)(int16_t)((e) >> 16))

/* status codes */
enum {
    H264R_OK = 0,
    H264R_ERR_INVALID = -1,          /* bad argument / bad handle                                     */
    H264R_ERR_UNSUPPORTED = -2,      /* stream feature outside the supported subset                   */
    H264R_ERR_NOMEM = -3,            /* frame pool / staging exhausted or cudaMalloc failed           */
    H264R_ERR_CUDA = -4,             /* CUDA runtime error (h264r_last_cuda_error gives the text)     */
    H264R_ERR_STATE = -5,            /* call order violated (e.g. submit without begin)               */
    H264R_ERR_NODEVICE = -6          /* no CUDA device: the engine never falls back to the CPU        */
};

/* ------------------------------------------------------------------------------------------------ */
/* per-macroblock header: 32 bytes.  Snapshot of the mb_t fields the reconstruction path reads       */
/* (parser/macroblock.h:78-135; list extracted in SURVEY.md §8a row TY).                             */

typedef struct h264r_mb {
    uint8_t  mb_type;                /* mb_t::mb_type                                                 */
    uint8_t  flags;                  /* H264R_MB_FLAG_*                                               */
    uint16_t slice_idx;              /* index into the picture's slice table == mb_t::slice_nr        */
    uint8_t  cbp_luma;               /* CodedBlockPatternLuma   (bit b = 8x8 block b)                 */
    uint8_t  cbp_chroma;             /* CodedBlockPatternChroma (0, 1 = DC only, 2 = DC + AC)         */
    int8_t   qp_y;                   /* QpY      (== qp_scaled[0] at 8 bit)                           */
    int8_t   qp_c[2];                /* QpC[0,1] (== qp_scaled[1,2]), from update_qp, interpret_mb.cc:784 */
    uint8_t  intra16_mode;           /* Intra16x16PredMode                                            */
    uint8_t  chroma_mode;            /* intra_chroma_pred_mode                                        */
    uint8_t  reserved0;
    uint16_t cbp_blks;               /* cbp_blks[0] bits 0..15: 4x4 block (by*4+bx) has luma AC levels
                                        (transform.cc:433-436; 0xFFFF for I_PCM, interpret_mb.cc:421) */
    uint16_t coeff_count;            /* number of h264r_level entries of this MB (0 = no residual)     */
    uint32_t coeff_offset;           /* word index of the MB's first level in the picture's stream       */
    union {
        uint8_t intra_modes[8];      /* 16 nibbles, low nibble first: Intra4x4PredMode[luma4x4BlkIdx]
                                        (I_4x4) or Intra8x8PredMode[0..3] in nibbles 0..3 (I_8x8)     */
        struct {
            uint8_t sub_mb_type[4];      /* SubMbType[mbPartIdx]  (== mb_type for non-8x8 MBs)        */
            uint8_t sub_mb_pred_mode[4]; /* SubMbPredMode[mbPartIdx] after direct-mode resolution     */
        } inter;
    } u;
    uint32_t motion;                 /* inter MBs: (word index of the MB's first h264r_motion_entry in the
                                        picture's stream) << 4 | layout code, see h264r_pack_motion; 0 for
                                        intra MBs                                                       */
} h264r_mb;

/* Motion of one partition as the device reads it: 12 bytes = three words of the picture's stream.                */
typedef struct h264r_motion_entry {
    int16_t mv[2][2];                /* [list][x,y] quarter-pel                                        */
    int8_t  ref_idx[2];              /* pic_motion_params::ref_idx (may be 0 for an unused list, quirk 2) */
    int8_t  ref_pic[2];              /* identity of pic_motion_params::ref_pic: index into
                                        h264r_pic_params::ref_frames, -1 == nullptr.  For a list the
                                        block predicts from it is the picture motion compensation reads
                                        (== ref_pic_list[list][ref_idx], as the reference's parser sets
                                        it, parser/interpret_mv.cc) and the one the bS rule compares    */
} h264r_motion_entry;

/* The 16 pic_motion_params (framebuf/picture.h:66-71) of one MB, unpacked: 192 bytes, 4x4 blocks in raster order
 * (by*4+bx).  Host-side convenience form only (input of h264r_pack_motion / h264r_picture_fill): an MB sends just its
 * DISTINCT entries across PCIe.                                                                          */
typedef struct h264r_mb_motion {
    int16_t mv[2][16][2];            /* [list][blk][x,y]                                               */
    int8_t  ref_idx[2][16];
    int8_t  ref_pic[2][16];
} h264r_mb_motion;

/* Packs the motion of one MB: writes the distinct entries to out[] (at most 16) and returns the layout code:
 * 1 = one entry for the MB | 2 = rows 0-1 / rows 2-3 | 3 = columns 0-1 / columns 2-3 | 4 = the four 8x8 quadrants |
 * 5 = all sixteen 4x4 blocks.  Entries per code: h264r_motion_entries_of_code[].  The MB header then carries
 * motion = (word index of out[0] in the stream) << 4 | code.                                              */
int  h264r_pack_motion(const h264r_mb_motion* m, h264r_motion_entry out[16]);
extern const uint8_t h264r_motion_entries_of_code[6];
/* The inverse: the sixteen per-block entries of an MB from its packed form (`motion` = h264r_mb::motion). */
void h264r_unpack_motion(const uint32_t* stream, uint32_t motion, h264r_mb_motion* out);
/* The same for a whole picture held in unpacked form, into caller memory (what h264r_picture_fill does into the
 * staging): copies the levels to stream[0 .. num_levels) -- coeff_offset stays valid --, copies the headers to out_mbs
 * (may equal mbs), appends the packed motion of every inter MB and sets h264r_mb::motion.  Host only, no CUDA.
 * Returns the stream words used, or a negative status (H264R_ERR_NOMEM: stream_capacity too small). */
int64_t h264r_pack_picture(int num_mbs, const h264r_mb* mbs, const h264r_mb_motion* motion, const h264r_level* levels,
                           uint32_t num_levels, h264r_mb* out_mbs, uint32_t* stream, uint32_t stream_capacity);

/* per-slice table.  Fields of shr_t / pps_t read by the path plus the tables the reference builds per
 * slice header (Decoder::assign_quant_params -> Transform::init, transform.cc:173-302).              */
typedef struct h264r_slice {
    uint8_t  slice_type;                         /* H264R_{P,B,I}_SLICE                                */
    uint8_t  disable_deblocking_filter_idc;
    int8_t   filter_offset_a;                    /* shr.FilterOffsetA                                  */
    int8_t   filter_offset_b;
    uint8_t  luma_log2_weight_denom;
    uint8_t  chroma_log2_weight_denom;
    uint8_t  weighted_pred_flag;                 /* pps                                                */
    uint8_t  weighted_bipred_idc;                /* pps                                                */
    uint8_t  constrained_intra_pred_flag;        /* pps                                                */
    uint8_t  direct_spatial_mv_pred_flag;
    uint8_t  num_ref[2];                         /* slice_t::RefPicSize                                */
    int8_t   ref_pic_list[2][H264R_MAX_REFS];    /* RefPicList[list][ref_idx] -> index into ref_frames  */
    int8_t   wp_weight[2][3][H264R_MAX_REFS];    /* pred_weight_l[list][plane][ref_idx].weight (int8!)  */
    int8_t   wp_offset[2][3][H264R_MAX_REFS];
    int16_t  implicit_w1[H264R_MAX_REFS][H264R_MAX_REFS]; /* weighted_bipred_idc==2: weight1 for
                                                    (ref_idx0, ref_idx1); weight0 = 64 - weight1
                                                    (inter_prediction.cc:112-139, host precomputes)    */
    uint16_t level_scale_4x4[2][3][6][16];       /* [0 intra / 1 inter][plane][qp%6][j*4+i]             */
    uint16_t level_scale_8x8[2][6][64];          /* [0 intra / 1 inter][qp%6][j*8+i] (luma only, 4:2:0) */
    uint8_t  reserved[20];
} h264r_slice;

/* per-picture parameters */
enum { H264R_FRAME = 0, H264R_TOP_FIELD = 1, H264R_BOTTOM_FIELD = 2 };   /* h264r_pic_params::structure */
typedef int32_t h264r_frame;                     /* frame-pool id, >= 0                                */

typedef struct h264r_pic_params {
    int32_t     num_slices;
    int32_t     num_ref_frames;                  /* distinct reference pictures used by this picture   */
    h264r_frame ref_frames[H264R_MAX_REFS];      /* pool ids; motion.ref_pic / ref_pic_list index this  */
    int32_t     run_deblock;                     /* reference: some slice has idc != 1 and
                                                    used_for_reference in {0,1} (deblock.cc:631-643)   */
    int32_t     poc;                             /* informational (implicit weights are precomputed)   */
    int32_t     ref_poc[H264R_MAX_REFS];         /* informational: POC of ref_frames[i]                */
    uint8_t     ref_long_term[H264R_MAX_REFS];   /* informational                                      */
    int32_t     direct_8x8_inference_flag;       /* sps of THIS picture's stream (decoder.cc:239-242): streams
                                                    that share a context may differ                     */
    int32_t     structure;                       /* shr.structure: H264R_FRAME, or H264R_TOP_FIELD / H264R_BOTTOM_FIELD
                                                    for a field picture (field_pic_flag = 1, PAFF).  A field picture
                                                    is a picture of its own of half the frame height (what the
                                                    reference's storable_picture of a field is): the context is created
                                                    with PicHeightInMbs = FrameHeightInMbs / 2 and every pool frame holds
                                                    one field.  Differences on the path: field scans on the host side
                                                    (transform.cc:344-382), the chroma vector offset between fields of
                                                    different parity (inter_prediction.cc:352-354), mvlimit 2 and bS 3
                                                    on horizontal MB edges in the deblocking rule (deblock.cc:86, 106,
                                                    164, 188).  Frame and field pictures do not mix in one context:
                                                    see h264r_field_copy.  MBAFF is unsupported.            */
    uint8_t     ref_structure[H264R_MAX_REFS];   /* structure of ref_frames[i] (all H264R_FRAME for a frame picture,
                                                    fields for a field picture)                            */
} h264r_pic_params;

/* context parameters (picture size = sps_t; every stream that shares the context has this size) */
typedef struct h264r_seq_params {
    int32_t width_mbs;                           /* PicWidthInMbs                                      */
    int32_t height_mbs;                          /* FrameHeightInMbs                                   */
    int32_t direct_8x8_inference_flag;           /* not read by the engine (h264r_pic_params carries it per
                                                    picture); kept for single-stream callers           */
    int32_t max_frames;                          /* frame pool capacity                                */
    int32_t max_pictures_in_flight;              /* staging slots: pictures begun but not yet copied to HBM */
    int32_t max_slices_per_picture;
    int32_t max_levels_per_picture;              /* sizes the stream of a picture: capacity in words =
                                                    this + 48 * MBs (worst-case motion); 0 = worst case
                                                    (384 levels per MB)                                */
} h264r_seq_params;

/* Host staging of one picture, handed out by h264r_picture_begin (pinned memory owned by the context).  The layout IS
 * the HBM layout: submit + flush copy [mbs | slices used] and [stream used] as they are, nothing is repacked.
 *   stream: 32-bit words.  An MB's levels are coeff_count consecutive words at coeff_offset (any order inside the MB);
 *           an inter MB's motion entries are 3 words each at (motion >> 4).  The parser side appends both as it goes:
 *           levels while the residual is parsed (Decoder::coeff_*), motion entries when the MB is complete
 *           (Decoder::decode). */
typedef struct h264r_pic_buffers {
    h264r_mb*        mbs;                        /* [width_mbs*height_mbs], raster order               */
    h264r_slice*     slices;                     /* [max_slices_per_picture]                           */
    uint32_t*        stream;                     /* [stream_capacity] levels and motion entries        */
    uint32_t         stream_capacity;            /* words                                              */
    int32_t          picture;                    /* handle for _update / _fill / _submit               */
} h264r_pic_buffers;

typedef struct h264r_ctx h264r_ctx;

/* ------------------------------------------------------------------------------------------------ */
/* entry points                                                                                       */

/* replaces: slice_t owning a `Decoder decoder` (parser/slice.h:173) + storable_picture allocation     */
int  h264r_create(h264r_ctx** out, int device, const h264r_seq_params* sp);
void h264r_destroy(h264r_ctx* ctx);

/* replaces: new storable_picture(...) planes (framebuf/picture.cc:15-83); device Y/Cb/Cr, uint8        */
int  h264r_frame_alloc(h264r_ctx* ctx, h264r_frame* out);
int  h264r_frame_release(h264r_ctx* ctx, h264r_frame f);

/* replaces: init_picture + Decoder::init (core/slice_data.cc:149-313, 618).  Reserves a staging slot; blocks while
 * all slots wait for their host->device copy (never on kernels).  Thread-safe. */
int  h264r_picture_begin(h264r_ctx* ctx, h264r_frame dst, const h264r_pic_params* pp, h264r_pic_buffers* out);
/* Replaces the picture parameters given to h264r_picture_begin while the picture is still being filled: a parser
 * learns the number of slices, further reference pictures (slices of one picture may list different references)
 * and whether any slice enables the deblocking filter only as it goes along (core/slice_data.cc:618 runs per slice). */
int  h264r_picture_update(h264r_ctx* ctx, int32_t picture, const h264r_pic_params* pp);
/* Convenience for callers that hold a complete unpacked description (generators, tests): copies mbs / slices / levels
 * into the staging of `picture`, packs every inter MB's motion behind the levels and sets h264r_mb::motion.  Runs on
 * the caller's thread (O(MBs)); returns the stream words used (pass them to submit) or a negative status. */
int64_t h264r_picture_fill(h264r_ctx* ctx, int32_t picture, const h264r_mb* mbs, const h264r_mb_motion* motion,
                           const h264r_slice* slices, int num_slices, const h264r_level* levels, uint32_t num_levels);
/* replaces: Decoder::deblock_filter at exit_picture (framebuf/picture.cc:253): the picture is complete on the host
 * side; it is queued.  `stream_words` = words of the stream actually used.  O(slices): the per-MB content is checked
 * by the kernels that read it (out-of-range values are clamped and reported by h264r_wait as H264R_ERR_INVALID).
 * Thread-safe; pictures are launched in submission order. */
int  h264r_picture_submit(h264r_ctx* ctx, int32_t picture, uint32_t stream_words);
/* launches everything queued: pictures are grouped into dependency waves (a picture whose references
 * are produced by a queued picture goes to a later wave); every wave is one batched launch sequence.  */
int  h264r_flush(h264r_ctx* ctx);
/* blocks until frame `f` is reconstructed (the wave that writes it; later pictures keep running), or, with
 * f < 0, until everything queued has finished, downloads included.  Returns H264R_ERR_INVALID if a kernel met a
 * macroblock description outside its domain since the last wait (slice index, level range, QP, reference slot,
 * macroblock type ...: the value was clamped, the picture is wrong, nothing was accessed out of bounds). */
int  h264r_wait(h264r_ctx* ctx, h264r_frame f);

/* replaces: write_out_picture reading imgY/imgUV (framebuf/output.cc:109-227)                          */
int  h264r_frame_download(h264r_ctx* ctx, h264r_frame f, uint8_t* y, uint8_t* cb, uint8_t* cr,
                          int pitch_y, int pitch_c);
/* test/seed helper: set a frame's samples (e.g. an externally decoded reference picture)               */
int  h264r_frame_upload(h264r_ctx* ctx, h264r_frame f, const uint8_t* y, const uint8_t* cb, const uint8_t* cr,
                        int pitch_y, int pitch_c);

/* replaces: dpb_split_field / dpb_combine_field_yuv (framebuf/dpb.cc), for a stream that switches between frame and field
 * pictures (PAFF).  Such a stream uses TWO contexts on one device: frame_ctx holds its frame pictures, field_ctx -- created
 * with half the height -- its field pictures.  When the reference lists of a picture name a picture that exists in the other
 * form only, the caller converts it: the lines of field `parity` (0 top, 1 bottom) of `frame` are copied to `field`
 * (to_field != 0: one half of dpb_split_field) or from it (to_field == 0: one half of dpb_combine_field_yuv), device to
 * device, asynchronously, ordered behind the picture that produces the source and ahead of every later use of both.
 * H264R_ERR_STATE if a picture that was submitted but not flushed yet writes the source or names the destination. */
int  h264r_field_copy(h264r_ctx* frame_ctx, h264r_frame frame, h264r_ctx* field_ctx, h264r_frame field, int parity, int to_field);

/* Asynchronous variant: the copy is ordered after the wave that produces `f` and runs on the engine's D2H
 * stream, overlapping later waves; destination should be pinned (h264r_host_alloc).  h264r_wait(ctx, -1) joins. */
int  h264r_frame_download_async(h264r_ctx* ctx, h264r_frame f, uint8_t* y, uint8_t* cb, uint8_t* cr,
                                int pitch_y, int pitch_c);
/* The display rectangle of a frame (frame_cropping of the SPS; offsets in luma samples, even for 4:2:0), ordered
 * after the wave that produces `f` only -- pictures queued behind it keep running.  Blocks until the copy is done.
 * replaces: img2buf with crop_left/right/top/bottom in write_out_picture (framebuf/output.cc:30-104, 147-187)   */
int  h264r_frame_download_cropped(h264r_ctx* ctx, h264r_frame f, int crop_left, int crop_right, int crop_top,
                                  int crop_bottom, uint8_t* y, uint8_t* cb, uint8_t* cr, int pitch_y, int pitch_c);
/* pinned host memory for download destinations */
void* h264r_host_alloc(size_t bytes);
void  h264r_host_free(void* p);

/* host helper restating inter_prediction.cc:112-139 (implicit bi-prediction weights)                   */
void h264r_implicit_weights(int cur_poc, int poc0, int poc1, int long_term0, int long_term1, int* w0, int* w1);
/* host helper restating transform.cc:265-302 (set_quant) for 8-bit 4:2:0: qmatrix[0..5] 4x4 lists
 * (Intra Y,Cb,Cr, Inter Y,Cb,Cr) and qmatrix[6..7] 8x8 lists (Intra Y, Inter Y), raster order          */
void h264r_build_level_scale(h264r_slice* s, const int* const qmatrix4x4[6], const int* const qmatrix8x8[2]);

/* counters / diagnostics */
typedef struct h264r_stats {
    uint64_t kernel_launches;        /* number of kernels launched by this context so far              */
    uint64_t pictures;               /* pictures reconstructed                                         */
    uint64_t macroblocks;
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t waves;                  /* batched launch sequences                                       */
} h264r_stats;
int         h264r_get_stats(h264r_ctx* ctx, h264r_stats* out);
const char* h264r_strerror(int code);
const char* h264r_last_cuda_error(h264r_ctx* ctx);
int         h264r_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* H264RECON_H_ */
