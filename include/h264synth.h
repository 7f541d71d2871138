/*
 * h264synth.h -- deterministic synthetic macroblock-data generator (SURVEY.md §8d "Synthetic inputs").
 *
 * Produces, picture by picture in decode order, exactly what the host-side entropy decoder would hand to
 * the reconstruction boundary: h264r_mb headers, h264r_mb_motion, h264r_slice tables and raw coefficient
 * levels, in the layout of include/h264recon.h.  The same bytes feed the CUDA engine, the CPU restatement
 * (oracle/port_recon.c) and the reference's own Decoder (oracle/ref_harness.cc).
 *
 * RNG = splitmix64, seed = 0x4832363400000000 + (config << 16) + (stream << 8); one state per stream.
 * Only legal data is produced (intra modes whose neighbours exist, CBP-consistent levels, int32-safe
 * dequantisation, weight denominators <= 6, |mv_y| < 2048): see SURVEY.md §8a quirks 3, 6, 10, 11.
 */
#ifndef H264SYNTH_H_
#define H264SYNTH_H_

#include "h264recon.h"

#ifdef __cplusplus
extern "C" {
#endif

/* BASELINE.json configs[0..4] */
enum {
    H264S_CFG_CIF_BASELINE = 1,      /* 22x18 MB, I P P P, 4x4 transform, 1 ref                        */
    H264S_CFG_720P_MAIN    = 2,      /* 80x45, I B B P, bi-pred + explicit/implicit WP, 2 slices/odd    */
    H264S_CFG_1080P_HIGH   = 3,      /* 120x68, 8x8 transform, I8x8, scaling lists, constrained intra   */
    H264S_CFG_4K_HIGH      = 4,      /* 240x135, I/P/B, low QP, deblock offsets +-6, all-intra frames   */
    H264S_CFG_MULTI_1080P  = 5,      /* 64 x config 3 with distinct seeds                              */
    H264S_CFG_1080I_FIELDS = 6       /* 120x34: the FIELD pictures of a 1080i stream (field_pic_flag = 1 on every picture,
                                        parities alternating in display order), otherwise the rules of config 3:
                                        references of both parities, field deblocking rules (not in BASELINE.json:
                                        SURVEY.md 8f-4, the PAFF part of the feature tail)                         */
};

typedef struct h264s_stream h264s_stream;

typedef struct h264s_pic_info {
    int32_t pic_index;               /* decode-order index of this picture within the stream           */
    int32_t pic_type;                /* H264R_{P,B,I}_SLICE of all its slices                          */
    int32_t used_for_reference;      /* nal_ref_idc != 0                                               */
    int32_t poc;
    int32_t num_refs;                /* == pic_params.num_ref_frames                                   */
    int32_t ref_pic_index[H264R_MAX_REFS]; /* decode-order index of pic_params.ref_frames[i]           */
    int32_t last_use_of_ref[H264R_MAX_REFS]; /* 1 if no later picture of the stream references it       */
    uint32_t num_levels;             /* entries written to the level list                                */
} h264s_pic_info;

/* width_mbs/height_mbs/num_frames <= 0 select the config's own values */
h264s_stream* h264s_open(int config, int stream_idx, int width_mbs, int height_mbs, int num_frames);
void          h264s_close(h264s_stream* s);
void          h264s_get_seq(const h264s_stream* s, h264r_seq_params* sp, int* num_frames);

/* Generates the next picture.  mbs/motion must hold width_mbs*height_mbs entries, slices at least 4, levels
 * `level_capacity` entries (384 per MB always suffices).  pp->ref_frames[] is left as -1: the caller maps
 * info->ref_pic_index[] to frame handles.  Returns 1 if a picture was produced, 0 at end of stream, -1 if the
 * level list overflowed. */
int h264s_next(h264s_stream* s, h264s_pic_info* info, h264r_pic_params* pp, h264r_mb* mbs,
               h264r_mb_motion* motion, h264r_slice* slices, h264r_level* levels, uint32_t level_capacity);

/* Algorithmic (compulsory) HBM bytes of one picture, SURVEY.md §8d: every input byte read once, every output byte
 * written once, ideal reference fetch without halo.
 *   out[0] whole path : per MB 32 (header) + 4 per transmitted level + 192 (motion, inter MBs)
 *                       + 96 per used prediction list and 8x8 quadrant (= 384 per list and MB) + 384 (output)
 *   out[1] inter kernel: the above restricted to inter MBs
 *   out[2] intra kernel: the above restricted to intra MBs
 *   out[3] deblock kernel (its own pass): per filtered MB 32 + 384 read + 384 written (+192 motion if inter)
 *   out[4] inter MBs, out[5] intra MBs, out[6] MBs with at least one level, out[7] filtered MBs */
void h264s_account(const h264r_mb* mbs, const h264r_slice* slices, int nmb, int run_deblock, uint64_t out[8]);

#ifdef __cplusplus
}
#endif
#endif
