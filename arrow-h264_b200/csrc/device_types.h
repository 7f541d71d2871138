// Device-side view of one queued picture and the launch interface between engine.cu (host runtime) and
// kernels.cu (sm_100a kernels).
#ifndef H264R_DEVICE_TYPES_H_
#define H264R_DEVICE_TYPES_H_

#include "h264recon.h"
#include <cuda.h>                    // CUtensorMap (type only; the encoder entry point is fetched through the runtime)
#include <cuda_runtime.h>

namespace h264r {

// Frame in HBM: one allocation, planes at fixed offsets, unpadded (kernels clamp coordinates, which is
// bit-identical to the reference's padded planes + block pre-clamp, SURVEY.md §8a derived facts).
struct FrameGeom {
    int width_mbs, height_mbs;
    int pitch_y, pitch_c;            // bytes; tight: 16 * width_mbs and 8 * width_mbs
    size_t off_cb, off_cr, bytes;    // plane offsets inside the allocation
};

// TMA descriptors of one frame (the -DH264R_INTER_TMA=1 build of recon_inter2_kernel fetches interior reference windows
// with cp.async.bulk.tensor): luma = 2-D uint8 {16 W, 16 H}, box 32 x 13; chroma = 3-D uint8 {8 W, 8 H, 2 planes}, box
// 32 x 5 x 2.  The boxes are 32 bytes wide for 13 / 5 useful ones because a box must START on a 16-byte boundary (measured:
// any other x origin raises "illegal instruction", scripts/probes/tma_probe3.cu), while a window starts at any sample.
// Out-of-bounds bytes are zero-filled, so only windows inside the picture take this path.
struct __align__(64) FrameMaps {
    CUtensorMap luma, chroma;
    int ok, pad[15];                 // 0: no descriptors (odd picture width: chroma pitch not a multiple of 16 bytes)
};

// Deblock descriptor of one MB, output of the parallel pre-pass (deblock_prep_kernel), input of the deblock wavefront: 64 bytes.
//   bs[dir * 2 + (edge >> 1)], nibble (edge & 1) * 4 + group = boundary strength 0..4 of the 4-sample group of that edge
//   par[plane][type], type 0 = left MB edge, 1 = internal edges, 2 = top MB edge: the filter thresholds of
//   filter_edge (deblock.cc:469-474 + tables :294-324) packed as alpha | beta << 8 | tc0[bS=1] << 13 | tc0[2] << 18 |
//   tc0[3] << 23, so that tc0(bS) = (par >> (8 + 5 * bS)) & 31.
struct DeblockDesc {
    uint32_t bs[4];
    uint32_t par[12];                // [Y, Cb, Cr][type 0..2] contiguous (9 words), 3 words unused
};

struct DevPicture {
    const h264r_mb*        mbs;
    const h264r_slice*     slices;
    const uint32_t*        stream;                    // levels and packed motion entries (h264recon.h h264r_pic_buffers)
    int16_t*               resid;                     // [nmb][384] residual plane, device only (residual_kernel)
    uint8_t*               dst;                       // frame base
    const uint8_t*         ref[H264R_MAX_REFS];       // frame bases of pic_params.ref_frames[]; unused slots: a dummy frame
    const FrameMaps*       ref_maps[H264R_MAX_REFS];  // their TMA descriptors
    DeblockDesc*           desc;                      // [nmb], device only
    uint64_t*              mbox;                      // [nmb][24], device only: row-to-row mailboxes { 4 samples, epoch }
    uint32_t*              mb_done;                   // [nmb], device only: epoch stamp of the launch that reconstructed the intra MB
    uint32_t*              intra_list;                // [nmb], device only: raster-ordered addresses of the intra MBs (intra_list_kernel)
    uint32_t*              intra_count;               // device only: entries of intra_list
    uint32_t               stream_words;              // words of `stream` in use (bounds of coeff_offset / motion)
    int                    num_slices;
    int                    num_refs;
    int                    run_deblock;
    int                    all_intra;                 // every slice is an I slice: row wavefront; otherwise inter + sparse intra kernels
    int                    direct8x8;                 // direct_8x8_inference_flag of the picture's stream
    // field pictures (h264r_pic_params::structure != H264R_FRAME); all 0 for frame pictures
    int                    field;                     // 1: mvlimit 2 and bS 3 on horizontal MB edges (deblock.cc:86, 106, 164, 188)
    int                    chroma_dy;                 // -2 (top field) / +2 (bottom field): added to the vertical chroma vector ...
    uint32_t               ref_opposite;              // ... for the reference slots of the other parity (inter_prediction.cc:352-354)
};

struct WaveLaunch {
    const DevPicture* pics;          // device array
    int   num_pics;
    int*  tickets;                   // device: work-ticket counters ([0] intra rows, [1] deblock, [2] sparse intra), zeroed per wave (64 ints)
    uint32_t* wave_max;              // device word of this wave: largest intra_count of its pictures (intra_list_kernel)
    uint32_t* err;                   // host-mapped word: kernels OR in a bit when they meet a description outside its domain
    FrameGeom geom;
    int   any_inter, any_deblock;
    int   any_intra_rows;            // some picture of the wave is all-intra: row wavefront kernel
    int   any_field;                 // some picture of the wave is a field picture: recon_inter2_kernel<true>
    uint32_t epoch;                  // stamp of this launch sequence (mailboxes, DevPicture::mb_done)
};

// Kernel launchers of one wave (kernels.cu).  Returns the number of kernels launched (0 when the wave has no work of that kind).
// KERNEL_RESID, KERNEL_DBPREP and KERNEL_LIST need nothing but the picture description (side stream, a wave ahead).
enum { KERNEL_RESID = 0, KERNEL_INTER = 1, KERNEL_INTRA = 2, KERNEL_DBPREP = 3, KERNEL_DEBLOCK = 4, KERNEL_LIST = 5, KERNEL_KINDS = 6 };
int launch_wave_kernel(const WaveLaunch& w, int which, cudaStream_t stream);
const char* wave_kernel_name(int which);

} // namespace h264r
#endif
