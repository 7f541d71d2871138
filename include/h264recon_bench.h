/*
 * h264recon_bench.h -- measurement hooks of libh264recon.so.  NOT part of the drop-in boundary (include/h264recon.h):
 * nothing a decoder needs is declared here.  bench.py, the profiling scripts and a few tests use these to time the
 * kernels on HBM-resident inputs, to drive the public entry points from several feeder threads the way a set of
 * parser threads would, and to measure the box's own host<->device copy ceiling.
 */
#ifndef H264RECON_BENCH_H_
#define H264RECON_BENCH_H_

#include "h264recon.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Re-runs the last flush (the picture descriptions of the last flush must still be in HBM: nothing has been begun
 * since).  flags: H264R_REPLAY_H2D re-issues the host->device copies of every picture description from the pinned
 * staging; without it only the kernels run on the HBM-resident inputs.  H264R_REPLAY_TIME_KERNELS brackets every
 * kernel with CUDA events.  ms_out[0] = whole replay (CUDA events on the compute stream), ms_out[1 + k] = time inside
 * kernel kind k, launches_out[1 + k] = its launches; h264r_bench_kernel_name(k) names kind k (NULL past the last). */
#define H264R_REPLAY_H2D           1
#define H264R_REPLAY_TIME_KERNELS  2
#define H264R_REPLAY_ASYNC         4   /* enqueue only (like h264r_flush): no host synchronisation, no timings;
                                          the caller joins with h264r_wait.  Lets downloads overlap the kernels. */
#define H264R_BENCH_MAX_KERNELS    8
int  h264r_replay_last_flush(h264r_ctx* ctx, int iterations, int flags, float ms_out[1 + H264R_BENCH_MAX_KERNELS],
                             int launches_out[1 + H264R_BENCH_MAX_KERNELS]);
const char* h264r_bench_kernel_name(int kind);

/* One pre-parsed picture as a parser thread would have left it in the staging buffers: the two blocks that cross PCIe. */
typedef struct h264r_bench_picture {
    h264r_pic_params pp;             /* ref_frames[] already hold frame handles                          */
    h264r_frame      dst;
    int32_t          stream_id;      /* pictures of one stream are fed by one thread, in list order      */
    const void*      head;           /* [mbs | slices]: width_mbs*height_mbs h264r_mb, then pp.num_slices h264r_slice */
    const uint32_t*  stream;
    uint32_t         stream_words;
    int32_t          pitch_y;        /* of `out` (chroma pitch = pitch_y / 2)                            */
    uint8_t*         out;            /* pinned destination of the reconstructed frame (Y|Cb|Cr), or NULL */
} h264r_bench_picture;

/* End to end through the PUBLIC entry points: `threads` feeder threads (stand-ins for one parser thread per stream;
 * feeder t owns the streams with stream_id % threads == t) each do h264r_picture_begin -> memcpy of head and stream
 * into the staging -> h264r_picture_submit for their pictures, in list order.  The calling thread flushes whenever
 * `flush_every` pictures have been submitted (h264r_flush) and enqueues h264r_frame_download_async of every flushed
 * picture with an `out`.  The list is fed `steps` times back to back; one join at the end (h264r_wait(ctx, -1)).
 * Returns the wall time of the whole call in seconds, or a negative status;
 * host_fill_s = seconds the feeder threads spent inside begin + memcpy + submit (summed over threads),
 * host_flush_s = seconds the calling thread spent inside h264r_flush + the download calls. */
double h264r_bench_feed(h264r_ctx* ctx, const h264r_bench_picture* pics, int num_pics, int num_mbs, int threads,
                        int flush_every, int steps, double* host_fill_s, double* host_flush_s);

/* Box ceiling: plain pinned cudaMemcpyAsync of `h2d_bytes` host->device and `d2h_bytes` device->host in chunks of
 * `chunk` bytes on two streams of device `device`, concurrently, `iterations` times; no kernels, no engine.
 * gbs_out[0] = host->device GB/s, [1] = device->host GB/s, [2] = seconds per iteration with both directions running. */
int  h264r_bench_copy_ceiling(int device, size_t h2d_bytes, size_t d2h_bytes, size_t chunk, int iterations, double gbs_out[3]);

#ifdef __cplusplus
}
#endif
#endif /* H264RECON_BENCH_H_ */
