// Host runtime behind the C ABI of include/h264recon.h: frame pool, pinned staging slots, dependency waves,
// batched launches.  C++ host code + CUDA runtime only (no PyTorch, no NCCL: streams/GOPs are independent).
#include "device_types.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <utility>
#include <algorithm>

using namespace h264r;

static_assert(sizeof(h264r_mb) == 32, "h264r_mb must be 32 bytes");
static_assert(sizeof(h264r_mb_motion) == 192, "h264r_mb_motion must be 192 bytes");
static_assert(sizeof(DeblockDesc) == 64, "DeblockDesc must be 64 bytes");
static_assert(sizeof(h264r_slice) % 16 == 0, "h264r_slice must keep 16-byte alignment in arrays");

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Frame {
    uint8_t* dev = nullptr;
    bool used = false;
    int write_wave = -1, read_wave = -1;      // bookkeeping inside one flush
    cudaEvent_t ready = nullptr;              // ev_done of the wave that last wrote this frame (owned by the pool)
    cudaEvent_t read_done = nullptr;          // recorded on the D2H stream after the last asynchronous download
    bool pending_read = false;                // a download was enqueued since the frame was last (re)written
    int group = -1;                           // stream group of the picture that last wrote the frame (affinity of its successors)
    int ready_group = -1;                     // group whose record `ready` belongs to
    uint64_t ready_flush = 0;                 // flush in which `ready` was assigned (events are recycled per flush)
    std::vector<std::pair<cudaEvent_t, int> > readers;   // (ev_done, group) of the records of this flush that read the frame
};

enum SlotState { SLOT_FREE = 0, SLOT_FILLING, SLOT_QUEUED, SLOT_INFLIGHT };

struct Slot {
    uint8_t* host = nullptr;                  // pinned
    uint8_t* dev = nullptr;
    DeblockDesc* dev_desc = nullptr;          // device only: output of the deblock pre-pass
    int16_t* dev_resid = nullptr;             // device only: residual plane [nmb][384]
    uint8_t* host_motion = nullptr;           // pinned, host only: the full per-MB motion array the parser side fills
    uint32_t motion_entries = 0;              // packed 12-byte motion entries behind the level list
    uint32_t intra_count = 0;                 // intra-MB address list behind the packed motion (mixed pictures only)
    uint32_t* dev_mb_done = nullptr;          // device only: per-MB epoch stamps of the sparse intra kernel
    uint64_t* dev_mbox = nullptr;             // device only: deblock mailboxes [nmb][24]
    SlotState state = SLOT_FREE;
    h264r_pic_params pp;
    h264r_frame dst = -1;
    uint32_t used_levels = 0;
    int has_intra = 0, has_inter = 0;
    int wave = 0;
    int group = 0;
};

struct WaveCopy { int slot; size_t bytes; };

struct WaveRecord {
    WaveLaunch launch;
    std::vector<WaveCopy> copies;             // H2D copies of the picture descriptions of this wave
    std::vector<int> dst_frames;              // frames written by this wave
    int group = 0;                            // stream group that runs the record
    std::vector<cudaEvent_t> deps;            // ev_done of records of OTHER groups this one must follow (cross-group references)
    cudaEvent_t ev_h2d = nullptr;             // recorded on the H2D stream after the wave's copies
    cudaEvent_t ev_done = nullptr;            // recorded on the compute stream after the wave's kernels
    cudaEvent_t ev_side = nullptr;            // recorded on the side stream after the wave's neighbour-independent kernels
    cudaEvent_t ev_inter = nullptr;           // recorded on the compute stream after the wave's inter kernel
};

} // namespace

enum { kMaxGroups = 4 };

struct h264r_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;            // compute
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaStream_t s_side = nullptr;            // kernels that depend on nothing but the picture description (motion
                                              // expansion, residual, deblock descriptors): they run ahead of, and
                                              // underneath, the latency-bound wavefront kernels of earlier waves
    cudaEvent_t ev_fork = nullptr;
    // Independent streams (closed GOPs) are spread over `num_groups` groups, each with its own compute + side stream,
    // ticket counters: the latency-bound wavefront kernels of one group run underneath the
    // throughput-bound kernels of the others.  g_main[0] == stream, g_side[0] == s_side.
    int num_groups = 1;
    int group_policy = 0;                     // 0: by stream (reference affinity, round robin); 1: by role (see h264r_create)
    int rr_group = 0;                         // round-robin cursor for pictures without a queued predecessor
    cudaStream_t g_main[kMaxGroups] = { nullptr, nullptr, nullptr, nullptr };
    cudaStream_t g_side[kMaxGroups] = { nullptr, nullptr, nullptr, nullptr };
    cudaEvent_t g_tail[kMaxGroups] = { nullptr, nullptr, nullptr, nullptr };
    size_t sync_ints_per_group = 0;
    uint64_t flush_serial = 0;
    bool cross_group = false;                 // the last flush has references across groups
    bool side_gate = false;                   // H264R_SIDE_GATE=1: side kernels of wave k+1 start when the inter kernel of wave k has
                                              // finished, i.e. underneath its wavefront kernels (measured: no gain, 34.7 vs 34.6 ms/step)
    cudaEvent_t gate[kMaxGroups] = { nullptr, nullptr, nullptr, nullptr };   // ev_inter of the group's latest record
    std::vector<cudaEvent_t> event_pool;      // reused across flushes
    size_t events_used = 0;
    std::vector<cudaEvent_t> timer_events;    // H264R_REPLAY_TIME_KERNELS
    h264r_seq_params seq;
    FrameGeom geom;
    int nmb = 0;
    size_t off_mbs = 0, off_slices = 0, off_levels = 0, slot_bytes = 0;   // levels are followed by the packed motion
    uint32_t level_capacity = 0;
    std::vector<Frame> frames;
    std::vector<Slot> slots;
    std::vector<int> queue;                   // slot indexes in submission order
    std::vector<uint32_t> scratch_list;       // intra-MB addresses of the picture being submitted
    uint32_t epoch = 0;                       // launch-sequence counter (DevPicture::mb_done stamps)
    int filling = -1;
    DevPicture* h_pics = nullptr;             // pinned, [2][max_pictures_in_flight]: alternating halves per flush
    DevPicture* d_pics = nullptr;
    cudaEvent_t table_ev[2] = { nullptr, nullptr };   // end of the flush that last used each half
    int table_idx = 0;
    int* d_sync = nullptr;                    // per group: ticket counters (64 ints)
    std::vector<WaveRecord> last_waves;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    h264r_stats stats;
    char cuda_err[256];
};

namespace {

int cuda_fail(h264r_ctx* c, cudaError_t e, const char* what)
{
    snprintf(c->cuda_err, sizeof(c->cuda_err), "%s: %s", what, cudaGetErrorString(e));
    return H264R_ERR_CUDA;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call); } while (0)

// Packed motion.  Entry = mv[0], mv[1] (int16 x, y each), ref_idx[0], ref_idx[1], ref_pic[0], ref_pic[1]: 12 bytes.
// code 0 none | 1 one entry for the MB | 2 rows 0-1 / rows 2-3 | 3 columns 0-1 / columns 2-3 | 4 quadrants | 5 all 16.
const uint32_t kPackedEntries[6] = { 0, 1, 2, 2, 4, 16 };
inline void motion_entry(const h264r_mb_motion& m, int blk, uint8_t* e)
{
    memcpy(e, m.mv[0][blk], 4); memcpy(e + 4, m.mv[1][blk], 4);
    e[8] = (uint8_t)m.ref_idx[0][blk]; e[9] = (uint8_t)m.ref_idx[1][blk];
    e[10] = (uint8_t)m.ref_pic[0][blk]; e[11] = (uint8_t)m.ref_pic[1][blk];
}
int pack_motion(const h264r_mb_motion& m, uint8_t* out)
{
    uint8_t e[16][12];
    for (int b = 0; b < 16; ++b) motion_entry(m, b, e[b]);
    auto same = [&](int a, int b) { return memcmp(e[a], e[b], 12) == 0; };
    bool quad = true;
    for (int q = 0; q < 4 && quad; ++q) {
        const int b0 = (q >> 1) * 8 + (q & 1) * 2;
        quad = same(b0, b0 + 1) && same(b0, b0 + 4) && same(b0, b0 + 5);
    }
    if (!quad) { memcpy(out, e, sizeof(e)); return 5; }
    const bool top = same(0, 2), bottom = same(8, 10), left = same(0, 8), right = same(2, 10);
    if (top && bottom && left) { memcpy(out, e[0], 12); return 1; }
    if (top && bottom) { memcpy(out, e[0], 12); memcpy(out + 12, e[8], 12); return 2; }
    if (left && right) { memcpy(out, e[0], 12); memcpy(out + 12, e[2], 12); return 3; }
    memcpy(out, e[0], 12); memcpy(out + 12, e[2], 12); memcpy(out + 24, e[8], 12); memcpy(out + 36, e[10], 12);
    return 4;
}

bool frame_ok(const h264r_ctx* c, h264r_frame f) { return f >= 0 && f < (int)c->frames.size() && c->frames[f].used; }

// device frame -> host planes on `stream`: one copy when the destination is one tight contiguous block
int copy_frame_d2h(h264r_ctx* ctx, cudaStream_t stream, const uint8_t* d, uint8_t* y, uint8_t* cb, uint8_t* cr, int pitch_y, int pitch_c)
{
    const FrameGeom& g = ctx->geom;
    const int w = g.width_mbs * 16, h = g.height_mbs * 16;
    const size_t ny = (size_t)w * h, nc = ny / 4;
    if (pitch_y == w && pitch_c == w / 2) {
        if (cb == y + ny && cr == cb + nc) { CU(cudaMemcpyAsync(y, d, ny + 2 * nc, cudaMemcpyDeviceToHost, stream)); return H264R_OK; }
        CU(cudaMemcpyAsync(y, d, ny, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(cb, d + g.off_cb, nc, cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(cr, d + g.off_cr, nc, cudaMemcpyDeviceToHost, stream));
        return H264R_OK;
    }
    CU(cudaMemcpy2DAsync(y, pitch_y, d, g.pitch_y, w, h, cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpy2DAsync(cb, pitch_c, d + g.off_cb, g.pitch_c, w / 2, h / 2, cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpy2DAsync(cr, pitch_c, d + g.off_cr, g.pitch_c, w / 2, h / 2, cudaMemcpyDeviceToHost, stream));
    return H264R_OK;
}

cudaEvent_t take_event(h264r_ctx* c)
{
    if (c->events_used == c->event_pool.size()) {
        cudaEvent_t ev = nullptr;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        c->event_pool.push_back(ev);
    }
    return c->event_pool[c->events_used++];
}

// Runs the recorded waves of the last flush on three streams:
//   H2D stream   : the picture descriptions of a wave (they wait for the previous run of the same wave, whose staging
//                  in HBM they overwrite);
//   side stream  : the kernels that need nothing but the description -- motion expansion, residual, deblock
//                  descriptors.  They run ahead, underneath the latency-bound wavefront kernels of earlier waves;
//   compute      : inter, intra, deblock of the wave, after the side kernels of the wave and the waves before it.
// With time_kernels everything runs on the compute stream, one kernel at a time, bracketed by events.
int run_waves(h264r_ctx* ctx, bool h2d, bool time_kernels, float* ms_kernel, int* launches, bool fork_join)
{
    size_t timer_used = 0;
    struct Pending { int kind; size_t ev; };
    std::vector<Pending> pend;
    const int G = ctx->num_groups;
    if (!time_kernels && (fork_join || ctx->cross_group)) {
        // every stream of every group starts behind everything enqueued so far on every compute stream (picture table
        // upload on `stream`; with cross-group references also the previous run of the other groups)
        for (int gi = 1; gi < G; ++gi) { CU(cudaEventRecord(ctx->g_tail[gi], ctx->g_main[gi])); CU(cudaStreamWaitEvent(ctx->stream, ctx->g_tail[gi], 0)); }
        CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
        for (int gi = 0; gi < G; ++gi) {
            if (gi > 0) CU(cudaStreamWaitEvent(ctx->g_main[gi], ctx->ev_fork, 0));
            CU(cudaStreamWaitEvent(ctx->g_side[gi], ctx->ev_fork, 0));
        }
    }
    auto launch = [&](WaveRecord& rec, int kind, cudaStream_t st) -> int {
        if (time_kernels) {
            while (ctx->timer_events.size() < timer_used + 2) {
                cudaEvent_t ev = nullptr;
                if (cudaEventCreate(&ev) != cudaSuccess) return H264R_ERR_CUDA;
                ctx->timer_events.push_back(ev);
            }
            if (cudaEventRecord(ctx->timer_events[timer_used], st) != cudaSuccess) return H264R_ERR_CUDA;
        }
        const int launched = launch_wave_kernel(rec.launch, kind, st);
        if (launched) {
            ctx->stats.kernel_launches += (uint64_t)launched;
            if (launches) launches[kind + 1] += launched;
            if (time_kernels) {
                if (cudaEventRecord(ctx->timer_events[timer_used + 1], st) != cudaSuccess) return H264R_ERR_CUDA;
                pend.push_back({ kind, timer_used });
                timer_used += 2;
            }
        }
        return H264R_OK;
    };
    for (WaveRecord& rec : ctx->last_waves) {
        cudaStream_t main = time_kernels ? ctx->stream : ctx->g_main[rec.group];
        cudaStream_t side = time_kernels ? ctx->stream : ctx->g_side[rec.group];
        if (h2d) {
            CU(cudaStreamWaitEvent(ctx->s_h2d, rec.ev_done, 0));          // no-op before the first record
            for (const WaveCopy& c : rec.copies) {
                Slot& s = ctx->slots[c.slot];
                CU(cudaMemcpyAsync(s.dev, s.host, c.bytes, cudaMemcpyHostToDevice, ctx->s_h2d));
                ctx->stats.h2d_bytes += c.bytes;
            }
            CU(cudaEventRecord(rec.ev_h2d, ctx->s_h2d));
            CU(cudaStreamWaitEvent(side, rec.ev_h2d, 0));
        }
        rec.launch.epoch = ++ctx->epoch;
        // side kernels: their outputs (expanded motion, residual plane, deblock descriptors) are per picture slot; the
        // previous run of this record must have consumed them
        if (!time_kernels) CU(cudaStreamWaitEvent(side, rec.ev_done, 0));
        // Throughput-bound side kernels are scheduled underneath the latency-bound wavefront kernels (intra, deblock) of
        // the wave before, not against its inter kernel: they start when that inter kernel has finished.
        if (!time_kernels && ctx->side_gate && ctx->gate[rec.group]) CU(cudaStreamWaitEvent(side, ctx->gate[rec.group], 0));
        { const int rc = launch(rec, KERNEL_RESID, side); if (rc != H264R_OK) return rc; }
        { const int rc = launch(rec, KERNEL_DBPREP, side); if (rc != H264R_OK) return rc; }
        if (!time_kernels) {
            CU(cudaEventRecord(rec.ev_side, side));
            CU(cudaStreamWaitEvent(main, rec.ev_side, 0));
            for (cudaEvent_t dep : rec.deps) CU(cudaStreamWaitEvent(main, dep, 0));
        }
        // write-after-read: a frame still being downloaded (asynchronously, on the D2H stream) is not overwritten
        for (int f : rec.dst_frames) {
            Frame& fr = ctx->frames[f];
            if (fr.pending_read) { CU(cudaStreamWaitEvent(main, fr.read_done, 0)); fr.pending_read = false; }
        }
        CU(cudaMemsetAsync(rec.launch.tickets, 0, sizeof(int) * 64, main));      // the wave's ticket counters
        { const int rc = launch(rec, KERNEL_INTER, main); if (rc != H264R_OK) return rc; }
        if (!time_kernels && ctx->side_gate) { CU(cudaEventRecord(rec.ev_inter, main)); ctx->gate[rec.group] = rec.ev_inter; }
        { const int rc = launch(rec, KERNEL_INTRA, main); if (rc != H264R_OK) return rc; }
        { const int rc = launch(rec, KERNEL_DEBLOCK, main); if (rc != H264R_OK) return rc; }
        CU(cudaGetLastError());
        CU(cudaEventRecord(rec.ev_done, main));
        ctx->stats.waves += 1;
        ctx->stats.pictures += (uint64_t)rec.launch.num_pics;
        ctx->stats.macroblocks += (uint64_t)rec.launch.num_pics * ctx->nmb;
    }
    if (!time_kernels && fork_join)                        // `stream` ends behind every group
        for (int gi = 1; gi < G; ++gi) { CU(cudaEventRecord(ctx->g_tail[gi], ctx->g_main[gi])); CU(cudaStreamWaitEvent(ctx->stream, ctx->g_tail[gi], 0)); }
    if (time_kernels && ms_kernel) {
        CU(cudaStreamSynchronize(ctx->stream));
        for (const Pending& p : pend) {
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, ctx->timer_events[p.ev], ctx->timer_events[p.ev + 1]));
            ms_kernel[p.kind + 1] += ms;
        }
    }
    return H264R_OK;
}

} // namespace

extern "C" {

int h264r_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* h264r_strerror(int code)
{
    switch (code) {
    case H264R_OK: return "ok";
    case H264R_ERR_INVALID: return "invalid argument";
    case H264R_ERR_UNSUPPORTED: return "unsupported stream feature (8-bit 4:2:0 frame pictures only)";
    case H264R_ERR_NOMEM: return "out of frames, staging slots or device memory";
    case H264R_ERR_CUDA: return "CUDA runtime error";
    case H264R_ERR_STATE: return "call order violated";
    case H264R_ERR_NODEVICE: return "no CUDA device (this engine has no CPU fallback)";
    default: return "unknown error";
    }
}

const char* h264r_last_cuda_error(h264r_ctx* ctx) { return ctx ? ctx->cuda_err : ""; }

int h264r_create(h264r_ctx** out, int device, const h264r_seq_params* sp)
{
    if (!out || !sp || sp->width_mbs <= 0 || sp->height_mbs <= 0 || sp->max_frames <= 0 ||
        sp->max_pictures_in_flight <= 0 || sp->max_slices_per_picture <= 0) return H264R_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return H264R_ERR_NODEVICE;
    if (device < 0 || device >= ndev) return H264R_ERR_INVALID;
    h264r_ctx* ctx = new h264r_ctx();
    ctx->device = device; ctx->seq = *sp; ctx->cuda_err[0] = 0;
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
    if (e != cudaSuccess) { delete ctx; return H264R_ERR_CUDA; }

    FrameGeom& g = ctx->geom;
    g.width_mbs = sp->width_mbs; g.height_mbs = sp->height_mbs;
    g.pitch_y = sp->width_mbs * 16;            // tight: a whole frame is one contiguous 1.5*W*H block
    g.pitch_c = sp->width_mbs * 8;
    g.off_cb = (size_t)g.pitch_y * sp->height_mbs * 16;
    g.off_cr = g.off_cb + (size_t)g.pitch_c * sp->height_mbs * 8;
    g.bytes  = align_up(g.off_cr + (size_t)g.pitch_c * sp->height_mbs * 8, 256);
    ctx->nmb = sp->width_mbs * sp->height_mbs;

    ctx->off_mbs = 0;
    ctx->off_slices = align_up(ctx->off_mbs + sizeof(h264r_mb) * ctx->nmb, 256);
    ctx->off_levels = align_up(ctx->off_slices + sizeof(h264r_slice) * sp->max_slices_per_picture, 256);
    ctx->level_capacity = sp->max_levels_per_picture > 0 ? (uint32_t)sp->max_levels_per_picture
                                                          : (uint32_t)H264R_COEFFS_PER_MB * (uint32_t)ctx->nmb;
    // worst case of the packed motion: 16 entries of 12 bytes per MB
    // and of the intra-MB address list: 4 bytes per MB
    ctx->slot_bytes = align_up(ctx->off_levels + sizeof(h264r_level) * (size_t)ctx->level_capacity + (sizeof(h264r_mb_motion) + 4) * ctx->nmb, 256);

    ctx->frames.resize(sp->max_frames);
    ctx->slots.resize(sp->max_pictures_in_flight);
    // one pinned and one device arena for all staging slots
    uint8_t* h_arena = nullptr; uint8_t* d_arena = nullptr; DeblockDesc* d_desc = nullptr; int16_t* d_resid = nullptr;
    uint8_t* h_motion = nullptr; uint32_t* d_done = nullptr; uint64_t* d_mbox = nullptr;
    const size_t mbox_words = (size_t)24 * ctx->nmb;
    const size_t motion_bytes = sizeof(h264r_mb_motion) * (size_t)ctx->nmb;
    const size_t arena = ctx->slot_bytes * sp->max_pictures_in_flight;
    e = cudaHostAlloc((void**)&h_arena, arena, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_arena, arena);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_desc, sizeof(DeblockDesc) * (size_t)ctx->nmb * sp->max_pictures_in_flight);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_resid, sizeof(int16_t) * H264R_COEFFS_PER_MB * (size_t)ctx->nmb * sp->max_pictures_in_flight);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&h_motion, motion_bytes * sp->max_pictures_in_flight, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_done, sizeof(uint32_t) * (size_t)ctx->nmb * sp->max_pictures_in_flight);
    if (e == cudaSuccess) e = cudaMemset(d_done, 0, sizeof(uint32_t) * (size_t)ctx->nmb * sp->max_pictures_in_flight);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_mbox, sizeof(uint64_t) * mbox_words * sp->max_pictures_in_flight);
    if (e == cudaSuccess) e = cudaMemset(d_mbox, 0, sizeof(uint64_t) * mbox_words * sp->max_pictures_in_flight);   // epoch 0 is never used
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->h_pics, sizeof(DevPicture) * 2 * sp->max_pictures_in_flight, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_pics, sizeof(DevPicture) * 2 * sp->max_pictures_in_flight);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->table_ev[i], cudaEventDisableTiming);
    {   // H264R_STREAM_GROUPS = number of stream groups (1..4, default 1), each with its own compute + side stream,
        // ticket counters; H264R_GROUP_POLICY says how pictures are dealt to them:
        //   stream (default): whole streams (GOP chains), round robin;
        //   role: group 0 = the reference CHAIN (pictures that later pictures of the same flush predict from: I/P) on
        //         high-priority streams, groups 1.. = LEAVES (pictures nobody in the flush references: B).
        // Measured on B200, 64 x 1080p x 16 pictures (profiles/r1_groups_experiment.txt): 1 group 37.2 ms/step,
        // 2 groups by role 39.3, 3 by role 40.6, 2 by stream 39.8.  The wavefront kernels are bound by their critical
        // path; co-scheduled throughput kernels lengthen every step of that path (there is no intra-SM priority), so
        // concurrency between groups loses more than it fills.  Groups only pay when a flush holds few pictures per wave.
        const char* env = getenv("H264R_STREAM_GROUPS");
        const char* pol = getenv("H264R_GROUP_POLICY");
        ctx->num_groups = env ? atoi(env) : 1;
        if (ctx->num_groups < 1) ctx->num_groups = 1;
        if (ctx->num_groups > kMaxGroups) ctx->num_groups = kMaxGroups;
        ctx->group_policy = (pol && strcmp(pol, "role") == 0 && ctx->num_groups > 1) ? 1 : 0;
    }
    // Priorities: the compute streams (inter, intra, deblock: the dependency chain of the pictures) above the side
    // streams (motion expansion, residual, deblock descriptors: work that only has to be ready a wave ahead), so that
    // side CTAs fill the SMs the wavefront kernels leave idle instead of competing with the inter kernel.
    // Measured (profiles/r1_groups_experiment.txt): 35.9 ms/step against 34.6 with equal priorities -- the compute
    // stream then waits for late side kernels at every wave start -- so equal is the default; H264R_SIDE_PRIORITY=low
    // selects the lower side priority.  With the role policy the chain group sits above the leaf groups.
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);          // numerically lower = higher priority
    const char* sp_env = getenv("H264R_SIDE_PRIORITY");
    { const char* ge = getenv("H264R_SIDE_GATE"); ctx->side_gate = ge && atoi(ge) != 0; }
    const bool side_low = sp_env && strcmp(sp_env, "low") == 0;
    if (e == cudaSuccess) {
        // the streams created above have the default priority: replace them
        cudaStreamDestroy(ctx->stream); cudaStreamDestroy(ctx->s_side);
        e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->s_side, cudaStreamNonBlocking, side_low ? prio_lo : prio_hi);
    }
    ctx->g_main[0] = ctx->stream; ctx->g_side[0] = ctx->s_side;
    for (int gi = 0; gi < ctx->num_groups && e == cudaSuccess; ++gi) {
        if (gi > 0) {
            const int main_prio = ctx->group_policy == 1 ? (prio_hi < prio_lo ? prio_hi + 1 : prio_lo) : prio_hi;
            e = cudaStreamCreateWithPriority(&ctx->g_main[gi], cudaStreamNonBlocking, main_prio);
            if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->g_side[gi], cudaStreamNonBlocking, side_low ? prio_lo : main_prio);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->g_tail[gi], cudaEventDisableTiming);
    }
    ctx->sync_ints_per_group = 64;                     // ticket counters; the rows of the wavefront kernels talk through mailboxes
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_sync, sizeof(int) * ctx->sync_ints_per_group * ctx->num_groups);
    if (e != cudaSuccess) {
        snprintf(ctx->cuda_err, sizeof(ctx->cuda_err), "allocation: %s", cudaGetErrorString(e));
        if (h_arena) cudaFreeHost(h_arena);
        if (d_arena) cudaFree(d_arena);
        if (d_desc) cudaFree(d_desc);
        if (d_resid) cudaFree(d_resid);
        if (h_motion) cudaFreeHost(h_motion);
        if (d_done) cudaFree(d_done);
        if (d_mbox) cudaFree(d_mbox);
        if (ctx->h_pics) cudaFreeHost(ctx->h_pics);
        if (ctx->d_pics) cudaFree(ctx->d_pics);
        if (ctx->d_sync) cudaFree(ctx->d_sync);
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return H264R_ERR_NOMEM;
    }
    for (int i = 0; i < sp->max_pictures_in_flight; ++i) {
        ctx->slots[i].host = h_arena + ctx->slot_bytes * i;
        ctx->slots[i].dev = d_arena + ctx->slot_bytes * i;
        ctx->slots[i].dev_desc = d_desc + (size_t)ctx->nmb * i;
        ctx->slots[i].dev_resid = d_resid + (size_t)H264R_COEFFS_PER_MB * ctx->nmb * i;
        ctx->slots[i].host_motion = h_motion + motion_bytes * i;
        ctx->slots[i].dev_mb_done = d_done + (size_t)ctx->nmb * i;
        ctx->slots[i].dev_mbox = d_mbox + mbox_words * i;
    }
    *out = ctx;
    return H264R_OK;
}

void h264r_destroy(h264r_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int gi = 0; gi < kMaxGroups; ++gi) {
        if (gi > 0 && ctx->g_main[gi]) cudaStreamDestroy(ctx->g_main[gi]);
        if (gi > 0 && ctx->g_side[gi]) cudaStreamDestroy(ctx->g_side[gi]);
        if (ctx->g_tail[gi]) cudaEventDestroy(ctx->g_tail[gi]);
    }
    for (Frame& f : ctx->frames) { if (f.dev) cudaFree(f.dev); if (f.read_done) cudaEventDestroy(f.read_done); }
    if (!ctx->slots.empty()) { cudaFreeHost(ctx->slots[0].host); cudaFree(ctx->slots[0].dev); cudaFree(ctx->slots[0].dev_desc); cudaFree(ctx->slots[0].dev_resid);
                                cudaFreeHost(ctx->slots[0].host_motion); cudaFree(ctx->slots[0].dev_mb_done); cudaFree(ctx->slots[0].dev_mbox); }
    cudaFreeHost(ctx->h_pics); cudaFree(ctx->d_pics); cudaFree(ctx->d_sync);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
    for (int i = 0; i < 2; ++i) if (ctx->table_ev[i]) cudaEventDestroy(ctx->table_ev[i]);
    for (cudaEvent_t ev : ctx->event_pool) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->timer_events) cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->stream); cudaStreamDestroy(ctx->s_h2d); cudaStreamDestroy(ctx->s_d2h); cudaStreamDestroy(ctx->s_side);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    delete ctx;
}

int h264r_frame_alloc(h264r_ctx* ctx, h264r_frame* out)
{
    if (!ctx || !out) return H264R_ERR_INVALID;
    cudaSetDevice(ctx->device);
    for (size_t i = 0; i < ctx->frames.size(); ++i)
        if (!ctx->frames[i].used) {
            if (!ctx->frames[i].dev) CU(cudaMalloc((void**)&ctx->frames[i].dev, ctx->geom.bytes));
            ctx->frames[i].used = true;
            *out = (h264r_frame)i;
            return H264R_OK;
        }
    return H264R_ERR_NOMEM;
}

int h264r_frame_release(h264r_ctx* ctx, h264r_frame f)
{
    if (!ctx || !frame_ok(ctx, f)) return H264R_ERR_INVALID;
    ctx->frames[f].used = false;               // memory is kept for reuse; queued work that references it stays valid
    return H264R_OK;
}

int h264r_picture_begin(h264r_ctx* ctx, h264r_frame dst, const h264r_pic_params* pp, h264r_pic_buffers* out)
{
    if (!ctx || !pp || !out || !frame_ok(ctx, dst)) return H264R_ERR_INVALID;
    if (ctx->filling >= 0) return H264R_ERR_STATE;
    if (pp->num_slices <= 0 || pp->num_slices > ctx->seq.max_slices_per_picture) return H264R_ERR_INVALID;
    if (pp->num_ref_frames < 0 || pp->num_ref_frames > H264R_MAX_REFS) return H264R_ERR_INVALID;
    for (int i = 0; i < pp->num_ref_frames; ++i)
        if (pp->ref_frames[i] < 0 || pp->ref_frames[i] >= (int)ctx->frames.size() || !ctx->frames[pp->ref_frames[i]].dev)
            return H264R_ERR_INVALID;
    int s = -1;
    for (int attempt = 0; attempt < 2 && s < 0; ++attempt) {
        for (size_t i = 0; i < ctx->slots.size(); ++i) if (ctx->slots[i].state == SLOT_FREE) { s = (int)i; break; }
        if (s < 0 && attempt == 0) {
            // staging of flushed pictures is reusable once their H2D copies have drained
            cudaSetDevice(ctx->device);
            CU(cudaStreamSynchronize(ctx->stream));
            for (Slot& t : ctx->slots) if (t.state == SLOT_INFLIGHT) t.state = SLOT_FREE;
        }
    }
    if (s < 0) return H264R_ERR_NOMEM;
    Slot& sl = ctx->slots[s];
    sl.state = SLOT_FILLING; sl.pp = *pp; sl.dst = dst; sl.used_levels = 0;
    ctx->filling = s;
    out->mbs = reinterpret_cast<h264r_mb*>(sl.host + ctx->off_mbs);
    out->motion = reinterpret_cast<h264r_mb_motion*>(sl.host_motion);
    out->slices = reinterpret_cast<h264r_slice*>(sl.host + ctx->off_slices);
    out->levels = reinterpret_cast<h264r_level*>(sl.host + ctx->off_levels);
    out->level_capacity = ctx->level_capacity;
    return H264R_OK;
}

int h264r_picture_update(h264r_ctx* ctx, const h264r_pic_params* pp)
{
    if (!ctx || !pp) return H264R_ERR_INVALID;
    if (ctx->filling < 0) return H264R_ERR_STATE;
    if (pp->num_slices <= 0 || pp->num_slices > ctx->seq.max_slices_per_picture) return H264R_ERR_INVALID;
    if (pp->num_ref_frames < 0 || pp->num_ref_frames > H264R_MAX_REFS) return H264R_ERR_INVALID;
    for (int i = 0; i < pp->num_ref_frames; ++i)
        if (pp->ref_frames[i] < 0 || pp->ref_frames[i] >= (int)ctx->frames.size() || !ctx->frames[pp->ref_frames[i]].dev)
            return H264R_ERR_INVALID;
    ctx->slots[ctx->filling].pp = *pp;
    return H264R_OK;
}

int h264r_picture_submit(h264r_ctx* ctx, uint32_t num_levels)
{
    if (!ctx) return H264R_ERR_INVALID;
    if (ctx->filling < 0) return H264R_ERR_STATE;
    if (num_levels > ctx->level_capacity) return H264R_ERR_INVALID;
    Slot& sl = ctx->slots[ctx->filling];
    sl.used_levels = num_levels;
    // validate what would otherwise become an out-of-bounds access on the device
    h264r_mb* mbs = reinterpret_cast<h264r_mb*>(sl.host + ctx->off_mbs);
    // The per-MB motion (192 bytes, the 16 pic_motion_params the parser side filled) crosses PCIe in packed form: only
    // the distinct entries of an MB (1 when all 16 blocks agree, 2 for halves, 4 for quadrants, else 16), 12 bytes each,
    // right behind the level list.  reserved2 of the header = first entry << 4 | code; the kernels read the packed
    // entries directly (kernels.cu packed_entry).  Intra MBs send nothing (their motion is never read).
    const h264r_mb_motion* motion = reinterpret_cast<const h264r_mb_motion*>(sl.host_motion);
    uint8_t* const packed = sl.host + ctx->off_levels + sizeof(h264r_level) * (size_t)num_levels;
    uint32_t entries = 0;
    int has_intra = 0, has_inter = 0, bad = 0, unsupported = 0;
    ctx->scratch_list.clear();
    for (int i = 0; i < ctx->nmb; ++i) {
        h264r_mb& m = mbs[i];
        if (m.flags & H264R_MB_FLAG_INTRA) { has_intra = 1; m.reserved2 = 0; ctx->scratch_list.push_back((uint32_t)i); }
        else {
            has_inter = 1;
            const int code = pack_motion(motion[i], packed + (size_t)12 * entries);
            m.reserved2 = entries << 4 | (uint32_t)code;
            entries += kPackedEntries[code];
        }
        if (m.slice_idx >= sl.pp.num_slices) bad = 1;
        if (m.coeff_count && ((uint64_t)m.coeff_offset + m.coeff_count > num_levels)) bad = 1;   // positions are checked on the device
        if (m.mb_type > H264R_MB_IPCM || m.mb_type == 11) unsupported = 1;        // SI and friends
        if (!(m.flags & H264R_MB_FLAG_INTRA) && m.mb_type > H264R_MB_8x8) bad = 1;
        if (m.qp_y < 0 || m.qp_y > 51 || m.qp_c[0] < 0 || m.qp_c[0] > 51 || m.qp_c[1] < 0 || m.qp_c[1] > 51) bad = 1;
    }
    const h264r_slice* slices = reinterpret_cast<const h264r_slice*>(sl.host + ctx->off_slices);
    for (int k = 0; k < sl.pp.num_slices; ++k) {
        if (slices[k].slice_type > H264R_I_SLICE) unsupported = 1;               // SP / SI
        for (int list = 0; list < 2; ++list)
            for (int i = 0; i < H264R_MAX_REFS; ++i)
                if (slices[k].ref_pic_list[list][i] >= sl.pp.num_ref_frames) bad = 1;
    }
    if (bad || unsupported) {
        sl.state = SLOT_FREE; ctx->filling = -1;
        return unsupported ? H264R_ERR_UNSUPPORTED : H264R_ERR_INVALID;
    }
    sl.has_intra = has_intra; sl.has_inter = has_inter; sl.motion_entries = entries;
    // Mixed pictures reconstruct their (few) intra MBs one warp each, ordered by per-MB dependencies: the raster-ordered
    // address list follows the packed motion.  All-intra pictures run the row wavefront and need no list.
    sl.intra_count = 0;
    if (has_intra && has_inter) {
        sl.intra_count = (uint32_t)ctx->scratch_list.size();
        memcpy(packed + (size_t)12 * entries, ctx->scratch_list.data(), sizeof(uint32_t) * ctx->scratch_list.size());
    }
    sl.state = SLOT_QUEUED;
    ctx->queue.push_back(ctx->filling);
    ctx->filling = -1;
    return H264R_OK;
}

int h264r_flush(h264r_ctx* ctx)
{
    if (!ctx) return H264R_ERR_INVALID;
    if (ctx->filling >= 0) return H264R_ERR_STATE;
    if (ctx->queue.empty()) return H264R_OK;
    cudaSetDevice(ctx->device);

    // ---- dependency waves: RAW on reference frames, WAR/WAW on the destination frame ----
    for (Frame& f : ctx->frames) { f.write_wave = -1; f.read_wave = -1; }
    int num_waves = 0;
    for (int qi : ctx->queue) {
        Slot& s = ctx->slots[qi];
        int w = 0;
        for (int i = 0; i < s.pp.num_ref_frames; ++i) w = std::max(w, ctx->frames[s.pp.ref_frames[i]].write_wave + 1);
        Frame& d = ctx->frames[s.dst];
        w = std::max(w, std::max(d.write_wave, d.read_wave) + 1);
        if (d.write_wave < 0 && d.read_wave < 0) w = std::max(w, 0);
        s.wave = w;
        d.write_wave = w;
        for (int i = 0; i < s.pp.num_ref_frames; ++i) {
            Frame& r = ctx->frames[s.pp.ref_frames[i]];
            r.read_wave = std::max(r.read_wave, w);
        }
        num_waves = std::max(num_waves, w + 1);
    }
    // ---- stream groups: a picture joins the group of the picture that produced its first known reference ----
    const int G = ctx->num_groups;
    ctx->flush_serial += 1;
    for (Frame& f : ctx->frames) f.readers.clear();
    if (ctx->group_policy == 1) {
        // by role: a picture some later picture of this flush predicts from belongs to the chain (group 0), the others
        // are leaves, dealt round robin to groups 1..G-1
        std::vector<char> referenced(ctx->frames.size(), 0);
        int rr = 0;
        for (size_t k = ctx->queue.size(); k-- > 0; ) {
            Slot& s = ctx->slots[ctx->queue[k]];
            s.group = referenced[s.dst] ? 0 : 1 + (rr++ % (G - 1));
            referenced[s.dst] = 0;                                      // an earlier picture in the same frame is another picture
            for (int i = 0; i < s.pp.num_ref_frames; ++i) referenced[s.pp.ref_frames[i]] = 1;
            ctx->frames[s.dst].group = s.group;
        }
    } else
    for (int qi : ctx->queue) {
        Slot& s = ctx->slots[qi];
        int grp = -1;
        for (int i = 0; i < s.pp.num_ref_frames && grp < 0; ++i) grp = ctx->frames[s.pp.ref_frames[i]].group;
        if (grp < 0 || grp >= G) { grp = ctx->rr_group; ctx->rr_group = (ctx->rr_group + 1) % G; }
        s.group = grp;
        ctx->frames[s.dst].group = grp;
    }
    std::vector<int> order(ctx->queue);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const Slot& x = ctx->slots[a]; const Slot& y = ctx->slots[b];
        return x.wave != y.wave ? x.wave < y.wave : x.group < y.group;
    });

    // ---- records = runs of equal (wave, group) in that order ----
    std::vector<int> rec_begin;
    for (size_t k = 0; k < order.size(); ++k) {
        const Slot& c = ctx->slots[order[k]];
        if (k == 0 || c.wave != ctx->slots[order[k - 1]].wave || c.group != ctx->slots[order[k - 1]].group) rec_begin.push_back((int)k);
    }
    rec_begin.push_back((int)order.size());

    // ---- device picture table (one upload for all records) ----
    ctx->last_waves.clear();
    // the picture table alternates between two halves so that this flush never overwrites what the previous
    // (possibly still running) flush reads
    ctx->table_idx ^= 1;
    const size_t table_off = (size_t)ctx->table_idx * ctx->seq.max_pictures_in_flight;
    CU(cudaEventSynchronize(ctx->table_ev[ctx->table_idx]));
    DevPicture* const h_table = ctx->h_pics + table_off;
    DevPicture* const d_table = ctx->d_pics + table_off;
    for (size_t r = 0; r + 1 < rec_begin.size(); ++r)
        for (int k = rec_begin[r]; k < rec_begin[r + 1]; ++k) {
            Slot& s = ctx->slots[order[k]];
            DevPicture& p = h_table[k];
            memset(&p, 0, sizeof(p));
            p.mbs = reinterpret_cast<const h264r_mb*>(s.dev + ctx->off_mbs);
            p.packed_motion = s.dev + ctx->off_levels + sizeof(h264r_level) * (size_t)s.used_levels;
            p.intra_list = reinterpret_cast<const uint32_t*>(p.packed_motion + (size_t)12 * s.motion_entries);
            p.intra_count = (int)s.intra_count;
            p.mb_done = s.dev_mb_done;
            p.mbox = s.dev_mbox;
            p.slices = reinterpret_cast<const h264r_slice*>(s.dev + ctx->off_slices);
            p.levels = reinterpret_cast<const h264r_level*>(s.dev + ctx->off_levels);
            p.resid = s.dev_resid;
            p.dst = ctx->frames[s.dst].dev;
            p.desc = s.dev_desc;
            for (int i = 0; i < H264R_MAX_REFS; ++i)
                p.ref[i] = i < s.pp.num_ref_frames ? ctx->frames[s.pp.ref_frames[i]].dev : ctx->frames[s.dst].dev;
            p.run_deblock = s.pp.run_deblock; p.has_intra = s.has_intra; p.has_inter = s.has_inter;
        }
    CU(cudaMemcpyAsync(d_table, h_table, sizeof(DevPicture) * order.size(), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += sizeof(DevPicture) * order.size();

    // ---- per record: what to copy, what to launch, which records of other groups to follow ----
    ctx->events_used = 0;
    for (int gi = 0; gi < kMaxGroups; ++gi) ctx->gate[gi] = nullptr;   // the pool's events get new roles
    ctx->cross_group = false;
    for (size_t r = 0; r + 1 < rec_begin.size(); ++r) {
        const int b = rec_begin[r], e = rec_begin[r + 1];
        WaveRecord rec;
        rec.group = ctx->slots[order[b]].group;
        WaveLaunch& L = rec.launch;
        L.pics = d_table + b; L.num_pics = e - b; L.tickets = ctx->d_sync + ctx->sync_ints_per_group * rec.group; L.geom = ctx->geom;
        L.direct8x8 = ctx->seq.direct_8x8_inference_flag;
        L.any_inter = L.any_intra = L.any_deblock = L.any_intra_rows = L.max_intra_sparse = 0; L.epoch = 0;
        rec.ev_h2d = take_event(ctx); rec.ev_done = take_event(ctx); rec.ev_side = take_event(ctx); rec.ev_inter = take_event(ctx);
        if (!rec.ev_h2d || !rec.ev_done || !rec.ev_side || !rec.ev_inter) return H264R_ERR_CUDA;
        auto follow = [&](cudaEvent_t ev, int grp) {
            if (!ev || grp == rec.group) return;
            if (std::find(rec.deps.begin(), rec.deps.end(), ev) == rec.deps.end()) rec.deps.push_back(ev);
            ctx->cross_group = true;
        };
        for (int k = b; k < e; ++k) {
            Slot& s = ctx->slots[order[k]];
            rec.copies.push_back({ order[k], ctx->off_levels + sizeof(h264r_level) * (size_t)s.used_levels + (size_t)12 * s.motion_entries + sizeof(uint32_t) * s.intra_count });
            L.any_inter |= s.has_inter; L.any_intra |= s.has_intra; L.any_deblock |= s.pp.run_deblock;
            L.any_intra_rows |= (s.has_intra && !s.has_inter);
            L.max_intra_sparse = std::max(L.max_intra_sparse, (int)s.intra_count);
            // read-after-write on the references, write-after-read / write-after-write on the destination, as far as
            // the other side belongs to this flush and to another group (same group: stream order; earlier flushes:
            // the fork at the start of the run)
            for (int i = 0; i < s.pp.num_ref_frames; ++i) {
                Frame& f = ctx->frames[s.pp.ref_frames[i]];
                if (f.ready_flush == ctx->flush_serial) follow(f.ready, f.ready_group);
            }
            Frame& d = ctx->frames[s.dst];
            if (d.ready_flush == ctx->flush_serial) follow(d.ready, d.ready_group);
            for (const std::pair<cudaEvent_t, int>& rd : d.readers) follow(rd.first, rd.second);
        }
        for (int k = b; k < e; ++k) {
            Slot& s = ctx->slots[order[k]];
            for (int i = 0; i < s.pp.num_ref_frames; ++i) ctx->frames[s.pp.ref_frames[i]].readers.push_back(std::make_pair(rec.ev_done, rec.group));
        }
        for (int k = b; k < e; ++k) {
            Frame& d = ctx->frames[ctx->slots[order[k]].dst];
            d.ready = rec.ev_done; d.ready_group = rec.group; d.ready_flush = ctx->flush_serial; d.readers.clear();
            rec.dst_frames.push_back(ctx->slots[order[k]].dst);
        }
        ctx->last_waves.push_back(rec);
    }
    for (int qi : ctx->queue) ctx->slots[qi].state = SLOT_INFLIGHT;   // reusable once the streams have drained (h264r_wait)
    ctx->queue.clear();
    const int rc = run_waves(ctx, true, false, nullptr, nullptr, true);
    if (rc == H264R_OK) CU(cudaEventRecord(ctx->table_ev[ctx->table_idx], ctx->stream));
    return rc;
}

int h264r_wait(h264r_ctx* ctx, h264r_frame f)
{
    if (!ctx) return H264R_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (f >= 0) {
        // one frame: the wave that produces it (pictures queued behind it keep running; staging slots stay in flight)
        if (f >= (int)ctx->frames.size() || !ctx->frames[f].dev) return H264R_ERR_INVALID;
        if (ctx->frames[f].ready) CU(cudaEventSynchronize(ctx->frames[f].ready));
        return H264R_OK;
    }
    CU(cudaStreamSynchronize(ctx->s_h2d));
    for (int gi = 0; gi < ctx->num_groups; ++gi) { CU(cudaStreamSynchronize(ctx->g_side[gi])); CU(cudaStreamSynchronize(ctx->g_main[gi])); }
    CU(cudaStreamSynchronize(ctx->s_d2h));
    for (Slot& t : ctx->slots) if (t.state == SLOT_INFLIGHT) t.state = SLOT_FREE;
    return H264R_OK;
}

int h264r_replay_last_flush(h264r_ctx* ctx, int iterations, int flags, float ms_out[6], int launches_out[6])
{
    if (!ctx || iterations <= 0) return H264R_ERR_INVALID;
    if (ctx->last_waves.empty() || ctx->filling >= 0 || !ctx->queue.empty()) return H264R_ERR_STATE;
    cudaSetDevice(ctx->device);
    int rc;
    if (flags & H264R_REPLAY_ASYNC) {
        for (int it = 0; it < iterations; ++it) {
            rc = run_waves(ctx, (flags & H264R_REPLAY_H2D) != 0, false, nullptr, nullptr, false);
            if (rc != H264R_OK) return rc;
        }
        return H264R_OK;
    }
    rc = h264r_wait(ctx, -1);
    if (rc != H264R_OK) return rc;
    float ms[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
    int launches[6] = { 0, 0, 0, 0, 0, 0 };
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int it = 0; it < iterations; ++it) {
        rc = run_waves(ctx, (flags & H264R_REPLAY_H2D) != 0, (flags & H264R_REPLAY_TIME_KERNELS) != 0, ms, launches, true);
        if (rc != H264R_OK) return rc;
    }
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaEventSynchronize(ctx->ev1));
    CU(cudaGetLastError());
    CU(cudaEventElapsedTime(&ms[0], ctx->ev0, ctx->ev1));
    if (ms_out) for (int i = 0; i < 6; ++i) ms_out[i] = ms[i];
    if (launches_out) for (int i = 0; i < 6; ++i) launches_out[i] = launches[i];
    return H264R_OK;
}

void* h264r_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void h264r_host_free(void* p) { if (p) cudaFreeHost(p); }

int h264r_frame_download_async(h264r_ctx* ctx, h264r_frame f, uint8_t* y, uint8_t* cb, uint8_t* cr, int pitch_y, int pitch_c)
{
    if (!ctx || f < 0 || f >= (int)ctx->frames.size() || !ctx->frames[f].dev || !y || !cb || !cr) return H264R_ERR_INVALID;
    cudaSetDevice(ctx->device);
    const FrameGeom& g = ctx->geom;
    const int w = g.width_mbs * 16, h = g.height_mbs * 16;
    if (pitch_y < w || pitch_c < w / 2) return H264R_ERR_INVALID;
    const uint8_t* d = ctx->frames[f].dev;
    if (ctx->frames[f].ready) CU(cudaStreamWaitEvent(ctx->s_d2h, ctx->frames[f].ready, 0));
    { const int rc = copy_frame_d2h(ctx, ctx->s_d2h, d, y, cb, cr, pitch_y, pitch_c); if (rc != H264R_OK) return rc; }
    Frame& fr = ctx->frames[f];
    if (!fr.read_done) CU(cudaEventCreateWithFlags(&fr.read_done, cudaEventDisableTiming));
    CU(cudaEventRecord(fr.read_done, ctx->s_d2h));
    fr.pending_read = true;
    ctx->stats.d2h_bytes += (uint64_t)w * h * 3 / 2;
    return H264R_OK;
}

int h264r_frame_download_cropped(h264r_ctx* ctx, h264r_frame f, int crop_left, int crop_right, int crop_top, int crop_bottom,
                                 uint8_t* y, uint8_t* cb, uint8_t* cr, int pitch_y, int pitch_c)
{
    if (!ctx || f < 0 || f >= (int)ctx->frames.size() || !ctx->frames[f].dev || !y || !cb || !cr) return H264R_ERR_INVALID;
    cudaSetDevice(ctx->device);
    const FrameGeom& g = ctx->geom;
    const int W = g.width_mbs * 16, H = g.height_mbs * 16;
    if (crop_left < 0 || crop_right < 0 || crop_top < 0 || crop_bottom < 0 || ((crop_left | crop_right | crop_top | crop_bottom) & 1) ||
        crop_left + crop_right >= W || crop_top + crop_bottom >= H) return H264R_ERR_INVALID;
    const int w = W - crop_left - crop_right, h = H - crop_top - crop_bottom;
    if (pitch_y < w || pitch_c < w / 2) return H264R_ERR_INVALID;
    Frame& fr = ctx->frames[f];
    const uint8_t* d = fr.dev;
    if (fr.ready) CU(cudaStreamWaitEvent(ctx->s_d2h, fr.ready, 0));
    CU(cudaMemcpy2DAsync(y, pitch_y, d + (size_t)crop_top * g.pitch_y + crop_left, g.pitch_y, w, h, cudaMemcpyDeviceToHost, ctx->s_d2h));
    const size_t coff = (size_t)(crop_top / 2) * g.pitch_c + crop_left / 2;
    CU(cudaMemcpy2DAsync(cb, pitch_c, d + g.off_cb + coff, g.pitch_c, w / 2, h / 2, cudaMemcpyDeviceToHost, ctx->s_d2h));
    CU(cudaMemcpy2DAsync(cr, pitch_c, d + g.off_cr + coff, g.pitch_c, w / 2, h / 2, cudaMemcpyDeviceToHost, ctx->s_d2h));
    if (!fr.read_done) CU(cudaEventCreateWithFlags(&fr.read_done, cudaEventDisableTiming));
    CU(cudaEventRecord(fr.read_done, ctx->s_d2h));
    fr.pending_read = true;
    CU(cudaStreamSynchronize(ctx->s_d2h));
    ctx->stats.d2h_bytes += (uint64_t)w * h * 3 / 2;
    return H264R_OK;
}

int h264r_frame_download(h264r_ctx* ctx, h264r_frame f, uint8_t* y, uint8_t* cb, uint8_t* cr, int pitch_y, int pitch_c)
{
    if (!ctx || f < 0 || f >= (int)ctx->frames.size() || !ctx->frames[f].dev || !y || !cb || !cr) return H264R_ERR_INVALID;
    cudaSetDevice(ctx->device);
    const FrameGeom& g = ctx->geom;
    const int w = g.width_mbs * 16, h = g.height_mbs * 16;
    if (pitch_y < w || pitch_c < w / 2) return H264R_ERR_INVALID;
    const uint8_t* d = ctx->frames[f].dev;
    { const int rc = copy_frame_d2h(ctx, ctx->stream, d, y, cb, cr, pitch_y, pitch_c); if (rc != H264R_OK) return rc; }
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += (uint64_t)w * h * 3 / 2;
    return H264R_OK;
}

int h264r_frame_upload(h264r_ctx* ctx, h264r_frame f, const uint8_t* y, const uint8_t* cb, const uint8_t* cr, int pitch_y, int pitch_c)
{
    if (!ctx || !frame_ok(ctx, f) || !y || !cb || !cr) return H264R_ERR_INVALID;
    cudaSetDevice(ctx->device);
    const FrameGeom& g = ctx->geom;
    const int w = g.width_mbs * 16, h = g.height_mbs * 16;
    if (pitch_y < w || pitch_c < w / 2) return H264R_ERR_INVALID;
    uint8_t* d = ctx->frames[f].dev;
    CU(cudaMemcpy2DAsync(d, g.pitch_y, y, pitch_y, w, h, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpy2DAsync(d + g.off_cb, g.pitch_c, cb, pitch_c, w / 2, h / 2, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpy2DAsync(d + g.off_cr, g.pitch_c, cr, pitch_c, w / 2, h / 2, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stats.h2d_bytes += (uint64_t)w * h * 3 / 2;
    return H264R_OK;
}

int h264r_get_stats(h264r_ctx* ctx, h264r_stats* out)
{
    if (!ctx || !out) return H264R_ERR_INVALID;
    *out = ctx->stats;
    return H264R_OK;
}

} // extern "C"
