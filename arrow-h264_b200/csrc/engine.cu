// Host runtime behind the C ABI of include/h264recon.h: frame pool, pinned staging slots, dependency waves,
// batched launches.  C++ host code + CUDA runtime only (no PyTorch, no NCCL: streams/GOPs are independent).
//
// Threading (include/h264recon.h): h264r_picture_begin / _update / _fill / _submit are thread-safe -- one parser thread
// per stream fills its own picture, the slot table and the submission queue sit behind one mutex that is never held
// while a picture is filled or a CUDA call blocks.  Everything else belongs to the one thread that owns the context.
#include "device_types.h"
#include "h264recon_bench.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <deque>
#include <mutex>
#include <utility>
#include <vector>

using namespace h264r;

static_assert(sizeof(h264r_mb) == 32, "h264r_mb must be 32 bytes");
static_assert(sizeof(h264r_motion_entry) == 12, "h264r_motion_entry must be 12 bytes");
static_assert(sizeof(h264r_mb_motion) == 192, "h264r_mb_motion must be 192 bytes");
static_assert(sizeof(DeblockDesc) == 64, "DeblockDesc must be 64 bytes");
static_assert(sizeof(h264r_slice) % 16 == 0, "h264r_slice must keep 16-byte alignment in arrays");

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Frame {
    uint8_t* dev = nullptr;
    FrameMaps* maps = nullptr;                // device: TMA descriptors of this frame (entry of h264r_ctx::d_maps)
    bool used = false;
    int write_wave = -1, read_wave = -1;      // bookkeeping inside one flush
    cudaEvent_t ready = nullptr;              // ev_done of the wave that last wrote this frame (owned by the event ring)
    cudaEvent_t read_done = nullptr;          // recorded on the D2H stream after the last asynchronous download
    bool pending_read = false;                // a download was enqueued since the frame was last (re)written
};

enum SlotState { SLOT_FREE = 0, SLOT_FILLING, SLOT_QUEUED, SLOT_INFLIGHT };

// One staging slot = the description of one picture: pinned host memory and its HBM twin, same layout:
//   [ h264r_mb x nmb | h264r_slice x max_slices | stream words ]
struct Slot {
    uint8_t* host = nullptr;                  // pinned
    uint8_t* dev = nullptr;
    DeblockDesc* dev_desc = nullptr;          // device only: output of the deblock pre-pass
    int16_t* dev_resid = nullptr;             // device only: residual plane [nmb][384]
    uint32_t* dev_mb_done = nullptr;          // device only: per-MB epoch stamps of the sparse intra kernel
    uint32_t* dev_intra_list = nullptr;       // device only: [nmb + 1] = count, then the addresses of the picture's intra MBs
    uint64_t* dev_mbox = nullptr;             // device only: row-to-row mailboxes [nmb][24]
    SlotState state = SLOT_FREE;
    h264r_pic_params pp;
    h264r_frame dst = -1;
    uint32_t stream_words = 0;
    int all_intra = 0;
    int wave = 0;
    cudaEvent_t ev_h2d = nullptr;             // copy of this description to HBM done: the pinned staging can be refilled
    cudaEvent_t ev_done = nullptr;            // kernels that read the HBM twin done: the twin can be overwritten
};

struct WaveCopy { int slot; size_t head_bytes, stream_bytes; cudaEvent_t prev_done; };

struct WaveRecord {
    WaveLaunch launch;
    std::vector<WaveCopy> copies;             // H2D copies of the picture descriptions of this wave
    std::vector<int> dst_frames;              // frames written by this wave
    cudaEvent_t ev_h2d = nullptr;             // recorded on the H2D stream after the wave's copies
    cudaEvent_t ev_done = nullptr;            // recorded on the compute stream after the wave's kernels
    cudaEvent_t ev_side = nullptr;            // recorded on the side stream after the wave's description-only kernels
};

struct TableRegion { size_t begin, end; cudaEvent_t done; };

} // namespace

struct h264r_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;            // compute: inter, intra, deblock of every wave, in order
    cudaStream_t s_side = nullptr;            // kernels that depend on nothing but the picture description (residual, deblock
                                              // descriptors): they run ahead of, and underneath, the latency-bound
                                              // wavefront kernels of earlier waves
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t d2h_waited = nullptr;         // last event the D2H stream was made to wait for (frames of one wave share it)
    // events: a ring, taken in order; an event is reused only after kEventRing later takes, long after its work is done
    // (waiting for a re-recorded event is merely conservative)
    std::vector<cudaEvent_t> event_ring;
    size_t event_next = 0;
    std::vector<cudaEvent_t> timer_events;    // H264R_REPLAY_TIME_KERNELS
    h264r_seq_params seq;
    FrameGeom geom;
    int nmb = 0;
    size_t off_mbs = 0, off_slices = 0, off_stream = 0, slot_bytes = 0;
    uint32_t stream_capacity = 0;             // words
    std::vector<Frame> frames;
    uint8_t* dummy_frame = nullptr;           // what unused reference slots point at (never written)
    FrameMaps* d_maps = nullptr;              // [max_frames + 1] (the last one: the dummy frame)
    void* encode_tiled = nullptr;             // cuTensorMapEncodeTiled, through cudaGetDriverEntryPoint (no libcuda link)
    std::mutex mu;                            // slots[].state, free_slots, inflight, queue
    std::vector<Slot> slots;
    std::vector<int> free_slots;
    std::deque<int> inflight;                 // flushed slots in flush order: their staging frees up in this order
    std::vector<int> queue;                   // slot indexes in submission order
    uint32_t epoch = 0;                       // launch-sequence counter (mailboxes, DevPicture::mb_done stamps)
    DevPicture* h_pics = nullptr;             // pinned ring of picture tables, one contiguous region per flush
    DevPicture* d_pics = nullptr;
    size_t table_entries = 0, table_next = 0;
    std::deque<TableRegion> table_busy;       // regions of flushes that may still be running
    int* d_tickets = nullptr;                 // 64 ints, zeroed per wave on the compute stream
    uint32_t* d_wave_words = nullptr;         // ring of per-wave device words (largest intra MB count of the wave)
    size_t wave_word_next = 0;
    uint32_t* h_err = nullptr;                // host-mapped error word the kernels OR into
    uint32_t* d_err = nullptr;
    std::vector<WaveRecord> last_waves;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    h264r_stats stats;
    char cuda_err[256];
};

namespace {

enum { kEventRing = 8192 };

int cuda_fail(h264r_ctx* c, cudaError_t e, const char* what)
{
    snprintf(c->cuda_err, sizeof(c->cuda_err), "%s: %s", what, cudaGetErrorString(e));
    return H264R_ERR_CUDA;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call); } while (0)

bool frame_ok(const h264r_ctx* c, h264r_frame f) { return f >= 0 && f < (int)c->frames.size() && c->frames[f].used; }

bool pic_params_ok(const h264r_ctx* c, const h264r_pic_params* pp)
{
    if (pp->num_slices <= 0 || pp->num_slices > c->seq.max_slices_per_picture) return false;
    if (pp->num_ref_frames < 0 || pp->num_ref_frames > H264R_MAX_REFS) return false;
    for (int i = 0; i < pp->num_ref_frames; ++i)
        if (pp->ref_frames[i] < 0 || pp->ref_frames[i] >= (int)c->frames.size() || !c->frames[pp->ref_frames[i]].dev) return false;
    // frame pictures reference frames, field pictures reference fields (the context holds pictures of one kind)
    if (pp->structure < H264R_FRAME || pp->structure > H264R_BOTTOM_FIELD) return false;
    for (int i = 0; i < pp->num_ref_frames; ++i)
        if (pp->structure == H264R_FRAME ? pp->ref_structure[i] != H264R_FRAME
                                         : (pp->ref_structure[i] != H264R_TOP_FIELD && pp->ref_structure[i] != H264R_BOTTOM_FIELD)) return false;
    return true;
}

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

// TMA descriptors of the frame at `dev` -> d_maps[index] (device memory; the kernels pass the address to cp.async.bulk.tensor)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int upload_frame_maps(h264r_ctx* ctx, uint8_t* dev, size_t index)
{
    alignas(64) FrameMaps m;
    memset(&m, 0, sizeof(m));
    const FrameGeom& g = ctx->geom;
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
    if (enc && (g.pitch_c % 16) == 0) {
        const cuuint64_t dimY[2] = { (cuuint64_t)g.pitch_y, (cuuint64_t)g.height_mbs * 16 }, strideY[1] = { (cuuint64_t)g.pitch_y };
        const cuuint32_t boxY[2] = { 32, 13 }, one[3] = { 1, 1, 1 };
        const cuuint64_t dimC[3] = { (cuuint64_t)g.pitch_c, (cuuint64_t)g.height_mbs * 8, 2 };
        const cuuint64_t strideC[2] = { (cuuint64_t)g.pitch_c, (cuuint64_t)(g.off_cr - g.off_cb) };
        const cuuint32_t boxC[3] = { 32, 5, 2 };
        const CUresult r0 = enc(&m.luma, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dev, dimY, strideY, boxY, one, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const CUresult r1 = enc(&m.chroma, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, dev + g.off_cb, dimC, strideC, boxC, one, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        m.ok = r0 == CUDA_SUCCESS && r1 == CUDA_SUCCESS;
    }
    CU(cudaMemcpy(ctx->d_maps + index, &m, sizeof(m), cudaMemcpyHostToDevice));
    return H264R_OK;
}

// cudaEventQuery without leaving cudaErrorNotReady behind as the thread's "last error"
bool event_done(cudaEvent_t ev)
{
    const cudaError_t e = cudaEventQuery(ev);
    if (e == cudaSuccess) return true;
    cudaGetLastError();
    return false;
}

cudaEvent_t take_event(h264r_ctx* c)
{
    if (c->event_ring.size() < kEventRing) {
        cudaEvent_t ev = nullptr;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        c->event_ring.push_back(ev);
        return ev;
    }
    cudaEvent_t ev = c->event_ring[c->event_next];
    c->event_next = (c->event_next + 1) % kEventRing;
    if (ev == c->d2h_waited) c->d2h_waited = nullptr;      // the event gets a new meaning
    return ev;
}

// Runs the recorded waves of the last flush on three streams:
//   H2D stream  : the picture descriptions of a wave, ahead of the kernels (a copy waits for the kernels that last read
//                 the HBM twin it overwrites);
//   side stream : the kernels that need nothing but the description -- residual, deblock descriptors.  They run ahead,
//                 underneath the latency-bound wavefront kernels of earlier waves;
//   compute     : inter, intra (row wavefront / sparse), deblock of the wave, after its side kernels.
// With time_kernels everything runs on the compute stream, one kernel at a time, bracketed by events.
int run_waves(h264r_ctx* ctx, bool h2d, bool time_kernels, float* ms_kernel, int* launches)
{
    size_t timer_used = 0;
    struct Pending { int kind; size_t ev; };
    std::vector<Pending> pend;
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
        cudaStream_t main = ctx->stream, side = time_kernels ? ctx->stream : ctx->s_side;
        if (h2d) {
            CU(cudaStreamWaitEvent(ctx->s_h2d, rec.ev_done, 0));          // replays: the previous run of this record; no-op the first time
            cudaEvent_t waited = nullptr;                                  // the slots of a wave mostly share their previous wave
            for (const WaveCopy& c : rec.copies) {
                Slot& s = ctx->slots[c.slot];
                if (c.prev_done && c.prev_done != waited) { CU(cudaStreamWaitEvent(ctx->s_h2d, c.prev_done, 0)); waited = c.prev_done; }
                // [mbs | slices used] and [stream used]: one copy when the unused slice entries between them are small
                const size_t gap = ctx->off_stream - c.head_bytes;
                if (c.stream_bytes && gap <= 32768) {
                    CU(cudaMemcpyAsync(s.dev, s.host, ctx->off_stream + c.stream_bytes, cudaMemcpyHostToDevice, ctx->s_h2d));
                    ctx->stats.h2d_bytes += ctx->off_stream + c.stream_bytes;
                } else {
                    CU(cudaMemcpyAsync(s.dev, s.host, c.head_bytes, cudaMemcpyHostToDevice, ctx->s_h2d));
                    if (c.stream_bytes) CU(cudaMemcpyAsync(s.dev + ctx->off_stream, s.host + ctx->off_stream, c.stream_bytes, cudaMemcpyHostToDevice, ctx->s_h2d));
                    ctx->stats.h2d_bytes += c.head_bytes + c.stream_bytes;
                }
            }
            CU(cudaEventRecord(rec.ev_h2d, ctx->s_h2d));
            CU(cudaStreamWaitEvent(side, rec.ev_h2d, 0));
        }
        rec.launch.epoch = ++ctx->epoch;
        // side kernels: their outputs (residual plane, deblock descriptors) are per picture slot; the previous user of the
        // slot's device buffers -- this record's previous run, or an earlier picture in the same slot -- must be done
        if (!time_kernels) {
            CU(cudaStreamWaitEvent(side, rec.ev_done, 0));
            cudaEvent_t waited = nullptr;
            for (const WaveCopy& c : rec.copies)
                if (c.prev_done && c.prev_done != waited) { CU(cudaStreamWaitEvent(side, c.prev_done, 0)); waited = c.prev_done; }
        }
        if (rec.launch.any_inter) CU(cudaMemsetAsync(rec.launch.wave_max, 0, sizeof(uint32_t), side));
        { const int rc = launch(rec, KERNEL_LIST, side); if (rc != H264R_OK) return rc; }
        { const int rc = launch(rec, KERNEL_RESID, side); if (rc != H264R_OK) return rc; }
        { const int rc = launch(rec, KERNEL_DBPREP, side); if (rc != H264R_OK) return rc; }
        if (!time_kernels) {
            CU(cudaEventRecord(rec.ev_side, side));
            CU(cudaStreamWaitEvent(main, rec.ev_side, 0));
        }
        // write-after-read: a frame still being downloaded (asynchronously, on the D2H stream) is not overwritten
        for (int f : rec.dst_frames) {
            Frame& fr = ctx->frames[f];
            if (fr.pending_read) { CU(cudaStreamWaitEvent(main, fr.read_done, 0)); fr.pending_read = false; }
        }
        CU(cudaMemsetAsync(rec.launch.tickets, 0, sizeof(int) * 64, main));      // the wave's ticket counters
        { const int rc = launch(rec, KERNEL_INTER, main); if (rc != H264R_OK) return rc; }
        { const int rc = launch(rec, KERNEL_INTRA, main); if (rc != H264R_OK) return rc; }
        { const int rc = launch(rec, KERNEL_DEBLOCK, main); if (rc != H264R_OK) return rc; }
        CU(cudaGetLastError());
        CU(cudaEventRecord(rec.ev_done, main));
        ctx->stats.waves += 1;
        ctx->stats.pictures += (uint64_t)rec.launch.num_pics;
        ctx->stats.macroblocks += (uint64_t)rec.launch.num_pics * ctx->nmb;
    }
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

// `n` contiguous entries of the picture-table ring that no running flush reads
int take_table(h264r_ctx* ctx, size_t n, size_t* begin)
{
    if (n > ctx->table_entries) return H264R_ERR_NOMEM;
    size_t b = ctx->table_next;
    if (b + n > ctx->table_entries) b = 0;
    // flushes finish in order: retire the finished ones, then wait for the oldest until nothing overlaps [b, b + n)
    while (!ctx->table_busy.empty() && event_done(ctx->table_busy.front().done)) ctx->table_busy.pop_front();
    auto overlaps = [&]() {
        for (const TableRegion& r : ctx->table_busy) if (!(r.end <= b || r.begin >= b + n)) return true;
        return false;
    };
    while (overlaps()) {
        CU(cudaEventSynchronize(ctx->table_busy.front().done));
        ctx->table_busy.pop_front();
    }
    *begin = b;
    ctx->table_next = b + n;
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
    case H264R_ERR_INVALID: return "invalid argument or picture description";
    case H264R_ERR_UNSUPPORTED: return "unsupported stream feature (8-bit 4:2:0 frame pictures only)";
    case H264R_ERR_NOMEM: return "out of frames, staging slots or device memory";
    case H264R_ERR_CUDA: return "CUDA runtime error";
    case H264R_ERR_STATE: return "call order violated";
    case H264R_ERR_NODEVICE: return "no CUDA device (this engine has no CPU fallback)";
    default: return "unknown error";
    }
}

const char* h264r_last_cuda_error(h264r_ctx* ctx) { return ctx ? ctx->cuda_err : ""; }

void h264r_destroy(h264r_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (Frame& f : ctx->frames) { if (f.dev) cudaFree(f.dev); if (f.read_done) cudaEventDestroy(f.read_done); }
    if (!ctx->slots.empty()) {
        if (ctx->slots[0].host) cudaFreeHost(ctx->slots[0].host);
        if (ctx->slots[0].dev) cudaFree(ctx->slots[0].dev);
        if (ctx->slots[0].dev_desc) cudaFree(ctx->slots[0].dev_desc);
        if (ctx->slots[0].dev_resid) cudaFree(ctx->slots[0].dev_resid);
        if (ctx->slots[0].dev_mb_done) cudaFree(ctx->slots[0].dev_mb_done);
        if (ctx->slots[0].dev_intra_list) cudaFree(ctx->slots[0].dev_intra_list);
        if (ctx->slots[0].dev_mbox) cudaFree(ctx->slots[0].dev_mbox);
    }
    if (ctx->dummy_frame) cudaFree(ctx->dummy_frame);
    if (ctx->d_maps) cudaFree(ctx->d_maps);
    if (ctx->h_pics) cudaFreeHost(ctx->h_pics);
    if (ctx->d_pics) cudaFree(ctx->d_pics);
    if (ctx->d_tickets) cudaFree(ctx->d_tickets);
    if (ctx->d_wave_words) cudaFree(ctx->d_wave_words);
    if (ctx->h_err) cudaFreeHost(ctx->h_err);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t ev : ctx->event_ring) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->timer_events) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->s_side) cudaStreamDestroy(ctx->s_side);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    delete ctx;
}

int h264r_create(h264r_ctx** out, int device, const h264r_seq_params* sp)
{
    if (!out || !sp || sp->width_mbs <= 0 || sp->height_mbs <= 0 || sp->max_frames <= 0 ||
        sp->max_pictures_in_flight <= 0 || sp->max_slices_per_picture <= 0 || sp->max_levels_per_picture < 0) return H264R_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return H264R_ERR_NODEVICE;
    if (device < 0 || device >= ndev) return H264R_ERR_INVALID;
    h264r_ctx* ctx = new h264r_ctx();
    ctx->device = device; ctx->seq = *sp; ctx->cuda_err[0] = 0;
    memset(&ctx->stats, 0, sizeof(ctx->stats));

    FrameGeom& g = ctx->geom;
    g.width_mbs = sp->width_mbs; g.height_mbs = sp->height_mbs;
    g.pitch_y = sp->width_mbs * 16;            // tight: a whole frame is one contiguous 1.5*W*H block
    g.pitch_c = sp->width_mbs * 8;
    g.off_cb = (size_t)g.pitch_y * sp->height_mbs * 16;
    g.off_cr = g.off_cb + (size_t)g.pitch_c * sp->height_mbs * 8;
    g.bytes  = align_up(g.off_cr + (size_t)g.pitch_c * sp->height_mbs * 8, 256);
    ctx->nmb = sp->width_mbs * sp->height_mbs;
    const size_t nmb = (size_t)ctx->nmb, nslots = (size_t)sp->max_pictures_in_flight;

    ctx->off_mbs = 0;
    ctx->off_slices = sizeof(h264r_mb) * nmb;                 // contiguous with the headers: [mbs | slices used] is one copy
    ctx->off_stream = align_up(ctx->off_slices + sizeof(h264r_slice) * sp->max_slices_per_picture, 256);
    // stream = levels + packed motion (worst case 16 entries of 3 words per MB)
    const size_t levels = sp->max_levels_per_picture > 0 ? (size_t)sp->max_levels_per_picture : (size_t)H264R_COEFFS_PER_MB * nmb;
    ctx->stream_capacity = (uint32_t)(levels + 48 * nmb);
    ctx->slot_bytes = align_up(ctx->off_stream + sizeof(uint32_t) * (size_t)ctx->stream_capacity, 256);
    ctx->frames.resize(sp->max_frames);
    ctx->slots.resize(nslots);
    ctx->table_entries = 4 * nslots;

    // every CUDA failure below takes the one cleanup path (h264r_destroy frees whatever exists)
    uint8_t* h_arena = nullptr; uint8_t* d_arena = nullptr; DeblockDesc* d_desc = nullptr; int16_t* d_resid = nullptr;
    uint32_t* d_done = nullptr; uint64_t* d_mbox = nullptr; uint32_t* d_list = nullptr;
    const size_t mbox_words = (size_t)24 * nmb;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
    bool alloc_phase = false;
    if (e == cudaSuccess) { alloc_phase = true; e = cudaHostAlloc((void**)&h_arena, ctx->slot_bytes * nslots, cudaHostAllocDefault); }
    if (h_arena) ctx->slots[0].host = h_arena;
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_arena, ctx->slot_bytes * nslots);
    if (d_arena) ctx->slots[0].dev = d_arena;
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_desc, sizeof(DeblockDesc) * nmb * nslots);
    if (d_desc) ctx->slots[0].dev_desc = d_desc;
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_resid, sizeof(int16_t) * H264R_COEFFS_PER_MB * nmb * nslots);
    if (d_resid) ctx->slots[0].dev_resid = d_resid;
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_done, sizeof(uint32_t) * nmb * nslots);
    if (d_done) ctx->slots[0].dev_mb_done = d_done;
    if (e == cudaSuccess) e = cudaMemset(d_done, 0, sizeof(uint32_t) * nmb * nslots);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_list, sizeof(uint32_t) * (nmb + 1) * nslots);
    if (d_list) ctx->slots[0].dev_intra_list = d_list;
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_wave_words, sizeof(uint32_t) * kEventRing);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_mbox, sizeof(uint64_t) * mbox_words * nslots);
    if (d_mbox) ctx->slots[0].dev_mbox = d_mbox;
    if (e == cudaSuccess) e = cudaMemset(d_mbox, 0, sizeof(uint64_t) * mbox_words * nslots);   // epoch 0 is never used
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->dummy_frame, g.bytes);
    if (e == cudaSuccess) e = cudaMemset(ctx->dummy_frame, 128, g.bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_maps, sizeof(FrameMaps) * ((size_t)sp->max_frames + 1));
    if (e == cudaSuccess) {
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ctx->encode_tiled, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) {
            ctx->encode_tiled = nullptr;                   // the TMA build then stays on its cp.async path
            cudaGetLastError();
        }
        if (upload_frame_maps(ctx, ctx->dummy_frame, (size_t)sp->max_frames) != H264R_OK) e = cudaErrorUnknown;
    }
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->h_pics, sizeof(DevPicture) * ctx->table_entries, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_pics, sizeof(DevPicture) * ctx->table_entries);
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_tickets, sizeof(int) * 64);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->h_err, sizeof(uint32_t), cudaHostAllocMapped);
    if (e == cudaSuccess) { *ctx->h_err = 0; e = cudaHostGetDevicePointer((void**)&ctx->d_err, ctx->h_err, 0); }
    if (e != cudaSuccess) {
        const bool nomem = alloc_phase && e == cudaErrorMemoryAllocation;
        cudaGetLastError();
        h264r_destroy(ctx);
        return nomem ? H264R_ERR_NOMEM : H264R_ERR_CUDA;
    }
    for (size_t i = 0; i < nslots; ++i) {
        Slot& s = ctx->slots[i];
        s.host = h_arena + ctx->slot_bytes * i;
        s.dev = d_arena + ctx->slot_bytes * i;
        s.dev_desc = d_desc + nmb * i;
        s.dev_resid = d_resid + (size_t)H264R_COEFFS_PER_MB * nmb * i;
        s.dev_mb_done = d_done + nmb * i;
        s.dev_intra_list = d_list + (nmb + 1) * i;
        s.dev_mbox = d_mbox + mbox_words * i;
        ctx->free_slots.push_back((int)(nslots - 1 - i));          // slot 0 is handed out first
    }
    *out = ctx;
    return H264R_OK;
}

int h264r_frame_alloc(h264r_ctx* ctx, h264r_frame* out)
{
    if (!ctx || !out) return H264R_ERR_INVALID;
    cudaSetDevice(ctx->device);
    for (size_t i = 0; i < ctx->frames.size(); ++i)
        if (!ctx->frames[i].used) {
            if (!ctx->frames[i].dev) {
                CU(cudaMalloc((void**)&ctx->frames[i].dev, ctx->geom.bytes));
                ctx->frames[i].maps = ctx->d_maps + i;
                const int rc = upload_frame_maps(ctx, ctx->frames[i].dev, i);
                if (rc != H264R_OK) return rc;
            }
            ctx->frames[i].used = true;
            *out = (h264r_frame)i;
            return H264R_OK;
        }
    return H264R_ERR_NOMEM;
}

int h264r_frame_release(h264r_ctx* ctx, h264r_frame f)
{
    if (!ctx || !frame_ok(ctx, f)) return H264R_ERR_INVALID;
    // The memory is kept.  Pictures already queued or running that read the frame stay valid: whatever is allocated
    // into it next is written by a later wave of the same compute stream (write-after-read is stream order, inside one
    // flush the wave assignment), after the asynchronous downloads of the old content (Frame::read_done).
    ctx->frames[f].used = false;
    return H264R_OK;
}

int h264r_picture_begin(h264r_ctx* ctx, h264r_frame dst, const h264r_pic_params* pp, h264r_pic_buffers* out)
{
    if (!ctx || !pp || !out || !frame_ok(ctx, dst) || !pic_params_ok(ctx, pp)) return H264R_ERR_INVALID;
    int s = -1;
    {
        std::unique_lock<std::mutex> lock(ctx->mu);
        for (;;) {
            // the staging of a flushed picture is reusable once its host->device copy has drained (never wait for kernels)
            while (!ctx->inflight.empty() && event_done(ctx->slots[ctx->inflight.front()].ev_h2d)) {
                const cudaEvent_t ev = ctx->slots[ctx->inflight.front()].ev_h2d;          // one event per wave: take the whole wave
                while (!ctx->inflight.empty() && ctx->slots[ctx->inflight.front()].ev_h2d == ev) {
                    ctx->slots[ctx->inflight.front()].state = SLOT_FREE;
                    ctx->free_slots.push_back(ctx->inflight.front());
                    ctx->inflight.pop_front();
                }
            }
            if (!ctx->free_slots.empty()) {
                s = ctx->free_slots.back();
                ctx->free_slots.pop_back();
                ctx->slots[s].state = SLOT_FILLING;
                break;
            }
            if (ctx->inflight.empty()) return H264R_ERR_NOMEM;    // every slot is being filled or queued: the owner must flush
            const cudaEvent_t oldest = ctx->slots[ctx->inflight.front()].ev_h2d;
            lock.unlock();
            cudaSetDevice(ctx->device);
            if (cudaEventSynchronize(oldest) != cudaSuccess) return H264R_ERR_CUDA;
            lock.lock();
        }
    }
    Slot& sl = ctx->slots[s];
    sl.pp = *pp; sl.dst = dst; sl.stream_words = 0;
    out->mbs = reinterpret_cast<h264r_mb*>(sl.host + ctx->off_mbs);
    out->slices = reinterpret_cast<h264r_slice*>(sl.host + ctx->off_slices);
    out->stream = reinterpret_cast<uint32_t*>(sl.host + ctx->off_stream);
    out->stream_capacity = ctx->stream_capacity;
    out->picture = s;
    return H264R_OK;
}

static Slot* filling_slot(h264r_ctx* ctx, int32_t picture)
{
    if (picture < 0 || picture >= (int)ctx->slots.size()) return nullptr;
    Slot& s = ctx->slots[picture];
    return s.state == SLOT_FILLING ? &s : nullptr;       // only the thread that began the picture moves it out of this state
}

int h264r_picture_update(h264r_ctx* ctx, int32_t picture, const h264r_pic_params* pp)
{
    if (!ctx || !pp) return H264R_ERR_INVALID;
    Slot* s = filling_slot(ctx, picture);
    if (!s) return H264R_ERR_STATE;
    if (!pic_params_ok(ctx, pp)) return H264R_ERR_INVALID;
    s->pp = *pp;
    return H264R_OK;
}

int64_t h264r_picture_fill(h264r_ctx* ctx, int32_t picture, const h264r_mb* mbs, const h264r_mb_motion* motion,
                           const h264r_slice* slices, int num_slices, const h264r_level* levels, uint32_t num_levels)
{
    if (!ctx || !mbs || !motion || !slices || (num_levels && !levels)) return H264R_ERR_INVALID;
    Slot* s = filling_slot(ctx, picture);
    if (!s) return H264R_ERR_STATE;
    if (num_slices <= 0 || num_slices > ctx->seq.max_slices_per_picture) return H264R_ERR_INVALID;
    h264r_mb* out_mbs = reinterpret_cast<h264r_mb*>(s->host + ctx->off_mbs);
    uint32_t* stream = reinterpret_cast<uint32_t*>(s->host + ctx->off_stream);
    memcpy(s->host + ctx->off_slices, slices, sizeof(h264r_slice) * (size_t)num_slices);
    const int64_t words = h264r_pack_picture(ctx->nmb, mbs, motion, levels, num_levels, out_mbs, stream, ctx->stream_capacity);
    return words;
}

int h264r_picture_submit(h264r_ctx* ctx, int32_t picture, uint32_t stream_words)
{
    if (!ctx) return H264R_ERR_INVALID;
    Slot* sp = filling_slot(ctx, picture);
    if (!sp) return H264R_ERR_STATE;
    Slot& sl = *sp;
    int rc = H264R_OK;
    if (stream_words > ctx->stream_capacity) rc = H264R_ERR_INVALID;
    // O(slices): slice tables only.  The macroblocks are checked by the kernels that read them.
    const h264r_slice* slices = reinterpret_cast<const h264r_slice*>(sl.host + ctx->off_slices);
    int all_intra = 1;
    for (int k = 0; k < sl.pp.num_slices && rc == H264R_OK; ++k) {
        if (slices[k].slice_type > H264R_I_SLICE) rc = H264R_ERR_UNSUPPORTED;    // SP / SI
        if (slices[k].slice_type != H264R_I_SLICE) all_intra = 0;
        for (int list = 0; list < 2; ++list)
            for (int i = 0; i < H264R_MAX_REFS; ++i)
                if (slices[k].ref_pic_list[list][i] >= sl.pp.num_ref_frames) rc = H264R_ERR_INVALID;
    }
    sl.stream_words = stream_words;
    sl.all_intra = all_intra;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (rc != H264R_OK) { sl.state = SLOT_FREE; ctx->free_slots.push_back(picture); return rc; }
    sl.state = SLOT_QUEUED;
    ctx->queue.push_back(picture);
    return H264R_OK;
}

int h264r_flush(h264r_ctx* ctx)
{
    if (!ctx) return H264R_ERR_INVALID;
    std::vector<int> queue;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        queue.swap(ctx->queue);
    }
    if (queue.empty()) return H264R_OK;
    cudaSetDevice(ctx->device);

    // ---- dependency waves: RAW on reference frames, WAR/WAW on the destination frame ----
    for (Frame& f : ctx->frames) { f.write_wave = -1; f.read_wave = -1; }
    for (int qi : queue) {
        Slot& s = ctx->slots[qi];
        int w = 0;
        for (int i = 0; i < s.pp.num_ref_frames; ++i) w = std::max(w, ctx->frames[s.pp.ref_frames[i]].write_wave + 1);
        Frame& d = ctx->frames[s.dst];
        w = std::max(w, std::max(d.write_wave, d.read_wave) + 1);
        s.wave = w;
        d.write_wave = w;
        for (int i = 0; i < s.pp.num_ref_frames; ++i) {
            Frame& r = ctx->frames[s.pp.ref_frames[i]];
            r.read_wave = std::max(r.read_wave, w);
        }
    }
    std::vector<int> order(queue);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ctx->slots[a].wave < ctx->slots[b].wave; });
    std::vector<int> rec_begin;
    for (size_t k = 0; k < order.size(); ++k)
        if (k == 0 || ctx->slots[order[k]].wave != ctx->slots[order[k - 1]].wave) rec_begin.push_back((int)k);
    rec_begin.push_back((int)order.size());

    // ---- device picture table (one upload for all records) ----
    size_t table_off = 0;
    { const int rc = take_table(ctx, order.size(), &table_off); if (rc != H264R_OK) return rc; }
    DevPicture* const h_table = ctx->h_pics + table_off;
    DevPicture* const d_table = ctx->d_pics + table_off;
    for (size_t k = 0; k < order.size(); ++k) {
        Slot& s = ctx->slots[order[k]];
        DevPicture& p = h_table[k];
        memset(&p, 0, sizeof(p));
        p.mbs = reinterpret_cast<const h264r_mb*>(s.dev + ctx->off_mbs);
        p.slices = reinterpret_cast<const h264r_slice*>(s.dev + ctx->off_slices);
        p.stream = reinterpret_cast<const uint32_t*>(s.dev + ctx->off_stream);
        p.dst = ctx->frames[s.dst].dev;
        for (int i = 0; i < H264R_MAX_REFS; ++i) {
            const bool have = i < s.pp.num_ref_frames;
            p.ref[i] = have ? ctx->frames[s.pp.ref_frames[i]].dev : ctx->dummy_frame;
            p.ref_maps[i] = have ? ctx->frames[s.pp.ref_frames[i]].maps : ctx->d_maps + ctx->seq.max_frames;
        }
        p.desc = s.dev_desc;
        p.resid = s.dev_resid;
        p.mbox = s.dev_mbox;
        p.mb_done = s.dev_mb_done;
        p.intra_count = s.dev_intra_list; p.intra_list = s.dev_intra_list + 1;
        p.stream_words = s.stream_words;
        p.num_slices = s.pp.num_slices; p.num_refs = s.pp.num_ref_frames;
        p.run_deblock = s.pp.run_deblock; p.all_intra = s.all_intra;
        p.direct8x8 = s.pp.direct_8x8_inference_flag != 0;
        p.field = s.pp.structure != H264R_FRAME;
        p.chroma_dy = s.pp.structure == H264R_TOP_FIELD ? -2 : (s.pp.structure == H264R_BOTTOM_FIELD ? 2 : 0);
        for (int i = 0; i < s.pp.num_ref_frames; ++i)
            if (p.field && s.pp.ref_structure[i] != s.pp.structure) p.ref_opposite |= 1u << i;
    }
    // on the H2D stream, ahead of the descriptions: every kernel of the flush (side and compute stream) follows its wave's ev_h2d
    CU(cudaMemcpyAsync(d_table, h_table, sizeof(DevPicture) * order.size(), cudaMemcpyHostToDevice, ctx->s_h2d));
    ctx->stats.h2d_bytes += sizeof(DevPicture) * order.size();

    // ---- per record: what to copy, what to launch ----
    ctx->last_waves.clear();
    for (size_t r = 0; r + 1 < rec_begin.size(); ++r) {
        const int b = rec_begin[r], e = rec_begin[r + 1];
        WaveRecord rec;
        WaveLaunch& L = rec.launch;
        L.pics = d_table + b; L.num_pics = e - b; L.tickets = ctx->d_tickets; L.err = ctx->d_err; L.geom = ctx->geom;
        L.any_inter = L.any_deblock = L.any_intra_rows = L.any_field = 0; L.epoch = 0;
        L.wave_max = ctx->d_wave_words + ctx->wave_word_next;
        ctx->wave_word_next = (ctx->wave_word_next + 1) % kEventRing;
        rec.ev_h2d = take_event(ctx); rec.ev_done = take_event(ctx); rec.ev_side = take_event(ctx);
        if (!rec.ev_h2d || !rec.ev_done || !rec.ev_side) return H264R_ERR_CUDA;
        for (int k = b; k < e; ++k) {
            Slot& s = ctx->slots[order[k]];
            rec.copies.push_back({ order[k], ctx->off_slices + sizeof(h264r_slice) * (size_t)s.pp.num_slices,
                                   sizeof(uint32_t) * (size_t)s.stream_words, s.ev_done });
            L.any_inter |= !s.all_intra; L.any_intra_rows |= s.all_intra; L.any_deblock |= s.pp.run_deblock;
            L.any_field |= s.pp.structure != H264R_FRAME;
            Frame& d = ctx->frames[s.dst];
            d.ready = rec.ev_done;
            rec.dst_frames.push_back(s.dst);
        }
        ctx->last_waves.push_back(rec);
    }
    const int rc = run_waves(ctx, true, false, nullptr, nullptr);
    if (rc != H264R_OK) return rc;
    cudaEvent_t flush_done = ctx->last_waves.back().ev_done;
    ctx->table_busy.push_back({ table_off, table_off + order.size(), flush_done });
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        for (const WaveRecord& rec : ctx->last_waves)
            for (const WaveCopy& c : rec.copies) {
                Slot& s = ctx->slots[c.slot];
                s.ev_h2d = rec.ev_h2d; s.ev_done = rec.ev_done;
                s.state = SLOT_INFLIGHT;
                ctx->inflight.push_back(c.slot);
            }
    }
    return H264R_OK;
}

static int check_device_errors(h264r_ctx* ctx)
{
    if (*(volatile uint32_t*)ctx->h_err == 0) return H264R_OK;
    snprintf(ctx->cuda_err, sizeof(ctx->cuda_err), "picture description outside its domain (bits %u: 1 header, 2 motion, 4 level)", *ctx->h_err);
    *ctx->h_err = 0;
    return H264R_ERR_INVALID;
}

int h264r_wait(h264r_ctx* ctx, h264r_frame f)
{
    if (!ctx) return H264R_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (f >= 0) {
        // one frame: the wave that produces it (pictures queued behind it keep running; staging slots stay in flight)
        if (f >= (int)ctx->frames.size() || !ctx->frames[f].dev) return H264R_ERR_INVALID;
        if (ctx->frames[f].ready) CU(cudaEventSynchronize(ctx->frames[f].ready));
        return check_device_errors(ctx);
    }
    CU(cudaStreamSynchronize(ctx->s_h2d));
    CU(cudaStreamSynchronize(ctx->s_side));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaStreamSynchronize(ctx->s_d2h));
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        for (int i : ctx->inflight) { ctx->slots[i].state = SLOT_FREE; ctx->free_slots.push_back(i); }
        ctx->inflight.clear();
    }
    ctx->table_busy.clear();
    return check_device_errors(ctx);
}

// ---- include/h264recon_bench.h ----

const char* h264r_bench_kernel_name(int kind) { return wave_kernel_name(kind); }

int h264r_replay_last_flush(h264r_ctx* ctx, int iterations, int flags, float ms_out[1 + H264R_BENCH_MAX_KERNELS],
                            int launches_out[1 + H264R_BENCH_MAX_KERNELS])
{
    if (!ctx || iterations <= 0) return H264R_ERR_INVALID;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        if (ctx->last_waves.empty() || !ctx->queue.empty()) return H264R_ERR_STATE;
    }
    cudaSetDevice(ctx->device);
    int rc;
    if (flags & H264R_REPLAY_ASYNC) {
        for (int it = 0; it < iterations; ++it) {
            rc = run_waves(ctx, (flags & H264R_REPLAY_H2D) != 0, false, nullptr, nullptr);
            if (rc != H264R_OK) return rc;
        }
        return H264R_OK;
    }
    rc = h264r_wait(ctx, -1);
    if (rc != H264R_OK) return rc;
    float ms[1 + H264R_BENCH_MAX_KERNELS] = { 0.f };
    int launches[1 + H264R_BENCH_MAX_KERNELS] = { 0 };
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int it = 0; it < iterations; ++it) {
        rc = run_waves(ctx, (flags & H264R_REPLAY_H2D) != 0, (flags & H264R_REPLAY_TIME_KERNELS) != 0, ms, launches);
        if (rc != H264R_OK) return rc;
    }
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaEventSynchronize(ctx->ev1));
    CU(cudaGetLastError());
    CU(cudaEventElapsedTime(&ms[0], ctx->ev0, ctx->ev1));
    if (ms_out) for (int i = 0; i <= H264R_BENCH_MAX_KERNELS; ++i) ms_out[i] = ms[i];
    if (launches_out) for (int i = 0; i <= H264R_BENCH_MAX_KERNELS; ++i) launches_out[i] = launches[i];
    return H264R_OK;
}

// ---- downloads ----

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
    if (ctx->frames[f].ready && ctx->frames[f].ready != ctx->d2h_waited) {
        CU(cudaStreamWaitEvent(ctx->s_d2h, ctx->frames[f].ready, 0));
        ctx->d2h_waited = ctx->frames[f].ready;
    }
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

// dpb_split_field / dpb_combine_field_yuv (framebuf/dpb.cc) between a context of frames and a context of fields
int h264r_field_copy(h264r_ctx* frame_ctx, h264r_frame frame, h264r_ctx* field_ctx, h264r_frame field, int parity, int to_field)
{
    if (!frame_ctx || !field_ctx || frame_ctx == field_ctx || !frame_ok(frame_ctx, frame) || !frame_ok(field_ctx, field) ||
        parity < 0 || parity > 1) return H264R_ERR_INVALID;
    const FrameGeom& gf = frame_ctx->geom; const FrameGeom& gd = field_ctx->geom;
    if (frame_ctx->device != field_ctx->device || gf.width_mbs != gd.width_mbs || gf.height_mbs != 2 * gd.height_mbs) return H264R_ERR_INVALID;
    h264r_ctx* const sctx = to_field ? frame_ctx : field_ctx; h264r_ctx* const dctx = to_field ? field_ctx : frame_ctx;
    const h264r_frame sf = to_field ? frame : field, df = to_field ? field : frame;
    // The copy is enqueued now: pictures that were submitted but not flushed would run after it.  Neither picture may be
    // named by a queued picture (as destination of the source, or in any role of the destination).
    for (h264r_ctx* c : { sctx, dctx }) {
        std::lock_guard<std::mutex> lock(c->mu);
        for (int qi : c->queue) {
            const Slot& s = c->slots[qi];
            const h264r_frame f = c == sctx ? sf : df;
            if (s.dst == f) return H264R_ERR_STATE;
            if (c == dctx) for (int i = 0; i < s.pp.num_ref_frames; ++i) if (s.pp.ref_frames[i] == f) return H264R_ERR_STATE;
        }
    }
    cudaSetDevice(dctx->device);
    Frame& src = sctx->frames[sf]; Frame& dst = dctx->frames[df];
    cudaStream_t st = dctx->stream;                           // the destination's compute stream: behind every kernel that reads or writes it
    if (src.ready) { if (cudaStreamWaitEvent(st, src.ready, 0) != cudaSuccess) return cuda_fail(dctx, cudaGetLastError(), "h264r_field_copy"); }
    if (dst.pending_read) { if (cudaStreamWaitEvent(st, dst.read_done, 0) != cudaSuccess) return cuda_fail(dctx, cudaGetLastError(), "h264r_field_copy"); dst.pending_read = false; }
    const int w = gf.width_mbs * 16, hf = gd.height_mbs * 16;
    uint8_t* const fb = frame_ctx->frames[frame].dev; uint8_t* const db = field_ctx->frames[field].dev;
    struct Plane { size_t off_f, off_d; int pitch_f, pitch_d, width, rows; } planes[3] = {
        { 0, 0, gf.pitch_y, gd.pitch_y, w, hf }, { gf.off_cb, gd.off_cb, gf.pitch_c, gd.pitch_c, w / 2, hf / 2 }, { gf.off_cr, gd.off_cr, gf.pitch_c, gd.pitch_c, w / 2, hf / 2 } };
    for (const Plane& p : planes) {
        uint8_t* const fl = fb + p.off_f + (size_t)parity * p.pitch_f;     // the field's lines inside the frame: every second line
        uint8_t* const dl = db + p.off_d;
        const cudaError_t e = to_field ? cudaMemcpy2DAsync(dl, p.pitch_d, fl, 2 * (size_t)p.pitch_f, p.width, p.rows, cudaMemcpyDeviceToDevice, st)
                                       : cudaMemcpy2DAsync(fl, 2 * (size_t)p.pitch_f, dl, p.pitch_d, p.width, p.rows, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_fail(dctx, e, "h264r_field_copy");
    }
    // consumers of the destination wait for the copy; a later picture that overwrites the source waits for it too
    cudaEvent_t done = take_event(dctx);
    if (!done || cudaEventRecord(done, st) != cudaSuccess) return cuda_fail(dctx, cudaGetLastError(), "h264r_field_copy");
    dst.ready = done;                                          // (combining: the second field's event covers the first, same stream)
    if (!src.read_done && cudaEventCreateWithFlags(&src.read_done, cudaEventDisableTiming) != cudaSuccess) return cuda_fail(sctx, cudaGetLastError(), "h264r_field_copy");
    if (cudaEventRecord(src.read_done, st) != cudaSuccess) return cuda_fail(sctx, cudaGetLastError(), "h264r_field_copy");
    src.pending_read = true;
    return H264R_OK;
}

int h264r_get_stats(h264r_ctx* ctx, h264r_stats* out)
{
    if (!ctx || !out) return H264R_ERR_INVALID;
    *out = ctx->stats;
    return H264R_OK;
}

} // extern "C"
