// Measurement support declared in include/h264recon_bench.h (not part of the drop-in boundary): a multi-threaded feeder
// that drives the PUBLIC entry points the way a set of parser threads would, and the box's plain copy ceiling.
#include "h264recon_bench.h"

#include <cuda_runtime.h>

#include <emmintrin.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string.h>
#include <thread>
#include <vector>

namespace {

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// What a parser thread does to the staging memory is WRITE it, once, front to back.  The stand-in copies a pre-parsed
// picture there with streaming stores: the lines are written whole, so reading them into the cache first (the
// read-for-ownership of an ordinary store) would only add host-memory traffic next to the two DMA directions.
void stream_copy(void* dst, const void* src, size_t bytes)
{
    uint8_t* d = static_cast<uint8_t*>(dst);
    const uint8_t* s = static_cast<const uint8_t*>(src);
    while (bytes && (reinterpret_cast<uintptr_t>(d) & 15)) { *d++ = *s++; --bytes; }
    size_t n = bytes / 64;
    for (; n; --n, d += 64, s += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s)), b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 16)),
                      c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 32)), e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(d), a); _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), c); _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), e);
    }
    bytes &= 63;
    memcpy(d, s, bytes);
}

struct FeedShared {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int> submitted;          // picture indexes submitted since the last flush
    int total_submitted = 0;
    bool starving = false;               // a feeder found every staging slot filled or queued: flush now
    int error = 0;
};

} // namespace

extern "C" {

double h264r_bench_feed(h264r_ctx* ctx, const h264r_bench_picture* pics, int num_pics, int num_mbs, int threads, int flush_every,
                        int steps, double* host_fill_s, double* host_flush_s)
{
    if (!ctx || !pics || num_pics <= 0 || num_mbs <= 0 || threads <= 0 || flush_every <= 0 || steps <= 0) return (double)H264R_ERR_INVALID;
    FeedShared sh;
    std::vector<double> fill_s((size_t)threads, 0.0);
    const int total = num_pics * steps;
    const size_t head_mbs = sizeof(h264r_mb) * (size_t)num_mbs;

    // Feeder t owns the streams s with s % threads == t and feeds their pictures in list order, step after step: the
    // pictures of one stream are submitted in decode order (a picture follows its references), streams are independent.
    auto feeder = [&](int t) {
        for (int step = 0; step < steps; ++step)
            for (int i = 0; i < num_pics; ++i) {
                const h264r_bench_picture& p = pics[i];
                if (p.stream_id % threads != t) continue;
                const double t0 = now_s();
                h264r_pic_buffers bufs;
                int rc;
                for (;;) {
                    rc = h264r_picture_begin(ctx, p.dst, &p.pp, &bufs);
                    if (rc != H264R_ERR_NOMEM) break;
                    { std::lock_guard<std::mutex> lock(sh.mu); sh.starving = true; if (sh.error) return; }
                    sh.cv.notify_all();
                    std::this_thread::sleep_for(std::chrono::microseconds(50));
                }
                if (rc == H264R_OK) {
                    stream_copy(bufs.mbs, p.head, head_mbs + sizeof(h264r_slice) * (size_t)p.pp.num_slices);
                    stream_copy(bufs.stream, p.stream, sizeof(uint32_t) * (size_t)p.stream_words);
                    _mm_sfence();                                // the streaming stores are globally visible before the submit
                    rc = h264r_picture_submit(ctx, bufs.picture, p.stream_words);
                }
                fill_s[(size_t)t] += now_s() - t0;
                {
                    std::lock_guard<std::mutex> lock(sh.mu);
                    if (rc != H264R_OK) { if (!sh.error) sh.error = rc; }
                    else { sh.submitted.push_back(i); sh.total_submitted += 1; }
                    if (sh.error) { sh.cv.notify_all(); return; }
                }
                sh.cv.notify_all();
            }
    };

    const double t_start = now_s();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(feeder, t);

    // the calling thread owns the GPU side: flush + downloads
    double flush_s = 0.0;
    int flushed = 0, rc_main = H264R_OK;
    std::vector<int> batch;
    while (flushed < total && rc_main == H264R_OK) {
        {
            std::unique_lock<std::mutex> lock(sh.mu);
            sh.cv.wait(lock, [&] { return sh.error || sh.starving || (int)sh.submitted.size() >= flush_every || sh.total_submitted == total; });
            if (sh.error) { rc_main = sh.error; break; }
            batch.swap(sh.submitted);
            sh.starving = false;
        }
        if (batch.empty()) continue;
        const double t0 = now_s();
        rc_main = h264r_flush(ctx);
        for (size_t k = 0; k < batch.size() && rc_main == H264R_OK; ++k) {
            const h264r_bench_picture& p = pics[batch[k]];
            if (!p.out) continue;
            const size_t ny = (size_t)num_mbs * 256;
            rc_main = h264r_frame_download_async(ctx, p.dst, p.out, p.out + ny, p.out + ny + ny / 4, p.pitch_y, p.pitch_y / 2);
        }
        flush_s += now_s() - t0;
        flushed += (int)batch.size();
        batch.clear();
    }
    if (rc_main != H264R_OK) { std::lock_guard<std::mutex> lock(sh.mu); if (!sh.error) sh.error = rc_main; }
    sh.cv.notify_all();
    for (std::thread& th : pool) th.join();
    const int rc_wait = h264r_wait(ctx, -1);
    const double elapsed = now_s() - t_start;
    if (host_fill_s) { double s = 0.0; for (double v : fill_s) s += v; *host_fill_s = s; }
    if (host_flush_s) *host_flush_s = flush_s;
    const int rc = sh.error ? sh.error : rc_wait;
    return rc == H264R_OK ? elapsed : (double)rc;
}

int h264r_bench_copy_ceiling(int device, size_t h2d_bytes, size_t d2h_bytes, size_t chunk, int iterations, double gbs_out[3])
{
    if (!gbs_out || chunk == 0 || iterations <= 0) return H264R_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return H264R_ERR_NODEVICE;
    const int ring = 32;
    uint8_t *h_up = nullptr, *h_down = nullptr, *d_up = nullptr, *d_down = nullptr;
    cudaStream_t s_up = nullptr, s_down = nullptr;
    cudaEvent_t e[4] = { nullptr, nullptr, nullptr, nullptr };
    cudaError_t err = cudaHostAlloc((void**)&h_up, chunk * ring, cudaHostAllocDefault);
    if (err == cudaSuccess) err = cudaHostAlloc((void**)&h_down, chunk * ring, cudaHostAllocDefault);
    if (err == cudaSuccess) err = cudaMalloc((void**)&d_up, chunk * ring);
    if (err == cudaSuccess) err = cudaMalloc((void**)&d_down, chunk * ring);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&s_up, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&s_down, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && err == cudaSuccess; ++i) err = cudaEventCreate(&e[i]);
    if (err == cudaSuccess) { memset(h_up, 1, chunk * ring); err = cudaMemset(d_down, 2, chunk * ring); }
    double wall = 0.0;
    float ms_up = 0.f, ms_down = 0.f;
    if (err == cudaSuccess) {
        const size_t n_up = h2d_bytes / chunk, n_down = d2h_bytes / chunk;
        for (int pass = 0; pass < 2 && err == cudaSuccess; ++pass) {            // pass 0 warms up
            const int iters = pass ? iterations : 1;
            cudaDeviceSynchronize();
            const double t0 = now_s();
            cudaEventRecord(e[0], s_up); cudaEventRecord(e[2], s_down);
            for (int it = 0; it < iters; ++it) {
                // interleave the enqueues so that both directions are fed from the start
                for (size_t k = 0; k < (n_up > n_down ? n_up : n_down); ++k) {
                    if (k < n_up) cudaMemcpyAsync(d_up + (k % ring) * chunk, h_up + (k % ring) * chunk, chunk, cudaMemcpyHostToDevice, s_up);
                    if (k < n_down) cudaMemcpyAsync(h_down + (k % ring) * chunk, d_down + (k % ring) * chunk, chunk, cudaMemcpyDeviceToHost, s_down);
                }
            }
            cudaEventRecord(e[1], s_up); cudaEventRecord(e[3], s_down);
            err = cudaStreamSynchronize(s_up);
            if (err == cudaSuccess) err = cudaStreamSynchronize(s_down);
            wall = (now_s() - t0) / iters;
            if (err == cudaSuccess) { cudaEventElapsedTime(&ms_up, e[0], e[1]); cudaEventElapsedTime(&ms_down, e[2], e[3]); }
            if (pass) {
                gbs_out[0] = n_up && ms_up > 0 ? (double)n_up * chunk * iters / (ms_up * 1e-3) / 1e9 : 0.0;
                gbs_out[1] = n_down && ms_down > 0 ? (double)n_down * chunk * iters / (ms_down * 1e-3) / 1e9 : 0.0;
                gbs_out[2] = wall;
            }
        }
    }
    for (int i = 0; i < 4; ++i) if (e[i]) cudaEventDestroy(e[i]);
    if (s_up) cudaStreamDestroy(s_up);
    if (s_down) cudaStreamDestroy(s_down);
    if (h_up) cudaFreeHost(h_up);
    if (h_down) cudaFreeHost(h_down);
    if (d_up) cudaFree(d_up);
    if (d_down) cudaFree(d_down);
    return err == cudaSuccess ? H264R_OK : H264R_ERR_CUDA;
}

} // extern "C"
