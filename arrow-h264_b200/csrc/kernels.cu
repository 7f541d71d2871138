// sm_100a kernels of the H.264 macroblock-reconstruction path, per wave of pictures:
//
//   residual_kernel           : kernel_residual.cuh -- dequantisation, DC Hadamards, inverse transforms -> int16 residual plane
//   deblock_prep_kernel       : kernel_residual.cuh -- boundary strengths + thresholds -> 64-byte descriptor per MB
//   recon_inter2_kernel       : kernel_inter.cuh    -- inter MBs: MC + weighted prediction + residual add
//   recon_intra_kernel        : kernel_intra.cuh    -- all-intra pictures, row wavefront
//   intra_list_kernel         : kernel_intra.cuh    -- addresses of the intra MBs of P/B pictures (the host sends none)
//   recon_intra_sparse_kernel : kernel_intra.cuh    -- intra MBs of P/B pictures
//   deblock_kernel            : kernel_deblock.cuh  -- deblocking filter, row wavefront
//
// Arithmetic follows the reference (src/codec/h264/decoder/{transform,inter_prediction,intra_prediction,
// deblock}.cc); the line-by-line citations live in the CPU restatement oracle/port_recon.c, whose structure
// these kernels mirror.  All sample arithmetic is int32; results are bit-exact by construction.
#include <algorithm>

#include "kernel_residual.cuh"
#include "kernel_inter.cuh"
#include "kernel_intra.cuh"
#include "kernel_deblock.cuh"
// -DH264R_DEBLOCK_PACKED=1: the deblocking of a wave runs in deblock4_kernel (two sample lines per lane in fp16x2, four pictures
// per warp) instead of deblock_kernel.  Bit-exact, measured slower in round 2; not compiled by default.
#ifndef H264R_DEBLOCK_PACKED
#define H264R_DEBLOCK_PACKED 0
#endif
#if H264R_DEBLOCK_PACKED
#include "kernel_deblock4.cuh"
#endif

namespace h264r {

const char* wave_kernel_name(int which)
{
    switch (which) {
    case KERNEL_RESID:   return "residual_kernel";
    case KERNEL_DBPREP:  return "deblock_prep_kernel";
    case KERNEL_LIST:    return "intra_list_kernel";
    case KERNEL_INTER:   return "recon_inter2_kernel";
    case KERNEL_INTRA:   return "recon_intra_kernel + recon_intra_sparse_kernel";
    case KERNEL_DEBLOCK: return H264R_DEBLOCK_PACKED ? "deblock4_kernel" : "deblock_kernel";
    default:             return nullptr;
    }
}

int launch_wave_kernel(const WaveLaunch& w, int which, cudaStream_t stream)
{
    const int nmb = w.geom.width_mbs * w.geom.height_mbs;
    const int threads = kWarpsPerCta * 32;
    const int groups = (w.geom.height_mbs + kWarpsPerCta - 1) / kWarpsPerCta;
    if (which == KERNEL_RESID) {
        residual_kernel<<<dim3((nmb + kResidWarps - 1) / kResidWarps, 1, w.num_pics), kResidWarps * 32, 0, stream>>>(w.pics, w.geom, w.err);
        return 1;
    }
    if (which == KERNEL_DBPREP) {
        if (!w.any_deblock) return 0;
        const long long total = (long long)w.num_pics * nmb;
        deblock_prep_kernel<<<(int)((total + 127) / 128), 128, 0, stream>>>(w.pics, w.num_pics, w.geom);
        return 1;
    }
    if (which == KERNEL_LIST) {
        if (!w.any_inter) return 0;
        intra_list_kernel<<<w.num_pics, kListThreads, 0, stream>>>(w.pics, w.geom, w.wave_max);
        return 1;
    }
    if (which == KERNEL_INTER) {
        if (!w.any_inter) return 0;
        const dim3 grid((w.geom.width_mbs + 2 * kInter2Warps - 1) / (2 * kInter2Warps), w.geom.height_mbs, w.num_pics);
        if (w.any_field) recon_inter2_kernel<true><<<grid, kInter2Warps * 32, 0, stream>>>(w.pics, w.geom);
        else             recon_inter2_kernel<false><<<grid, kInter2Warps * 32, 0, stream>>>(w.pics, w.geom);
        return 1;
    }
    if (which == KERNEL_INTRA) {
        int n = 0;
        if (w.any_intra_rows) {
            recon_intra_kernel<<<w.num_pics * groups, threads, 0, stream>>>(w.pics, w.num_pics, w.tickets, w.geom, w.epoch);
            ++n;
        }
        if (w.any_inter) {                                   // pictures with P / B slices may hold intra MBs anywhere
            // persistent CTAs (the number of intra MBs is only known on the device): one round of resident CTAs, fewer for tiny waves
            const long long most = ((long long)w.num_pics * nmb + kSparseWarps - 1) / kSparseWarps;    // one warp per MB at the very most
            const int ctas = (int)std::min<long long>(most, 148LL * H264R_SPARSE_CTAS);
            recon_intra_sparse_kernel<<<ctas, kSparseWarps * 32, 0, stream>>>(w.pics, w.num_pics, w.tickets, w.wave_max, w.geom, w.epoch);
            ++n;
        }
        return n;
    }
    if (which == KERNEL_DEBLOCK) {
        if (!w.any_deblock) return 0;
#if H264R_DEBLOCK_PACKED
        deblock4_kernel<<<((w.num_pics + 3) / 4) * groups, threads, 0, stream>>>(w.pics, w.num_pics, w.tickets, w.geom, w.epoch);
#else
        deblock_kernel<<<((w.num_pics + 1) / 2) * groups, threads, 0, stream>>>(w.pics, w.num_pics, w.tickets, w.geom, w.epoch);
#endif
        return 1;
    }
    return 0;
}

} // namespace h264r
