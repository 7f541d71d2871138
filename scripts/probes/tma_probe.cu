// Probe: one 16 x 13 uint8 TMA box (cp.async.bulk.tensor.2d) per lane-group, descriptor (a) in global memory written by
// cudaMemcpy, (b) passed as a __grid_constant__ kernel parameter.  Build: nvcc -arch=sm_100a -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const CUtensorMap* gmap, const __grid_constant__ CUtensorMap pmap, int use_param, int x, int y, uint8_t* out, int issuers, int stage, int box_bytes)
{
    __shared__ __align__(128) uint8_t win[4][2048];
    __shared__ __align__(8) uint64_t bar;
    const int lane = threadIdx.x;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(saddr(&bar)), "r"(32) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (issuers >= 10) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    const CUtensorMap* map = use_param ? &pmap : gmap;
    if (stage == 2 && lane == 0) {               // plain (non-tensor) bulk copy of 208 bytes
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(saddr(&bar)), "r"(208) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(saddr(win[0])), "l"(out + 8192), "r"(208), "r"(saddr(&bar)) : "memory");
    } else if (stage == 1 && (lane & 7) == 0 && (lane >> 3) < issuers % 10) {   // up to four lanes issue a box each (divergent coordinates)
        const int q = lane >> 3;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(saddr(&bar)), "r"(box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(saddr(win[q])), "l"(map), "r"(x + q), "r"(y + 2 * q), "r"(saddr(&bar)) : "memory");
    } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(saddr(&bar)) : "memory");
    }
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(saddr(&bar)), "r"(0) : "memory");
    } while (!ok);
    __syncwarp();
    for (int i = lane; i < 4 * 2048; i += 32) out[i] = win[i / 2048][i % 2048];
}

int main(int argc, char** argv)
{
    const int issuers = argc > 1 ? atoi(argv[1]) : 4, first_param = argc > 2 ? atoi(argv[2]) : 1, stage = argc > 3 ? atoi(argv[3]) : 1;
    const int box_w = argc > 4 ? atoi(argv[4]) : 16, box_h = argc > 5 ? atoi(argv[5]) : 13, byver = argc > 6 ? atoi(argv[6]) : 0;
    const int W = 64, H = 48;
    uint8_t* h = new uint8_t[W * H];
    for (int i = 0; i < W * H; ++i) h[i] = (uint8_t)((i * 7 + (i / W) * 13) & 255);
    uint8_t *d, *dout; CUtensorMap* dmap;
    cudaMalloc(&d, W * H); cudaMalloc(&dout, 16384); cudaMalloc(&dmap, sizeof(CUtensorMap));
    cudaMemcpy(d, h, W * H, cudaMemcpyHostToDevice);
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaError_t e = byver ? cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &qr)
                          : cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    printf("entry point: %s qr=%d fn=%p\n", cudaGetErrorString(e), (int)qr, fn);
    alignas(64) CUtensorMap m;
    const cuuint64_t dim[2] = { W, H }, stride[1] = { W };
    const cuuint32_t box[2] = { (cuuint32_t)box_w, (cuuint32_t)box_h }, one[2] = { 1, 1 };
    CUresult r = ((Enc)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dim, stride, box, one, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    cudaMemcpy(dmap, &m, sizeof(m), cudaMemcpyHostToDevice);
    for (int it = 0; it < 2; ++it) {
        const int use_param = it == 0 ? first_param : !first_param;
        cudaMemset(dout, 0, 8192);
        probe<<<1, 32>>>(dmap, m, use_param, 5, 3, dout, issuers, stage, box_w * box_h);
        e = cudaDeviceSynchronize();
        printf("%s descriptor: %s\n", use_param ? "param" : "global", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        if (stage != 1) continue;
        static uint8_t res[8192];
        cudaMemcpy(res, dout, sizeof(res), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int q = 0; q < issuers % 10; ++q) for (int r2 = 0; r2 < box_h; ++r2) for (int c = 0; c < box_w; ++c)
            if (res[q * 2048 + r2 * box_w + c] != h[(3 + 2 * q + r2) * W + 5 + q + c]) ++bad;
        printf("  mismatches: %d\n", bad);
    }
    return 0;
}
