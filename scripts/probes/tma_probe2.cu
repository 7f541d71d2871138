// Probe 2: the CUDA C++ Programming Guide's own TMA example (libcu++ wrappers), 2-D uint8/int tile load.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdio.h>
#include <stdint.h>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

constexpr int BW = 16, BH = 16;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int* out)
{
    __shared__ alignas(128) int smem_buffer[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) {
        init(&bar, blockDim.x);
        cde::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = smem_buffer[i / BW][i % BW];
}

int main()
{
    const int W = 64, H = 64;
    int* h = new int[W * H];
    for (int i = 0; i < W * H; ++i) h[i] = i;
    int *d, *dout;
    cudaMalloc(&d, W * H * 4); cudaMalloc(&dout, BW * BH * 4);
    cudaMemcpy(d, h, W * H * 4, cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap m{};
    const cuuint64_t dim[2] = { W, H }, stride[1] = { W * 4 };
    const cuuint32_t box[2] = { BW, BH }, one[2] = { 1, 1 };
    CUresult r = ((Enc)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dim, stride, box, one, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    kernel<<<1, 128>>>(m, 8, 4, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("guide example: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        int res[BW * BH]; cudaMemcpy(res, dout, sizeof(res), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r2 = 0; r2 < BH; ++r2) for (int c = 0; c < BW; ++c) if (res[r2 * BW + c] != h[(4 + r2) * W + 8 + c]) ++bad;
        printf("  mismatches %d\n", bad);
    }
    return 0;
}
