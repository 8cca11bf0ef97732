// Probe 2: the libcu++ way (CUDA programming guide, "Using TMA to transfer multi-dimensional arrays")
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <stdlib.h>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

template <typename T, int BW, int BH>
__global__ void k(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int *out)
{
    __shared__ alignas(128) T smem_buffer[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
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


template <typename T, int BW, int BH> int run(CUtensorMapDataType dt, const char *name)
{
    const int W = 256, H = 256;
    std::vector<T> h(W * H);
    for (int i = 0; i < W * H; i++) h[i] = (T)(i * 3 + 1);
    T *d; int *o; cudaMalloc(&d, W * H * sizeof(T)); cudaMalloc(&o, 4096 * 4);
    cudaMemcpy(d, h.data(), W * H * sizeof(T), cudaMemcpyHostToDevice);
    void *fn; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    auto enc = (CUresult(*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                            const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
    CUtensorMap tm{};
    cuuint64_t dims[2] = {W, H}, strides[1] = {W * sizeof(T)};
    cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
    CUresult r = enc(&tm, dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    k<T, BW, BH><<<1, 128>>>(tm, 33, 9, o);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<int> g(BW * BH);
    cudaMemcpy(g.data(), o, BW * BH * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < BH; yy++) for (int xx = 0; xx < BW; xx++) if ((T)g[yy * BW + xx] != h[(9 + yy) * W + 33 + xx]) bad++;
    printf("%s box %dx%d: encode %d, %s, mismatches %d\n", name, BW, BH, (int)r, cudaGetErrorString(e), bad);
    return e != cudaSuccess;
}
int main(int argc, char **argv)
{
    int v = argc > 1 ? atoi(argv[1]) : 0;
    if (v == 0) return run<int, 16, 16>(CU_TENSOR_MAP_DATA_TYPE_INT32, "int32");
    if (v == 1) return run<uint16_t, 16, 16>(CU_TENSOR_MAP_DATA_TYPE_UINT16, "uint16");
    if (v == 2) return run<uint16_t, 24, 23>(CU_TENSOR_MAP_DATA_TYPE_UINT16, "uint16");
    if (v == 3) return run<uint16_t, 64, 16>(CU_TENSOR_MAP_DATA_TYPE_UINT16, "uint16");
    return 0;
}
