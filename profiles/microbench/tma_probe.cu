// Probe: 2-D TMA box loads of uint16 windows with the tensor map (A) passed as __grid_constant__ and (B) read from
// global memory, issued by several lanes of one warp with different coordinates.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <stdlib.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void k(const __grid_constant__ CUtensorMap pm, const CUtensorMap *gm, uint16_t *out, int nl, int BW, int BH)
{
    __shared__ __align__(128) uint16_t win[8][2048];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(nl * BW * BH * 2) : "memory");
    __syncwarp();
    if (threadIdx.x < nl) {
        const void *tm = MODE == 0 ? (const void *)&pm : (const void *)gm;
        int c0 = 3 + 5 * threadIdx.x, c1 = 2 + 3 * threadIdx.x;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(&win[threadIdx.x][0])), "l"(tm), "r"(c0), "r"(c1), "r"(s32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@!p bra W;\n\t}" ::"r"(s32(&bar)) : "memory");
    for (int i = threadIdx.x; i < nl * BW * BH; i += blockDim.x) out[i] = win[i / (BW * BH)][i % (BW * BH)];
}

int main(int argc, char **argv)
{
    const int BW = argc > 1 ? atoi(argv[1]) : 24, BH = argc > 2 ? atoi(argv[2]) : 23; const int only_mode = argc > 3 ? atoi(argv[3]) : -1;
    const int W = 512, H = 256;
    std::vector<uint16_t> h(W * H);
    for (int i = 0; i < W * H; i++) h[i] = (uint16_t)(i * 7 + (i >> 9));
    uint16_t *d, *o; cudaMalloc(&d, W * H * 2); cudaMalloc(&o, 8 * 2048 * 2);
    cudaMemcpy(d, h.data(), W * H * 2, cudaMemcpyHostToDevice);
    void *fn; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    auto enc = (CUresult(*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                            const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
    alignas(64) CUtensorMap tm;
    cuuint64_t dims[2] = {W, H}, strides[1] = {W * 2};
    cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)BH}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    CUtensorMap *gtm; cudaMalloc(&gtm, sizeof(tm)); cudaMemcpy(gtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; mode++)
        for (int nl = 1; nl <= 8; nl += 7) {
            cudaMemset(o, 0, 8 * 2048 * 2); if (only_mode >= 0 && mode != only_mode) continue;
            if (mode == 0) k<0><<<1, 64>>>(tm, gtm, o, nl, BW, BH); else k<1><<<1, 64>>>(tm, gtm, o, nl, BW, BH);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<uint16_t> g(8 * 2048);
            cudaMemcpy(g.data(), o, g.size() * 2, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int l = 0; l < nl; l++)
                for (int y = 0; y < BH; y++)
                    for (int x = 0; x < BW; x++)
                        if (g[l * BW * BH + y * BW + x] != h[(2 + 3 * l + y) * W + 3 + 5 * l + x]) bad++;
            printf("box %dx%d mode %s, %d lanes: %s, mismatches %d\n", BW, BH, mode ? "global-memory tensormap" : "param tensormap", nl, cudaGetErrorString(e), bad);
            if (e != cudaSuccess) return 1;
        }
    return 0;
}
