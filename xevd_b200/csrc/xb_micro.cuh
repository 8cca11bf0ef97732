// xb_micro.cuh -- batched leaf kernels (BASELINE.json config 5): the same device functions as the picture
// kernels, driven over arrays of independent blocks.
//   xb200_itdq_blocks_dev : xevd_itdq / xevdm_itdq on n contiguous blocks (xevd_itdq.c:494-542)
//   xb200_mc_blocks_dev   : xevd_mc_l / xevd_mc_c on n blocks of one reference plane (xevd_mc.h:66-74)
#pragma once
#include "xb_common.cuh"
#include "xb_itdq.cuh"
#include "xb_recon.cuh"

namespace xb {

// One CTA per group of blocks; a thread owns one line.  Intermediate in shared memory with +1 padding.
template <bool IQT>
__global__ void __launch_bounds__(256) k_itdq_blocks(const int16_t *__restrict__ in, int16_t *__restrict__ out, int n, int lw, int lh,
                                                       int qp, int bd, int blocks_per_cta)
{
    extern __shared__ int s_tmp[];
    const int w = 1 << lw, h = 1 << lh, ts = w + 1;
    const int first = blockIdx.x * blocks_per_cta;
    const int nb = min(blocks_per_cta, n - first);
    Dequant dq;
    dq.init(lw, lh, qp, bd, IQT);
    for (int t = threadIdx.x; t < nb * w; t += blockDim.x) {
        const int b = t >> lw, x = t & (w - 1);
        const int16_t *src = in + ((size_t)(first + b) << (lw + lh)) + x;
        int *dst = s_tmp + b * h * ts + x;
        itx_line_dyn<IQT>(lh, [&](int k) { return dq.apply(src[k * w]); }, [&](int nn, int v) { dst[nn * ts] = v; }, IQT ? 7 : 0);
    }
    __syncthreads();
    const int sh2 = IQT ? 12 - (bd - 8) : 19 - (bd - 8);
    for (int t = threadIdx.x; t < nb * h; t += blockDim.x) {
        const int b = t >> lh, y = t & (h - 1);
        const int *srow = s_tmp + (b * h + y) * ts;
        int16_t *drow = out + ((size_t)(first + b) << (lw + lh)) + y * w;
        itx_line_dyn<false>(lw, [&](int k) { return srow[k]; }, [&](int nn, int v) { drow[nn] = (int16_t)v; }, sh2);
    }
}

inline int launch_itdq_blocks(const int16_t *in, int16_t *out, int n, int lw, int lh, int qp, int bd, int iqt, cudaStream_t st)
{
    const int w = 1 << lw, h = 1 << lh;
    int per = 256 / (w > h ? w : h);
    if (per < 1) per = 1;
    const size_t smem = (size_t)per * h * (w + 1) * sizeof(int);
    const int grid = (n + per - 1) / per;
    if (iqt) k_itdq_blocks<true><<<grid, 256, smem, st>>>(in, out, n, lw, lh, qp, bd, per);
    else     k_itdq_blocks<false><<<grid, 256, smem, st>>>(in, out, n, lw, lh, qp, bd, per);
    return 1;
}

// One warp per 16x16 (luma) / 8x8.. tile of a block; mv = {gmv_x, gmv_y, ori_mv_x, ori_mv_y} per block.
__global__ void __launch_bounds__(256) k_mc_blocks(const pel *__restrict__ ref, int stride, int chroma, const int *__restrict__ mv,
                                                     pel *__restrict__ out, int n, int w, int h, int bd, int main_tables)
{
    __shared__ int16_t s_scr[8][kMcScratchPerWarp];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tw = min(w, 16), th = min(h, 16);
    const int tiles_x = w / tw, tiles_y = h / th, tiles = tiles_x * tiles_y;
    const int rpl = max(1, (tw * th) >> 5);
    for (long long t = (long long)blockIdx.x * 8 + warp; t < (long long)n * tiles; t += (long long)gridDim.x * 8) {
        const int b = (int)(t / tiles), ti = (int)(t % tiles);
        const int tx = (ti % tiles_x) * tw, ty = (ti / tiles_x) * th;
        const int gx = mv[b * 4 + 0], gy = mv[b * 4 + 1], ox = mv[b * 4 + 2], oy = mv[b * 4 + 3];
        int pr[8];
        if (!chroma) {
            const bool fx = ((ox | (ox >> 1) | (ox >> 2) | (ox >> 3)) & 1) != 0, fy = ((oy | (oy >> 1) | (oy >> 2) | (oy >> 3)) & 1) != 0;
            mc_tile<8>(ref + ((gy >> 4) + ty) * stride + (gx >> 4) + tx, stride, c_mc_l[main_tables][gx & 15], c_mc_l[main_tables][gy & 15],
                       fx, fy, tw, th, rpl, bd, s_scr[warp], lane, pr);
        } else {
            const bool fx = ((ox | (ox >> 1) | (ox >> 2) | (ox >> 3) | (ox >> 4)) & 1) != 0,
                       fy = ((oy | (oy >> 1) | (oy >> 2) | (oy >> 3) | (oy >> 4)) & 1) != 0;
            mc_tile<4>(ref + ((gy >> 5) + ty) * stride + (gx >> 5) + tx, stride, c_mc_c[main_tables][gx & 31], c_mc_c[main_tables][gy & 31],
                       fx, fy, tw, th, rpl, bd, s_scr[warp], lane, pr);
        }
        const int col = lane & (tw - 1), r0 = (lane / tw) * rpl;
        if (r0 < th) {
            pel *dst = out + (size_t)b * w * h + (ty + r0) * w + tx + col;
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (i < rpl) dst[i * w] = (pel)pr[i];
        }
    }
}

inline int launch_mc_blocks(const pel *ref, int stride, int chroma, const int *mv, pel *out, int n, int w, int h, int bd, int main_tables,
                            cudaStream_t st)
{
    if (w < 2 || h < 2 || w > 128 || h > 128 || (w & (w - 1)) || (h & (h - 1))) return XB200_ERR_INVALID_ARGUMENT;
    const int tw = w < 16 ? w : 16, th = h < 16 ? h : 16;
    const long long tiles = (long long)n * (w / tw) * (h / th);
    long long grid = (tiles + 7) / 8;
    if (grid > 148 * 32) grid = 148 * 32;
    k_mc_blocks<<<(int)grid, 256, 0, st>>>(ref, stride, chroma, mv, out, n, w, h, bd, main_tables);
    return 1;
}

}  // namespace xb
