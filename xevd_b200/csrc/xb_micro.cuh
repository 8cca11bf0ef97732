// xb_micro.cuh -- batched leaf kernels (BASELINE.json config 5): the same device functions as the picture
// kernels, driven over arrays of independent blocks.
//   xb200_itdq_blocks_dev : xevd_itdq / xevdm_itdq on n contiguous blocks (xevd_itdq.c:494-542)
//   xb200_mc_blocks_dev   : xevd_mc_l / xevd_mc_c on n blocks of one reference plane (xevd_mc.h:66-74)
#pragma once
#include "xb_common.cuh"
#include "xb_itdq.cuh"
#include "xb_recon.cuh"
#include "xb_recon2.cuh"

namespace xb {

// The transform passes are the ones of the picture kernel (row_pass / row_pass2 / col_pass / col_pass2, xb_recon2.cuh: IDP.2A first pass on
// packed s16 pairs, two neighbouring lines per thread for lines of at most 16 points, packed stores), so this measures the arithmetic
// the product runs.  A CTA handles 4096 samples' worth of blocks: coefficients staged by coalesced 16-byte loads, pass-1 results in shared
// memory (row stride w + 4 words: 16-byte row stores of 8 consecutive rows hit 8 distinct bank quads), residual staged and written back
// by coalesced 16-byte stores.
template <bool IQT>
__global__ void __launch_bounds__(256) k_itdq_blocks(const int16_t *__restrict__ in, int16_t *__restrict__ out, int n, int lw, int lh,
                                                       int mul, int shift, int wide, int bd, int per)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int w = 1 << lw, h = 1 << lh, bs = w * h, ts = w + 4;
    int16_t *s_coef = (int16_t *)s_raw;
    int16_t *s_res = s_coef + per * bs;
    int *s_tmp = (int *)(s_res + per * bs);
    const long long first = (long long)blockIdx.x * per;
    const int nb = (int)min((long long)per, n - first), tot = nb * bs;
    const int16_t *g_in = in + first * bs;
    int16_t *g_out = out + first * bs;
    for (int i = threadIdx.x; i < tot >> 3; i += blockDim.x) ((int4 *)s_coef)[i] = __ldg((const int4 *)g_in + i);
    for (int i = (tot & ~7) + threadIdx.x; i < tot; i += blockDim.x) s_coef[i] = g_in[i];
    __syncthreads();
    const int off = shift ? 1 << (shift - 1) : 0;
    {
        const int nr = IQT ? w : row_tasks(lw, lh, false), lnr = 31 - __clz(nr);       // IQT: the first pass runs along columns
        for (int t = threadIdx.x; t < nb * nr; t += blockDim.x) {
            const int b = t >> lnr, q = t & (nr - 1);
            if (IQT) {
                const int16_t *src = s_coef + b * bs + q;
                int *dst = s_tmp + b * h * ts + q;
                switch (lh) {
                case 1: row_pass<2, true>(src, w, dst, ts, mul, off, shift, wide); break;
                case 2: row_pass<4, true>(src, w, dst, ts, mul, off, shift, wide); break;
                case 3: row_pass<8, true>(src, w, dst, ts, mul, off, shift, wide); break;
                case 4: row_pass<16, true>(src, w, dst, ts, mul, off, shift, wide); break;
                case 5: row_pass<32, true>(src, w, dst, ts, mul, off, shift, wide); break;
                default: row_pass<64, true>(src, w, dst, ts, mul, off, shift, wide); break;
                }
            } else {
                const int r0 = lw <= 4 ? 2 * q : q;
                const int16_t *src = s_coef + b * bs + r0 * w;
                int *dst = s_tmp + (b * h + r0) * ts;
                switch (lw) {
                case 1: row_pass2<2>(src, w, dst, ts, mul, off, shift, wide); break;
                case 2: row_pass2<4>(src, w, dst, ts, mul, off, shift, wide); break;
                case 3: row_pass2<8>(src, w, dst, ts, mul, off, shift, wide); break;
                case 4: row_pass2<16>(src, w, dst, ts, mul, off, shift, wide); break;
                case 5: row_pass<32, false>(src, w, dst, ts, mul, off, shift, wide); break;
                default: row_pass<64, false>(src, w, dst, ts, mul, off, shift, wide); break;
                }
            }
        }
    }
    __syncthreads();
    {
        const int sh2 = (IQT ? 12 : 19) - (bd - 8);
        const int nc = IQT ? h : col_tasks(lw, lh, false), lnc = 31 - __clz(nc);        // IQT: the second pass runs along rows
        for (int t = threadIdx.x; t < nb * nc; t += blockDim.x) {
            const int b = t >> lnc, q = t & (nc - 1);
            if (IQT) {
                const int *src = s_tmp + (b * h + q) * ts;
                int16_t *dst = s_res + b * bs + q * w;
                switch (lw) {
                case 1: col_pass<2, true>(src, ts, dst, w, sh2); break;
                case 2: col_pass<4, true>(src, ts, dst, w, sh2); break;
                case 3: col_pass<8, true>(src, ts, dst, w, sh2); break;
                case 4: col_pass<16, true>(src, ts, dst, w, sh2); break;
                case 5: col_pass<32, true>(src, ts, dst, w, sh2); break;
                default: col_pass<64, true>(src, ts, dst, w, sh2); break;
                }
            } else {
                const int c0 = lh <= 4 ? 2 * q : q;
                const int *src = s_tmp + b * h * ts + c0;
                int16_t *dst = s_res + b * bs + c0;
                switch (lh) {
                case 1: col_pass2<2>(src, ts, dst, w, sh2); break;
                case 2: col_pass2<4>(src, ts, dst, w, sh2); break;
                case 3: col_pass2<8>(src, ts, dst, w, sh2); break;
                case 4: col_pass2<16>(src, ts, dst, w, sh2); break;
                case 5: col_pass<32, false>(src, ts, dst, w, sh2); break;
                default: col_pass<64, false>(src, ts, dst, w, sh2); break;
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < tot >> 3; i += blockDim.x) ((int4 *)g_out)[i] = ((const int4 *)s_res)[i];
    for (int i = (tot & ~7) + threadIdx.x; i < tot; i += blockDim.x) g_out[i] = s_res[i];
}

inline size_t itdq_blocks_smem(int lw, int lh, int per) { return (size_t)per * (1 << (lw + lh)) * 4 + (size_t)per * (1 << lh) * ((1 << lw) + 4) * 4; }
inline int launch_itdq_blocks(const int16_t *in, int16_t *out, int n, int lw, int lh, int qp, int bd, int iqt, cudaStream_t st)
{
    static const int dq[2][6] = {{40, 45, 51, 57, 64, 71}, {40, 45, 51, 57, 64, 72}};       // xevd_tbl_dq_scale_b / xevd_tbl_dq_scale (xevd_tbl.c:255-256)
    // xevd_itdq prologue (xevd_itdq.c:494-517)
    const int odd = (lw + lh) & 1;
    const int shift = 20 - 14 - (15 - bd - ((lw + lh) >> 1)) + (odd ? 8 : 0);
    const long long mul = (long long)(dq[iqt ? 1 : 0][qp % 6] << (qp / 6)) * (odd ? 181 : 1);
    if (shift < 0 || mul > 0x7fffffffLL) return XB200_ERR_INVALID_ARGUMENT;
    int per = 4096 >> (lw + lh);
    if (per < 1) per = 1;
    const size_t smem = itdq_blocks_smem(lw, lh, per);
    const int grid = (n + per - 1) / per;
    if (iqt) k_itdq_blocks<true><<<grid, 256, smem, st>>>(in, out, n, lw, lh, (int)mul, shift, mul >= 65536, bd, per);
    else     k_itdq_blocks<false><<<grid, 256, smem, st>>>(in, out, n, lw, lh, (int)mul, shift, mul >= 65536, bd, per);
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
