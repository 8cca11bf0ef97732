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

// xevd_mc_l / xevd_mc_c of n blocks (BASELINE config 5) with the prediction stages of k_recon_inter_v2's warp-slot variant: a warp takes
// 16x16 luma (8x8 chroma) tiles of the blocks one after the other - TMA box of the 8-sample-aligned window superset into the warp's own
// slot, horizontal stage with IDP.2A on packed sample pairs into vertical pairs, vertical stage, clipped result straight to the output
// block; the next tile's box is issued as soon as the horizontal stage is done.  mv = {gmv_x, gmv_y (1/16 luma, 1/32 chroma pel, absolute,
// clipped), ori_mv_x, ori_mv_y (the unclipped vector picks the variant 00 / n0 / 0n / nn, T3)} per block.
// (Round 1 ran the generic kernel's tile routine here - per-sample clamped loads, one IMAD per tap: 37-39 us per 4 Mi luma samples.)
struct McBlocksArgs {
    const CUtensorMap *tm;          // the 2-D map of the plane
    const int *mv;
    pel *out;
    int n, w, h, bd, main_tables, pad;      // pad: samples between the buffer origin and sample (0, 0)
};
// CHROMA: a warp takes two 8x8 tiles at a time, one per half-warp (the fused two-stage pass needs 16 lanes per tile)
template <bool CHROMA>
__global__ void __launch_bounds__(256) k_mc_blocks(const __grid_constant__ McBlocksArgs a)
{
    constexpr int T = CHROMA ? 8 : 16;                               // tile edge
    constexpr int kWin = CHROMA ? 2 * 640 : kWinLBytes;             // chroma: two windows of 24 x 11 samples (528 bytes, 128-byte aligned)
    __shared__ __align__(128) unsigned char s_win[8][kWin];
    __shared__ __align__(16) int s_m2[8][CHROMA ? 4 : kM2LWords];
    __shared__ __align__(8) uint64_t s_bar[8];
    __shared__ int s_taps[16 * 9 + 32 * 6];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 16 * 9) s_taps[tid] = __ldg(&c_taps5[a.main_tables][tid]);
    if (tid < 32 * 6) s_taps[16 * 9 + tid] = __ldg(&c_taps3[a.main_tables][tid]);
    if (tid < 8) mbar_init(&s_bar[tid], 1);
    __syncthreads();
    const int *s_t8 = s_taps, *s_t4 = s_taps + 16 * 9;
    auto ld_taps5 = [&](int ph, int set) { Taps5 t; const int *q = s_t8 + (ph & 15) * 9 + set * 3; t.r0 = q[0]; t.r1 = q[1]; t.r2 = q[2]; return t; };
    auto ld_taps3 = [&](int ph, int set) { Taps3 t; const int *q = s_t4 + (ph & 31) * 6 + set * 2; t.r0 = q[0]; t.r1 = q[1]; return t; };
    const int tw = min(a.w, T), th = min(a.h, T);
    const int tiles_x = a.w / tw, tiles = tiles_x * (a.h / th);
    constexpr int PER = CHROMA ? 2 : 1;                              // tiles per warp and step
    const long long total = (long long)a.n * tiles, step = (long long)gridDim.x * 8 * PER;
    uint64_t *bar = &s_bar[warp];
    const int fsh = CHROMA ? 5 : 4, fmask = (1 << fsh) - 1, lead = CHROMA ? 1 : 3;
    auto origin = [&](long long t, int &wx, int &wy) {              // window origin of tile t in padded-plane coordinates
        const int b = (int)(t / tiles), ti = (int)(t % tiles);
        wx = a.pad + (a.mv[b * 4 + 0] >> fsh) + (ti % tiles_x) * tw - lead;
        wy = a.pad + (a.mv[b * 4 + 1] >> fsh) + (ti / tiles_x) * th - lead;
    };
    auto issue = [&](long long t) {                                  // t: the warp's first tile of the step
        if (lane == 0) {
            int wx, wy;
            if (CHROMA) {
                const int n_t = t + 1 < total ? 2 : 1;
                mbar_expect_tx(bar, (uint32_t)n_t * 2u * kBoxCW * kBoxCH);
                for (int k = 0; k < n_t; k++) { origin(t + k, wx, wy); tma_load_2d(s_win[warp] + k * 640, a.tm, wx & ~7, wy, bar); }
            } else {
                origin(t, wx, wy);
                mbar_expect_tx(bar, 2u * kBoxLW * kBoxLH);
                tma_load_2d(s_win[warp], a.tm, wx & ~7, wy, bar);
            }
        }
    };
    const int maxv2 = ((1 << a.bd) - 1) * 0x00010001;
    const int s1 = min(4, a.bd - 8), s2 = max(8, 20 - a.bd);
    long long t0 = ((long long)blockIdx.x * 8 + warp) * PER;
    if (t0 < total) issue(t0);
    int phase = 0;
    for (; t0 < total; t0 += step) {
        const int hw = CHROMA ? lane >> 4 : 0;                      // half-warp = which of the step's tiles
        const long long t = t0 + hw;
        const bool have = t < total;
        const long long tt = have ? t : t0;
        const int b = (int)(tt / tiles), ti = (int)(tt % tiles);
        const int tx = (ti % tiles_x) * tw, ty = (ti / tiles_x) * th;
        const int gx = a.mv[b * 4 + 0], gy = a.mv[b * 4 + 1], ox = a.mv[b * 4 + 2], oy = a.mv[b * 4 + 3];
        // the variant follows the UNCLIPPED vector: a direction it calls integer is not filtered whatever phase the clipped vector has
        // (phase-0 taps with shift 6 are an exact copy), and nn differs from n0 / 0n in shifts and rounding only
        const bool fx = (ox & fmask) != 0, fy = (oy & fmask) != 0, two_d = fx && fy;
        const int phx = fx ? (gx & fmask) : 0, phy = fy ? (gy & fmask) : 0;
        int wx, wy;
        origin(tt, wx, wy);
        const int offx = wx & 7, par = offx & 1;
        mbar_wait(bar, phase & 1);
        phase++;
        pel *dst0 = a.out + (size_t)b * a.w * a.h + (size_t)ty * a.w + tx;
        if (!CHROMA) {
            if (lane < 24) {
                const int half = lane >= 12 ? 1 : 0, rp = lane - 12 * half;
                if (half * 8 < tw && 2 * rp < th + 7) {
                    const int *win = (const int *)s_win[warp] + (offx >> 1) + half * 4;
                    const Taps5 te = ld_taps5(phx, par), to = ld_taps5(phx, par + 1);
                    const int sh = two_d ? s1 : 6;
                    int hv[2][8];
#pragma unroll
                    for (int rr = 0; rr < 2; rr++) {
                        const int *rowp = win + (2 * rp + rr) * kWinLStrideW;
                        int q[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) q[j] = rowp[j];
#pragma unroll
                        for (int o = 0; o < 4; o++) {
                            hv[rr][2 * o] = fir5(te, q[o], q[o + 1], q[o + 2], q[o + 3], q[o + 4 < 8 ? o + 4 : 7], 0) >> sh;
                            hv[rr][2 * o + 1] = fir5(to, q[o], q[o + 1], q[o + 2], q[o + 3], q[o + 4 < 8 ? o + 4 : 7], 0) >> sh;
                        }
                    }
                    int4 *dst = (int4 *)(s_m2[warp] + half * 8 + rp * kM2LStrideW);
                    dst[0] = make_int4(pack16(hv[0][0], hv[1][0]), pack16(hv[0][1], hv[1][1]), pack16(hv[0][2], hv[1][2]), pack16(hv[0][3], hv[1][3]));
                    dst[1] = make_int4(pack16(hv[0][4], hv[1][4]), pack16(hv[0][5], hv[1][5]), pack16(hv[0][6], hv[1][6]), pack16(hv[0][7], hv[1][7]));
                }
            }
            __syncwarp();
            if (t0 + step < total) issue(t0 + step);
            const int cp = lane & 7, rg = lane >> 3;
            if (2 * cp < tw && 4 * rg < th) {
                const Taps5 te = ld_taps5(phy, 0), to = ld_taps5(phy, 1);
                const int sh = two_d ? s2 : 6, rnd = two_d ? (1 << (s2 - 1)) : 0;
                const int *m2 = s_m2[warp] + (2 * rg) * kM2LStrideW + 2 * cp;
                int P[6][2];
#pragma unroll
                for (int j = 0; j < 6; j++) { const int2 v = *(const int2 *)(m2 + j * kM2LStrideW); P[j][0] = v.x; P[j][1] = v.y; }
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    int e0 = fir5(te, P[q][0], P[q + 1][0], P[q + 2][0], P[q + 3][0], 0, rnd) >> sh;
                    int e1 = fir5(te, P[q][1], P[q + 1][1], P[q + 2][1], P[q + 3][1], 0, rnd) >> sh;
                    int o0 = fir5(to, P[q][0], P[q + 1][0], P[q + 2][0], P[q + 3][0], P[q + 4][0], rnd) >> sh;
                    int o1 = fir5(to, P[q][1], P[q + 1][1], P[q + 2][1], P[q + 3][1], P[q + 4][1], rnd) >> sh;
                    pel *d = dst0 + (size_t)(4 * rg + 2 * q) * a.w + 2 * cp;
                    if (4 * rg + 2 * q < th) *(int *)d = __vimin_s16x2_relu(pack16(e0, e1), maxv2);
                    if (4 * rg + 2 * q + 1 < th) *(int *)(d + a.w) = __vimin_s16x2_relu(pack16(o0, o1), maxv2);
                }
            }
            __syncwarp();           // the pair buffer is rewritten by the next tile
        } else {
            // both stages in one pass: a lane filters its 5 rows horizontally from the window and vertically from registers (2 columns x 2 rows)
            const int cp = lane & 3, rg = (lane >> 2) & 3;
            if (have && 2 * cp < tw && 2 * rg < th) {
                const int *win = (const int *)(s_win[warp] + hw * 640) + (offx >> 1) + cp + (2 * rg) * kWinCStrideW;
                const Taps3 he = ld_taps3(phx, par), ho = ld_taps3(phx, par + 1);
                const int sh1 = two_d ? s1 : 6;
                int h0[6], h1[6];
#pragma unroll
                for (int r = 0; r < 5; r++) {
                    const int q0 = win[r * kWinCStrideW], q1 = win[r * kWinCStrideW + 1], q2 = win[r * kWinCStrideW + 2];
                    h0[r] = fir3(he, q0, q1, q2, 0) >> sh1;
                    h1[r] = fir3(ho, q0, q1, q2, 0) >> sh1;
                }
                h0[5] = h1[5] = 0;
                int P[3][2];
#pragma unroll
                for (int j = 0; j < 3; j++) { P[j][0] = pack16(h0[2 * j], h0[2 * j + 1]); P[j][1] = pack16(h1[2 * j], h1[2 * j + 1]); }
                const Taps3 te = ld_taps3(phy, 0), to = ld_taps3(phy, 1);
                const int sh = two_d ? s2 : 6, rnd = two_d ? (1 << (s2 - 1)) : 0;
                const int e0 = fir3(te, P[0][0], P[1][0], 0, rnd) >> sh, e1 = fir3(te, P[0][1], P[1][1], 0, rnd) >> sh;
                const int o0 = fir3(to, P[0][0], P[1][0], P[2][0], rnd) >> sh, o1 = fir3(to, P[0][1], P[1][1], P[2][1], rnd) >> sh;
                pel *d = dst0 + (size_t)(2 * rg) * a.w + 2 * cp;
                *(int *)d = __vimin_s16x2_relu(pack16(e0, e1), maxv2);
                if (2 * rg + 1 < th) *(int *)(d + a.w) = __vimin_s16x2_relu(pack16(o0, o1), maxv2);
            }
            __syncwarp();
            if (t0 + step < total) issue(t0 + step);
        }
    }
}

inline int launch_mc_blocks(const CUtensorMap *tm, int pad, int plane, const int *mv, pel *out, int n, int w, int h, int bd, int main_tables,
                            cudaStream_t st)
{
    if (w < 2 || h < 2 || w > 128 || h > 128 || (w & (w - 1)) || (h & (h - 1))) return XB200_ERR_INVALID_ARGUMENT;
    const int T = plane ? 8 : 16;
    const int tw = w < T ? w : T, th = h < T ? h : T;
    const long long tiles = (long long)n * (w / tw) * (h / th);
    long long grid = (tiles + (plane ? 15 : 7)) / (plane ? 16 : 8);
    if (grid > 148 * 8) grid = 148 * 8;
    McBlocksArgs a;
    a.tm = tm; a.mv = mv; a.out = out; a.n = n; a.w = w; a.h = h; a.bd = bd; a.main_tables = main_tables; a.pad = pad;
    if (plane) k_mc_blocks<true><<<(int)grid, 256, 0, st>>>(a);
    else       k_mc_blocks<false><<<(int)grid, 256, 0, st>>>(a);
    return 1;
}

}  // namespace xb
