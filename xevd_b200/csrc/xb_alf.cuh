// Adaptive loop filter (Main profile, sps->tool_alf).
// Replaces mctx->fn_alf = xevd_alf -> alf_process -> alf_process_tile (src_main/xevdm.c:2105, src_main/xevdm_alf.c:901-1165,
// 1167-1273) for a picture that is one tile: 4x4-block classification (alf_derive_classification_blk, :38-208), the 7x7 diamond
// luma filter with 25 classes x 4 transposes (alf_filter_blk_7, :210-339) and the 5x5 diamond chroma filter (alf_filter_blk_5,
// :341-429).
//
// The reference filters every CTU from a copy of the pre-ALF picture, extended by replication, through a per-CTU window whose
// 3-sample margins are mirrored where no neighbour CTU exists.  With a single tile that only happens at the picture border, so the
// window rules collapse to a function of picture coordinates (alf_sample below) and the work can be tiled independently of the
// CTU size: one CTA per 32x32 luma tile (+ its two 16x16 chroma tiles), reading the pre-ALF copy and writing the picture.
#pragma once
#include "xb_common.cuh"

namespace xb {

struct AlfArgs {
    const pel *sy, *su, *sv;     // pre-ALF copy (same strides as the picture)
    pel *dy, *du, *dv;
    int s_l, s_c, w, h, log2_ctu, bd, w_ctu;
    const uint8_t *ctb_flag;     // device, one byte per CTU, or nullptr
    int16_t coef_l[25][13];
    int16_t coef_c[7];
    uint8_t enable[3];
};

constexpr int kAlfT = 32;        // luma tile edge; divides every CTU size the Main profile allows (32, 64, 128)

__constant__ uint8_t c_alf_perm[4][13] = {{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12}, {9, 4, 10, 8, 1, 5, 11, 7, 3, 0, 2, 6, 12},
                                          {0, 3, 2, 1, 8, 7, 6, 5, 4, 9, 10, 11, 12}, {9, 8, 10, 4, 3, 7, 11, 5, 1, 0, 2, 6, 12}};
__constant__ uint8_t c_alf_th[16] = {0, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4};
__constant__ uint8_t c_alf_trans[8] = {0, 1, 0, 2, 2, 3, 1, 3};

// Sample (gy, gx) of the window of the CTU whose rows are [cy0, cy1), gy in [-3, H+3), gx in [-3, W+3).
// Rows outside the picture mirror (-1 -> 1, H -> H-2: alf_process_tile :1098-1110, :1133-1146); columns outside the picture mirror
// on rows that belong to the CTU (:1060-1096) and replicate on rows taken from the CTU above / below, which come from the
// replicated copy (:1112-1131, :1148-1160, alf_copy_and_extend_tile :805).
__device__ __forceinline__ pel alf_sample(const pel *p, int s, int W, int H, int gy, int gx, int cy0, int cy1)
{
    bool own_row = gy >= cy0 && gy < cy1;
    if (gy < 0) { gy = -gy; own_row = true; }
    else if (gy >= H) { gy = 2 * H - gy - 2; own_row = true; }
    if (gx < 0) gx = own_row ? -gx : 0;
    else if (gx >= W) gx = own_row ? 2 * W - gx - 2 : W - 1;
    return p[(size_t)gy * s + gx];
}

__global__ void __launch_bounds__(256) k_alf(const __grid_constant__ AlfArgs a)
{
    __shared__ pel win[kAlfT + 6][kAlfT + 8];
    __shared__ ushort4 cell[kAlfT / 2 + 2][kAlfT / 2 + 2];   // per 2x2 cell: sum of |vertical|, |horizontal|, |diag0|, |diag1| Laplacians
    __shared__ int16_t filt[64][14];

    const int t = threadIdx.x;
    const int x0 = blockIdx.x * kAlfT, y0 = blockIdx.y * kAlfT;
    const int tw = min(kAlfT, a.w - x0), th = min(kAlfT, a.h - y0);
    const int ctu = 1 << a.log2_ctu;
    const int cy0 = (y0 >> a.log2_ctu) << a.log2_ctu, cy1 = min(cy0 + ctu, a.h);
    const bool luma_on = a.enable[0] && (!a.ctb_flag || a.ctb_flag[(y0 >> a.log2_ctu) * a.w_ctu + (x0 >> a.log2_ctu)]);
    const int maxv = (1 << a.bd) - 1;

    // Tiles whose window lies inside the picture need none of the mirroring rules: 8-byte loads of the aligned superset of every window
    // row (x0 - 4 .. x0 + 36; the planes start 16-byte aligned and x0 is a multiple of 32).  The per-sample path with its border logic
    // took 45 % of the kernel's instructions before (profiles/r1/alf_ncu_summary.txt).
    const bool interior = tw == kAlfT && th == kAlfT && x0 >= 4 && x0 + kAlfT + 4 <= a.w && y0 >= 3 && y0 + kAlfT + 3 <= a.h;
    if (luma_on) {
        if (interior) {
            for (int i = t; i < (kAlfT + 6) * 10; i += 256) {
                const int r = i / 10, q = i - r * 10;
                const short4 v = *(const short4 *)(a.sy + (size_t)(y0 - 3 + r) * a.s_l + x0 - 4 + 4 * q);
                pel *d = &win[r][4 * q - 1];                    // win[r][c] holds column x0 - 3 + c
                if (q > 0) d[0] = v.x;
                d[1] = v.y; d[2] = v.z;
                d[3] = v.w;                                     // q == 9: c 38 - inside the row's 40 entries, never read
            }
        } else
        for (int i = t; i < (kAlfT + 6) * (kAlfT + 6); i += 256) {
            const int r = i / (kAlfT + 6), c = i - r * (kAlfT + 6);
            if (r < th + 6 && c < tw + 6) win[r][c] = alf_sample(a.sy, a.s_l, a.w, a.h, y0 - 3 + r, x0 - 3 + c, cy0, cy1);
        }
        __syncthreads();
        // Laplacians per 2x2 cell over rows/cols -2 .. +size+1 (alf_derive_classification_blk :60-106)
        const int ncx = tw / 2 + 2, ncy = th / 2 + 2;
        for (int i = t; i < ncx * ncy; i += 256) {
            const int cr = i / ncx, cc = i - cr * ncx;
            int sv = 0, sh = 0, sd0 = 0, sd1 = 0;
#pragma unroll
            for (int dy = 0; dy < 2; dy++)
#pragma unroll
                for (int dx = 0; dx < 2; dx++) {
                    const int r = 2 * cr + 1 + dy, c = 2 * cc + 1 + dx;
                    const int p2 = (int16_t)(win[r][c] << 1);
                    sv += abs(p2 - win[r - 1][c] - win[r + 1][c]);
                    sh += abs(p2 - win[r][c - 1] - win[r][c + 1]);
                    sd0 += abs(p2 - win[r - 1][c - 1] - win[r + 1][c + 1]);
                    sd1 += abs(p2 - win[r + 1][c - 1] - win[r - 1][c + 1]);
                }
            cell[cr][cc] = make_ushort4((unsigned short)sv, (unsigned short)sh, (unsigned short)sd0, (unsigned short)sd1);
        }
        __syncthreads();
        // one thread per 4x4 block: 8x8 window sums, class and transpose (:108-206)
        if (t < 64) {
            const int by = t >> 3, bx = t & 7;
            if (by * 4 < th && bx * 4 < tw) {
                int sv = 0, sh = 0, sd0 = 0, sd1 = 0;
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const ushort4 q = cell[2 * by + i][2 * bx + j];
                        sv += q.x; sh += q.y; sd0 += q.z; sd1 += q.w;
                    }
                const int activity = min(15, (sv + sh) >> (a.bd - 2));
                int cls = c_alf_th[activity];
                int hv1, hv0, d1, d0, dir_hv, dir_d, hvd1, hvd0, main_dir, sec_dir;
                if (sv > sh) { hv1 = sv; hv0 = sh; dir_hv = 1; } else { hv1 = sh; hv0 = sv; dir_hv = 3; }
                if (sd0 > sd1) { d1 = sd0; d0 = sd1; dir_d = 0; } else { d1 = sd1; d0 = sd0; dir_d = 2; }
                if ((int)((unsigned)d1 * (unsigned)hv0) > (int)((unsigned)hv1 * (unsigned)d0)) { hvd1 = d1; hvd0 = d0; main_dir = dir_d; sec_dir = dir_hv; }
                else { hvd1 = hv1; hvd0 = hv0; main_dir = dir_hv; sec_dir = dir_d; }
                int strength = 0;
                if (hvd1 > 2 * hvd0) strength = 1;
                if (hvd1 * 2 > 9 * hvd0) strength = 2;
                if (strength) cls += (((main_dir & 1) << 1) + strength) * 5;
                const int tr = c_alf_trans[main_dir * 2 + (sec_dir >> 1)];
#pragma unroll
                for (int i = 0; i < 13; i++) filt[t][i] = a.coef_l[cls][c_alf_perm[tr][i]];
            }
        }
        __syncthreads();
        // thread = one row of one 4x4 block
        {
            const int blk = t >> 2, by = blk >> 3, bx = blk & 7;
            const int r = by * 4 + (t & 3), c0 = bx * 4;
            if (r < th && c0 < tw) {
                int f[13];
#pragma unroll
                for (int i = 0; i < 13; i++) f[i] = filt[blk][i];
                pel out[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int R = r + 3, C = c0 + k + 3;
#define P(dy, dx) ((int)win[R + (dy)][C + (dx)])
                    int sum = f[0] * (P(3, 0) + P(-3, 0))
                            + f[1] * (P(2, 1) + P(-2, -1)) + f[2] * (P(2, 0) + P(-2, 0)) + f[3] * (P(2, -1) + P(-2, 1))
                            + f[4] * (P(1, 2) + P(-1, -2)) + f[5] * (P(1, 1) + P(-1, -1)) + f[6] * (P(1, 0) + P(-1, 0)) + f[7] * (P(1, -1) + P(-1, 1)) + f[8] * (P(1, -2) + P(-1, 2))
                            + f[9] * (P(0, 3) + P(0, -3)) + f[10] * (P(0, 2) + P(0, -2)) + f[11] * (P(0, 1) + P(0, -1)) + f[12] * P(0, 0);
#undef P
                    out[k] = (pel)min(max((sum + 256) >> 9, 0), maxv);
                }
                *(short4 *)(a.dy + (size_t)(y0 + r) * a.s_l + x0 + c0) = make_short4(out[0], out[1], out[2], out[3]);
            }
        }
    }

    // chroma: 16x16 tile per plane, +-2 window, one sample per thread
    const int W_c = a.w >> 1, H_c = a.h >> 1, xc0 = x0 >> 1, yc0 = y0 >> 1, twc = tw >> 1, thc = th >> 1;
#pragma unroll 1
    for (int pl = 1; pl < 3; pl++) {
        if (!a.enable[pl]) continue;
        const pel *src = pl == 1 ? a.su : a.sv;
        pel *dst = pl == 1 ? a.du : a.dv;
        __syncthreads();
        if (interior) {
            // chroma window: columns xc0 - 2 .. xc0 + 17; aligned superset xc0 - 4 .. xc0 + 19 (xc0 is a multiple of 16)
            for (int i = t; i < (kAlfT / 2 + 4) * 6; i += 256) {
                const int r = i / 6, q = i - r * 6;
                const short4 v = *(const short4 *)(src + (size_t)(yc0 - 2 + r) * a.s_c + xc0 - 4 + 4 * q);
                pel *d = &win[r][4 * q - 2];                    // win[r][c] holds column xc0 - 2 + c
                if (q > 0) { d[0] = v.x; d[1] = v.y; }
                d[2] = v.z; d[3] = v.w;                         // q == 5: c 20, 21 - inside the row, never read
            }
        } else
        for (int i = t; i < (kAlfT / 2 + 4) * (kAlfT / 2 + 4); i += 256) {
            const int r = i / (kAlfT / 2 + 4), c = i - r * (kAlfT / 2 + 4);
            if (r < thc + 4 && c < twc + 4) win[r][c] = alf_sample(src, a.s_c, W_c, H_c, yc0 - 2 + r, xc0 - 2 + c, cy0 >> 1, cy1 >> 1);
        }
        __syncthreads();
        const int r = t >> 4, c = t & 15;
        if (r < thc && c < twc) {
            const int R = r + 2, C = c + 2;
#define P(dy, dx) ((int)win[R + (dy)][C + (dx)])
            int sum = a.coef_c[0] * (P(2, 0) + P(-2, 0)) + a.coef_c[1] * (P(1, 1) + P(-1, -1)) + a.coef_c[2] * (P(1, 0) + P(-1, 0)) + a.coef_c[3] * (P(1, -1) + P(-1, 1))
                    + a.coef_c[4] * (P(0, 2) + P(0, -2)) + a.coef_c[5] * (P(0, 1) + P(0, -1)) + a.coef_c[6] * P(0, 0);
#undef P
            dst[(size_t)(yc0 + r) * a.s_c + xc0 + c] = (pel)min(max((sum + 256) >> 9, 0), maxv);
        }
    }
}

}  // namespace xb
