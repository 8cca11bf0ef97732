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

constexpr int kAlfTabBytes = 100 * 16 * 2;

struct AlfArgs {
    const pel *sy, *su, *sv;     // pre-ALF copy (same strides as the picture)
    pel *dy, *du, *dv;
    int s_l, s_c, w, h, log2_ctu, bd, w_ctu;
    const uint8_t *ctb_flag;     // device, one byte per CTU, or nullptr
    int16_t coef_c[7];
    uint8_t enable[3];
    int n_tile_cols, n_tile_rows, tile_across;      // xb200_set_tiles: boundaries in CTUs
    uint16_t tile_col_bd[XB200_MAX_TILE_COLS + 1], tile_row_bd[XB200_MAX_TILE_ROWS + 1];
    const int4 *ftab;            // device: the 100 luma filters (class x 4 + transpose), 16 int16 each, coefficients in filtering order.  Kept
                                 // per context and uploaded only when the APS changes (as kernel parameters, the per-thread indexed copy
                                 // into shared memory was 11 % of the kernel's stall samples: divergent constant-bank reads)
};

constexpr int kAlfT = 32;        // luma tile edge; divides every CTU size the Main profile allows (32, 64, 128)

__constant__ uint8_t c_alf_th[16] = {0, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4};
__constant__ uint8_t c_alf_trans[8] = {0, 1, 0, 2, 2, 3, 1, 3};

// Where a window sample comes from.  The reference filters every CTU through a window cut from a copy of its TILE that was extended by
// replication (alf_copy_and_extend_tile :805), and builds the 3-sample margins of the window per CTU side (alf_process_tile :1000-1046):
// a side that is "available" takes the copy as it is - real neighbours inside the tile, the replicated border outside it; a side that is
// not mirrors the CTU's own samples (rows: the assembled window rows, margins included).  Which sides are available is decided against
// the tile (loop_filter_across_tiles_enabled_flag == 0, and always for a single tile) or against the picture (flag == 1;
// tile_boundary_check :844 is then handed width - 1 / height - 1, so only the picture's left and top sides ever mirror).
struct AlfGeom {
    int cx0, cx1, cy0, cy1;      // the CTU (clipped to the picture), luma or chroma units
    int tx0, tx1, ty0, ty1;      // its tile (clipped to the picture)
    bool al, ar, at, ab;         // sides of the CTU whose margin comes from the copy
};
__device__ __forceinline__ AlfGeom alf_geom(const AlfArgs &a, int x0, int y0)
{
    AlfGeom g;
    const int ctu = 1 << a.log2_ctu, cxi = x0 >> a.log2_ctu, cyi = y0 >> a.log2_ctu;
    g.cx0 = cxi << a.log2_ctu; g.cy0 = cyi << a.log2_ctu;
    g.cx1 = min(g.cx0 + ctu, a.w); g.cy1 = min(g.cy0 + ctu, a.h);
    int tc = 0, tr = 0;
    while (tc + 1 < a.n_tile_cols && cxi >= (int)a.tile_col_bd[tc + 1]) tc++;
    while (tr + 1 < a.n_tile_rows && cyi >= (int)a.tile_row_bd[tr + 1]) tr++;
    g.tx0 = (int)a.tile_col_bd[tc] << a.log2_ctu; g.ty0 = (int)a.tile_row_bd[tr] << a.log2_ctu;
    g.tx1 = min((int)a.tile_col_bd[tc + 1] << a.log2_ctu, a.w); g.ty1 = min((int)a.tile_row_bd[tr + 1] << a.log2_ctu, a.h);
    if (a.tile_across) { g.al = g.cx0 != 0; g.ar = true; g.at = g.cy0 != 0; g.ab = true; }
    else { g.al = g.cx0 != g.tx0; g.ar = g.cx1 != g.tx1; g.at = g.cy0 != g.ty0; g.ab = g.cy1 != g.ty1; }
    return g;
}
// Sample (gy, gx) of the window of that CTU; sh = 0 luma, 1 chroma (the geometry is in luma units and even)
__device__ __forceinline__ pel alf_sample(const pel *p, int s, const AlfGeom &g, int sh, int gy, int gx)
{
    const int cx0 = g.cx0 >> sh, cx1 = g.cx1 >> sh, cy0 = g.cy0 >> sh, cy1 = g.cy1 >> sh;
    bool own_row = gy >= cy0 && gy < cy1;
    if (gy < cy0) { if (!g.at) { gy = 2 * cy0 - gy; own_row = true; } }
    else if (gy >= cy1) { if (!g.ab) { gy = 2 * cy1 - gy - 2; own_row = true; } }
    if (own_row) {
        if (gx < cx0) { if (!g.al) gx = 2 * cx0 - gx; }
        else if (gx >= cx1) { if (!g.ar) gx = 2 * cx1 - gx - 2; }
    }
    gx = min(max(gx, g.tx0 >> sh), (g.tx1 >> sh) - 1);
    gy = min(max(gy, g.ty0 >> sh), (g.ty1 >> sh) - 1);
    return p[(size_t)gy * s + gx];
}

// The pre-ALF copy: the sample rows of the three planes (padding columns included, so each plane is one contiguous range), 16 bytes
// per thread and step, one launch.  (Three cudaMemcpy2DAsync calls took 36 us of the 123 us an ALF pass cost in round 1.)
struct AlfCopyArgs { const int4 *src[3]; int4 *dst[3]; unsigned n[3]; };
__global__ void __launch_bounds__(256) k_alf_copy(const __grid_constant__ AlfCopyArgs a)
{
    const unsigned step = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    xb_grid_wait();
    xb_grid_release();
#pragma unroll 1
    for (int pl = 0; pl < 3; pl++) {
        const int4 *__restrict__ s = a.src[pl];
        int4 *__restrict__ d = a.dst[pl];
        const unsigned n = a.n[pl];
        unsigned i = t0;
        for (; i + 3 * step < n; i += 4 * step) {          // four independent 16-byte loads in flight per thread
            const int4 v0 = s[i], v1 = s[i + step], v2 = s[i + 2 * step], v3 = s[i + 3 * step];
            d[i] = v0; d[i + step] = v1; d[i + 2 * step] = v2; d[i + 3 * step] = v3;
        }
        for (; i < n; i += step) d[i] = s[i];
    }
}

// Window samples are kept as 32-bit words: win[r][j] is luma sample (y0 - 3 + r, x0 - 4 + j) - the 16-byte-aligned superset of the
// 38 columns a 32x32 tile needs - so the filter phase reads a row of its 7x7 diamond as 16-byte shared-memory loads (17 per thread for 4
// output samples; the 16-bit window of round 1 took 100 two-byte loads) and rows of 40 words put a quarter-warp's 8 loads on 32
// distinct banks.
constexpr int kAlfWinW = kAlfT + 8;
constexpr int kAlfCW = kAlfT / 2 + 8;       // chroma window: samples xc0 - 4 .. xc0 + 19 (aligned superset of xc0 - 2 .. xc0 + 17)
constexpr int kAlfCH = kAlfT / 2 + 4;

__global__ void __launch_bounds__(256) k_alf(const __grid_constant__ AlfArgs a)
{
    __shared__ __align__(16) int win[(kAlfT + 6) * kAlfWinW];
    __shared__ ushort4 cell[kAlfT / 2 + 2][kAlfT / 2 + 2];   // per 2x2 cell: sum of |vertical|, |horizontal|, |diag0|, |diag1| Laplacians
    __shared__ __align__(16) int4 ftab[kAlfTabBytes / 16];   // the picture's 100 luma filters
    __shared__ uint8_t fsel[64];                             // filter of every 4x4 block of the tile: class x 4 + transpose

    const int t = threadIdx.x;
    const int x0 = blockIdx.x * kAlfT, y0 = blockIdx.y * kAlfT;
    xb_grid_wait();
    xb_grid_release();
    const int tw = min(kAlfT, a.w - x0), th = min(kAlfT, a.h - y0);
    const AlfGeom g = alf_geom(a, x0, y0);
    const bool luma_on = a.enable[0] && (!a.ctb_flag || a.ctb_flag[(y0 >> a.log2_ctu) * a.w_ctu + (x0 >> a.log2_ctu)]);
    const int maxv = (1 << a.bd) - 1;

    // Tiles whose window lies inside the picture need none of the mirroring rules: 8-byte global loads of the aligned superset of every
    // window row (the planes start 16-byte aligned and x0 is a multiple of 32), one 16-byte shared-memory store each.
    const bool interior = tw == kAlfT && th == kAlfT && x0 - 4 >= g.tx0 && x0 + kAlfT + 4 <= g.tx1 && y0 - 3 >= g.ty0 && y0 + kAlfT + 3 <= g.ty1;
    // the chroma windows of an interior tile (2 planes x 20 rows x 6 8-byte words = 240 loads, one per thread) are fetched now and parked in
    // registers: their latency hides behind the luma phases (the store that waited for them was 7 % of the stall samples)
    static_assert(2 * kAlfCH * (kAlfCW / 4) <= 256, "one chroma window word per thread");
    const int cpre_pl = t / (kAlfCH * (kAlfCW / 4)), cpre_k = t - cpre_pl * (kAlfCH * (kAlfCW / 4)), cpre_r = cpre_k / (kAlfCW / 4), cpre_q = cpre_k - cpre_r * (kAlfCW / 4);
    const bool cpre_on = interior && t < 2 * kAlfCH * (kAlfCW / 4) && a.enable[1 + (cpre_pl & 1)];
    short4 cpre = make_short4(0, 0, 0, 0);
    if (cpre_on) cpre = *(const short4 *)((cpre_pl ? a.sv : a.su) + (size_t)((y0 >> 1) - 2 + cpre_r) * a.s_c + (x0 >> 1) - 4 + 4 * cpre_q);
    if (luma_on) {
        if (t < kAlfTabBytes / 16) ftab[t] = __ldg(a.ftab + t);
        if (interior) {
            for (int i = t; i < (kAlfT + 6) * (kAlfWinW / 4); i += 256) {
                const int r = i / (kAlfWinW / 4), q = i - r * (kAlfWinW / 4);
                const short4 v = *(const short4 *)(a.sy + (size_t)(y0 - 3 + r) * a.s_l + x0 - 4 + 4 * q);
                *(int4 *)(win + r * kAlfWinW + 4 * q) = make_int4((int)(unsigned short)v.x, (int)(unsigned short)v.y, (int)(unsigned short)v.z, (int)(unsigned short)v.w);
            }
        } else
        for (int i = t; i < (kAlfT + 6) * (kAlfT + 6); i += 256) {
            const int r = i / (kAlfT + 6), c = i - r * (kAlfT + 6);
            if (r < th + 6 && c < tw + 6) win[r * kAlfWinW + c + 1] = alf_sample(a.sy, a.s_l, g, 0, y0 - 3 + r, x0 - 3 + c);
        }
        __syncthreads();
        // Laplacians per 2x2 cell over rows/cols -2 .. +size+1 (alf_derive_classification_blk :60-106).  A thread takes two neighbouring
        // cells: their 4x6 patch is the 4x8 words of two 16-byte loads per row, so the 18 x 9 tasks of a tile are one step of the CTA.
        // (Always the cells of a full tile: cells beyond a partial tile read window words nobody wrote and feed no valid block.)
        constexpr int ncp = (kAlfT / 2 + 2) / 2;
        if (t < (kAlfT / 2 + 2) * ncp) {
            const int cr = t / ncp, cp = t - cr * ncp;
            int pt[4][8];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int4 *rowp = (const int4 *)(win + (2 * cr + r) * kAlfWinW + 4 * cp);
                const int4 u = rowp[0], v = rowp[1];
                pt[r][0] = u.x; pt[r][1] = u.y; pt[r][2] = u.z; pt[r][3] = u.w; pt[r][4] = v.x; pt[r][5] = v.y; pt[r][6] = v.z; pt[r][7] = v.w;
            }
#pragma unroll
            for (int ce = 0; ce < 2; ce++) {
                int sv = 0, sh = 0, sd0 = 0, sd1 = 0;
#pragma unroll
                for (int dy = 1; dy < 3; dy++)
#pragma unroll
                    for (int dx = 2 + 2 * ce; dx < 4 + 2 * ce; dx++) {      // the cell's samples sit at words 2, 3 (first cell) and 4, 5 (second) of the patch
                        const int p2 = (int)(int16_t)(pt[dy][dx] << 1);
                        sv += abs(p2 - pt[dy - 1][dx] - pt[dy + 1][dx]);
                        sh += abs(p2 - pt[dy][dx - 1] - pt[dy][dx + 1]);
                        sd0 += abs(p2 - pt[dy - 1][dx - 1] - pt[dy + 1][dx + 1]);
                        sd1 += abs(p2 - pt[dy + 1][dx - 1] - pt[dy - 1][dx + 1]);
                    }
                cell[cr][2 * cp + ce] = make_ushort4((unsigned short)sv, (unsigned short)sh, (unsigned short)sd0, (unsigned short)sd1);
            }
        }
        __syncthreads();
        // four threads per 4x4 block: each sums one row of the block's 4x4 cells (its 8x8 window), a butterfly adds the rows; class and
        // transpose (:108-206) select one of the staged filters
        {
            const int blk = t >> 2, part = t & 3, by = blk >> 3, bx = blk & 7;
            int sv = 0, sh = 0, sd0 = 0, sd1 = 0;
            const bool valid = by * 4 < th && bx * 4 < tw;
            if (valid) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const ushort4 q = cell[2 * by + part][2 * bx + j];
                    sv += q.x; sh += q.y; sd0 += q.z; sd1 += q.w;
                }
            }
#pragma unroll
            for (int d = 1; d < 4; d <<= 1) {
                sv += __shfl_xor_sync(0xffffffffu, sv, d); sh += __shfl_xor_sync(0xffffffffu, sh, d);
                sd0 += __shfl_xor_sync(0xffffffffu, sd0, d); sd1 += __shfl_xor_sync(0xffffffffu, sd1, d);
            }
            if (valid) {
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
                if (part == 0) fsel[blk] = (uint8_t)(cls * 4 + tr);
            }
        }
        __syncthreads();
        // thread = one row of one 4x4 block
        {
            const int blk = t >> 2, by = blk >> 3, bx = blk & 7;
            const int r = by * 4 + (t & 3), c0 = bx * 4;
            if (r < th && c0 < tw) {
                int f[16];
                {
                    const int sel = fsel[blk];
                    const int4 fa = ftab[2 * sel], fb = ftab[2 * sel + 1];
                    const int fw[8] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
#pragma unroll
                    for (int i = 0; i < 8; i++) { f[2 * i] = (int)(int16_t)(fw[i] & 0xffff); f[2 * i + 1] = fw[i] >> 16; }
                }
                // rows R-3 .. R+3 of the window, columns c0 .. c0+11 (sample k of the row sits at column c0 + 4 + k)
                const int *base = win + (r + 3) * kAlfWinW + c0;
                int m0[12], m1[2][12], m2[2][12], m3[2][4];
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    const int4 v = *(const int4 *)(base + 4 * q);
                    m0[4 * q] = v.x; m0[4 * q + 1] = v.y; m0[4 * q + 2] = v.z; m0[4 * q + 3] = v.w;
                }
#pragma unroll
                for (int s2 = 0; s2 < 2; s2++) {
                    const int *b1 = base + (s2 ? 1 : -1) * kAlfWinW, *b2 = base + (s2 ? 2 : -2) * kAlfWinW, *b3 = base + (s2 ? 3 : -3) * kAlfWinW;
#pragma unroll
                    for (int q = 0; q < 3; q++) {
                        const int4 v = *(const int4 *)(b1 + 4 * q);
                        m1[s2][4 * q] = v.x; m1[s2][4 * q + 1] = v.y; m1[s2][4 * q + 2] = v.z; m1[s2][4 * q + 3] = v.w;
                    }
                    {
                        const int4 v = *(const int4 *)(b2 + 4);
                        m2[s2][3] = b2[3]; m2[s2][4] = v.x; m2[s2][5] = v.y; m2[s2][6] = v.z; m2[s2][7] = v.w; m2[s2][8] = b2[8];
                    }
                    {
                        const int4 v = *(const int4 *)(b3 + 4);
                        m3[s2][0] = v.x; m3[s2][1] = v.y; m3[s2][2] = v.z; m3[s2][3] = v.w;
                    }
                }
                pel out[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int C = 4 + k;         // s2 = 0: rows above (dy < 0), s2 = 1: rows below (dy > 0)
                    int sum = f[0] * (m3[1][k] + m3[0][k])
                            + f[1] * (m2[1][C + 1] + m2[0][C - 1]) + f[2] * (m2[1][C] + m2[0][C]) + f[3] * (m2[1][C - 1] + m2[0][C + 1])
                            + f[4] * (m1[1][C + 2] + m1[0][C - 2]) + f[5] * (m1[1][C + 1] + m1[0][C - 1]) + f[6] * (m1[1][C] + m1[0][C])
                            + f[7] * (m1[1][C - 1] + m1[0][C + 1]) + f[8] * (m1[1][C - 2] + m1[0][C + 2])
                            + f[9] * (m0[C + 3] + m0[C - 3]) + f[10] * (m0[C + 2] + m0[C - 2]) + f[11] * (m0[C + 1] + m0[C - 1]) + f[12] * m0[C];
                    out[k] = (pel)min(max((sum + 256) >> 9, 0), maxv);
                }
                *(short4 *)(a.dy + (size_t)(y0 + r) * a.s_l + x0 + c0) = make_short4(out[0], out[1], out[2], out[3]);
            }
        }
    }

    // chroma: the 16x16 tiles of both planes in one phase, +-2 window, thread = 2 neighbouring samples of one plane
    const int xc0 = x0 >> 1, yc0 = y0 >> 1, twc = tw >> 1, thc = th >> 1;
    if (!a.enable[1] && !a.enable[2]) return;
    __syncthreads();
    if (interior) {
        if (cpre_on) *(int4 *)(win + (cpre_pl * kAlfCH + cpre_r) * kAlfCW + 4 * cpre_q) =
            make_int4((int)(unsigned short)cpre.x, (int)(unsigned short)cpre.y, (int)(unsigned short)cpre.z, (int)(unsigned short)cpre.w);
    } else
    for (int i = t; i < 2 * kAlfCH * kAlfCH; i += 256) {
        const int pl = i / (kAlfCH * kAlfCH), k = i - pl * (kAlfCH * kAlfCH), r = k / kAlfCH, c = k - r * kAlfCH;
        if (a.enable[1 + pl] && r < thc + 4 && c < twc + 4)
            win[(pl * kAlfCH + r) * kAlfCW + c + 2] = alf_sample(pl ? a.sv : a.su, a.s_c, g, 1, yc0 - 2 + r, xc0 - 2 + c);
    }
    __syncthreads();
    {
        const int pl = t >> 7, r = (t >> 3) & 15, c = (t & 7) * 2;
        if (a.enable[1 + pl] && r < thc && c < twc) {
            // sample c of the tile sits at window column c + 4; rows r .. r+4 of the window are dy = -2 .. 2
            const int *base = win + (pl * kAlfCH + r + 2) * kAlfCW + c + 4;
            int n0[6], n1[2][6], n2[2][2];
#pragma unroll
            for (int q = 0; q < 3; q++) { const int2 v = *(const int2 *)(base - 2 + 2 * q); n0[2 * q] = v.x; n0[2 * q + 1] = v.y; }
#pragma unroll
            for (int s2 = 0; s2 < 2; s2++) {
                const int *b1 = base + (s2 ? 1 : -1) * kAlfCW, *b2 = base + (s2 ? 2 : -2) * kAlfCW;
#pragma unroll
                for (int q = 0; q < 3; q++) { const int2 v = *(const int2 *)(b1 - 2 + 2 * q); n1[s2][2 * q] = v.x; n1[s2][2 * q + 1] = v.y; }
                const int2 v = *(const int2 *)b2;
                n2[s2][0] = v.x; n2[s2][1] = v.y;
            }
            int o[2];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int C = 2 + k;
                const int sum = a.coef_c[0] * (n2[1][k] + n2[0][k]) + a.coef_c[1] * (n1[1][C + 1] + n1[0][C - 1]) + a.coef_c[2] * (n1[1][C] + n1[0][C])
                              + a.coef_c[3] * (n1[1][C - 1] + n1[0][C + 1]) + a.coef_c[4] * (n0[C + 2] + n0[C - 2]) + a.coef_c[5] * (n0[C + 1] + n0[C - 1])
                              + a.coef_c[6] * n0[C];
                o[k] = min(max((sum + 256) >> 9, 0), maxv);
            }
            pel *dst = (pl ? a.dv : a.du) + (size_t)(yc0 + r) * a.s_c + xc0 + c;
            if (c + 1 < twc) *(int *)dst = o[0] | (o[1] << 16);
            else dst[0] = (pel)o[0];
        }
    }
}

}  // namespace xb
