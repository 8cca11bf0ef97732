// xb_recon.cuh -- per-picture CU reconstruction kernel (inter CUs): one CTA reconstructs one CTU.
//
// Replaces xevd_ctu_row_rec_mt -> xevd_recon_tree -> xevd_recon_unit (src_base/xevd.c:1470,1019,678;
// Main src_main/xevdm.c:2463,1854,1230) for CUs whose motion is already resolved:
//   phase A  residual:  coefficient stream -> dequant -> inverse DCT-2 (two passes) -> s16 residual in smem
//            (xevd_sub_block_itdq, src_base/xevd_itdq.c:544-621; xevdm_sub_block_itdq, xevdm_itdq.c:790-887)
//   phase B  prediction + reconstruction: warp per 16x16 tile, reference window staged in shared memory,
//            separable 8-tap / 4-tap interpolation, bi-prediction average, + residual, clip, store
//            (xevd_mc, src_base/xevd_mc.c:469-557; xevd_recon, src_base/xevd_recon.c:36-68)
//   phase C  publish per-SCU maps (xevd_set_dec_info, src_base/xevd_util.c:1574-1650)
#pragma once
#include "xb_common.cuh"
#include "xb_itdq.cuh"

namespace xb {

constexpr int kReconThreads = 256;
constexpr int kReconWarps = kReconThreads / 32;
constexpr int kMcScratchPerWarp = 1536;     // int16 elements: luma window 23x24 + intermediate 23x16 + slack

// shared memory carve-up for a CTU of size S = 1 << log2_ctu (S = 64: ~39 KB)
struct ReconSmem {
    int S, Sc;
    int *tmp_y, *tmp_u, *tmp_v;            // pass-1 results, row stride S+1 / Sc+1 (bank-conflict free)
    int16_t *res_y, *res_u, *res_v;        // residual, row stride S+2 / Sc+2
    uint16_t *cu_of_scu;                   // CU index (relative to the CTU's first CU) covering each SCU: owner of its luma samples and maps
    uint16_t *cu_of_scu_c;                 // owner of its chroma samples: differs inside a local dual tree node (luma leaves + one chroma-only CU)
    int16_t *mc;                           // per-warp interpolation scratch (aliases tmp_*)
    __device__ __forceinline__ void carve(unsigned char *base, int log2_ctu)
    {
        S = 1 << log2_ctu; Sc = S >> 1;
        tmp_y = (int *)base;
        tmp_u = tmp_y + S * (S + 1);
        tmp_v = tmp_u + Sc * (Sc + 1);
        mc = (int16_t *)base;
        size_t tmp_bytes = sizeof(int) * (S * (S + 1) + 2 * Sc * (Sc + 1));
        size_t mc_bytes = sizeof(int16_t) * kMcScratchPerWarp * kReconWarps;
        size_t a = tmp_bytes > mc_bytes ? tmp_bytes : mc_bytes;
        a = (a + 15) & ~(size_t)15;
        res_y = (int16_t *)(base + a);
        res_u = res_y + S * (S + 2);
        res_v = res_u + Sc * (Sc + 2);
        cu_of_scu = (uint16_t *)(res_v + Sc * (Sc + 2));
        cu_of_scu_c = cu_of_scu + (S / 4) * (S / 4);
    }
    static size_t bytes(int log2_ctu)
    {
        int S = 1 << log2_ctu, Sc = S >> 1;
        size_t tmp_bytes = sizeof(int) * (S * (S + 1) + 2 * Sc * (Sc + 1));
        size_t mc_bytes = sizeof(int16_t) * kMcScratchPerWarp * kReconWarps;
        size_t a = tmp_bytes > mc_bytes ? tmp_bytes : mc_bytes;
        a = (a + 15) & ~(size_t)15;
        return a + sizeof(int16_t) * (S * (S + 2) + 2 * Sc * (Sc + 2)) + 2 * sizeof(uint16_t) * (S / 4) * (S / 4) + 16;
    }
};

// ---- motion vector clipping: xevd_mv_clip (src_base/xevd_mc.c:435-467) ---------------------------------
__device__ __forceinline__ void mv_clip(int x, int y, int pic_w, int pic_h, int w, int h, int mvx, int mvy, int &cx, int &cy)
{
    const int qx = x << 2, qy = y << 2, qw = w << 2, qh = h << 2;
    const int lo = -(128 << 2), hx = (pic_w - 1 + 128) << 2, hy = (pic_h - 1 + 128) << 2;
    cx = mvx; cy = mvy;
    if (qx + mvx < lo) cx = lo - qx;
    if (qy + mvy < lo) cy = lo - qy;
    if (qx + mvx + qw - 4 > hx) cx = hx - qx - qw + 4;
    if (qy + mvy + qh - 4 > hy) cy = hy - qy - qh + 4;
    cx = (int16_t)cx; cy = (int16_t)cy;
}

// ---- one interpolated tile, computed by one warp ----------------------------------------------------------
// Tile of tw x th samples (powers of two, <= 16).  Lane l owns column (l % tw) and `rpl` consecutive rows
// starting at (l / tw) * rpl.  Result in pr[0..rpl).  NTAP = 8 (luma) or 4 (chroma).
//   ref     : sample at integer position of the tile's top-left output (before the -NTAP/2+1 tap offset)
//   cx, cy  : taps for the horizontal / vertical phase;  fx, fy : variant selected by the UNCLIPPED mv (T3)
//   ld(r, c): sample of the tile's full (th + NTAP - 1) x (tw + NTAP - 1) window, (0, 0) = HALF samples up-left of the tile's
//             integer position
template <int NTAP, typename Ld>
__device__ __forceinline__ void mc_tile_ld(Ld ld, const int16_t *cx, const int16_t *cy, bool fx, bool fy, int tw, int th, int rpl, int bd,
                                           int16_t *scr, int lane, int (&pr)[8])
{
    constexpr int HALF = NTAP / 2 - 1;
    const int maxv = (1 << bd) - 1;
    const int ltw = 31 - __clz(tw);                 // tile widths are powers of two: shifts instead of divisions in the sample loops
    const int col = lane & (tw - 1), r0 = (lane >> ltw) * rpl;
    const bool active = r0 < th;
    if (!fx && !fy) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (i < rpl && active) pr[i] = ld(r0 + i + HALF, col + HALF);
        return;
    }
    // stage the reference window: rows [-HALF, th+NTAP-1-HALF) if fy, cols [-HALF, tw+NTAP-1-HALF) if fx
    const int wrows = fy ? th + NTAP - 1 : th, wcols = fx ? tw + NTAP - 1 : tw;
    const int wstride = 24;
    const int ro = fy ? 0 : HALF, co = fx ? 0 : HALF;
    for (int r = 0; r < wrows; r++)
        if (lane < wcols) scr[r * wstride + lane] = (int16_t)ld(r + ro, lane + co);
    __syncwarp();
    int16_t *win = scr;
    int16_t *mid = scr + 23 * wstride;              // horizontal-pass output, row stride tw
    if (fx) {
        int c[NTAP];
#pragma unroll
        for (int t = 0; t < NTAP; t++) c[t] = cx[t];
        const int s1 = fy ? min(4, bd - 8) : 6;
        for (int idx = lane; idx < wrows * tw; idx += 32) {
            const int r = idx >> ltw, cc = idx & (tw - 1);
            int acc = 0;
#pragma unroll
            for (int t = 0; t < NTAP; t++) acc += c[t] * win[r * wstride + cc + t];
            acc >>= s1;
            if (!fy) acc = xb_clip3(0, maxv, acc);   // 1-D: (sum + 0) >> 6, clipped (xevd_mc.c:203, T1)
            mid[idx] = (int16_t)acc;                 // 2-D: kept as s16 (xevd_mc.c:243,264)
        }
        __syncwarp();
        if (!fy) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (i < rpl && active) pr[i] = mid[(r0 + i) * tw + col];
            __syncwarp();
            return;
        }
    }
    {
        int c[NTAP];
#pragma unroll
        for (int t = 0; t < NTAP; t++) c[t] = cy[t];
        const int16_t *src = fx ? mid : win;
        const int sst = fx ? tw : wstride;
        const int s2 = fx ? max(8, 20 - bd) : 6, rnd = fx ? (1 << (s2 - 1)) : 0;
        if (active) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (i < rpl) {
                    int acc = rnd;
#pragma unroll
                    for (int t = 0; t < NTAP; t++) acc += c[t] * src[(r0 + i + t) * sst + col];
                    pr[i] = xb_clip3(0, maxv, acc >> s2);
                }
            }
        }
    }
    __syncwarp();
}

template <int NTAP>
__device__ __forceinline__ void mc_tile(const pel *__restrict__ ref, int stride, const int16_t *cx, const int16_t *cy,
                                        bool fx, bool fy, int tw, int th, int rpl, int bd, int16_t *scr, int lane, int (&pr)[8])
{
    constexpr int HALF = NTAP / 2 - 1;
    const pel *base = ref - HALF * stride - HALF;
    mc_tile_ld<NTAP>([&](int r, int c) { return (int)base[r * stride + c]; }, cx, cy, fx, fy, tw, th, rpl, bd, scr, lane, pr);
}

// prediction of one tile from one or two references, result pr[] (bi-pred averaged): xevd_mc body
template <int NTAP>
__device__ __forceinline__ void pred_tile(const XbFrameArgs &a, const XB200_CU &cu, int plane, int tx, int ty, int tw, int th,
                                          int rpl, int16_t *scr, int lane, int (&pr)[8])
{
    constexpr bool LUMA = NTAP == 8;
    const int cuw = 1 << cu.log2w, cuh = 1 << cu.log2h;
    const int bd = LUMA ? a.bd_l : a.bd_c;
    const int stride = LUMA ? a.s_l : a.s_c;
    int mvc[2][2];
    mv_clip(cu.x, cu.y, a.w, a.h, cuw, cuh, cu.mv[0][0], cu.mv[0][1], mvc[0][0], mvc[0][1]);
    mv_clip(cu.x, cu.y, a.w, a.h, cuw, cuh, cu.mv[1][0], cu.mv[1][1], mvc[1][0], mvc[1][1]);
    bool use[2] = {cu.refi[0] >= 0, cu.refi[1] >= 0};
    // identical motion: same picture (POC) and same clipped vector -> list 0 only (xevd_mc.c:513-519)
    if (use[0] && use[1] && a.ref_poc[0][cu.refi[0]] == a.ref_poc[1][cu.refi[1]] && mvc[0][0] == mvc[1][0] && mvc[0][1] == mvc[1][1])
        use[1] = false;
    int n = 0;
    int p0[8];
#pragma unroll 1
    for (int l = 0; l < 2; l++) {
        if (!use[l]) continue;
        const int ri = cu.refi[l];
        const pel *rp = plane == 0 ? a.ref_y[l][ri] : (plane == 1 ? a.ref_u[l][ri] : a.ref_v[l][ri]);
        const int mvx = mvc[l][0], mvy = mvc[l][1];
        int ix, iy, phx, phy;
        bool fx, fy;
        if (LUMA) {
            ix = cu.x + tx + (mvx >> 2); iy = cu.y + ty + (mvy >> 2);
            phx = (mvx & 3) << 2; phy = (mvy & 3) << 2;
            fx = (cu.mv[l][0] & 3) != 0; fy = (cu.mv[l][1] & 3) != 0;       // variant from the unclipped vector (T3)
        } else {
            ix = (cu.x >> 1) + tx + (mvx >> 3); iy = (cu.y >> 1) + ty + (mvy >> 3);
            phx = (mvx & 7) << 2; phy = (mvy & 7) << 2;
            fx = (cu.mv[l][0] & 7) != 0; fy = (cu.mv[l][1] & 7) != 0;
        }
        const int16_t *cx = LUMA ? c_mc_l[a.main_tables][phx] : c_mc_c[a.main_tables][phx];
        const int16_t *cy = LUMA ? c_mc_l[a.main_tables][phy] : c_mc_c[a.main_tables][phy];
        if (n == 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) p0[i] = pr[i];
        }
        mc_tile<NTAP>(rp + iy * stride + ix, stride, cx, cy, fx, fy, tw, th, rpl, bd, scr, lane, pr);      // one call site: code size matters here
        n++;
    }
    if (n == 2) {
#pragma unroll
        for (int i = 0; i < 8; i++) pr[i] = (p0[i] + pr[i] + 1) >> 1;      // xevd_average_16b_no_clip
    }
}

// ---- DMVR (Main, tool_dmvr): decoder-side motion vector refinement ----------------------------------------------------------
// xevdm_mc's DMVR branch (src_main/xevdm_mc.c:1860-2038): processDMVR (:1638-1825) per 16x16 sub-PU - bilinear search planes of both
// lists (xevdm_bl_mc_l), two rounds of a 5-point mirrored SAD search (xevd_DMVR_refine), parabolic sub-sample step
// (xevd_SubPelErrorSrfc), final 8/4-tap prediction from the window of the INITIAL vector extended by 2 (1) replicated samples
// (prefetch_for_mc + final_paddedMC_forDMVR).  One warp per sub-PU.

// does xevdm_mc refine this CU?  (:1893-1911)
__device__ __forceinline__ bool dmvr_applies(const XbFrameArgs &a, const XB200_CU &cu, int (&start)[2][2])
{
    if (!a.dmvr || cu.mode != XB200_MODE_INTER || !(cu.flags & XB200_CUF_DMVR) || cu.refi[0] < 0 || cu.refi[1] < 0) return false;
    const int w = 1 << cu.log2w, h = 1 << cu.log2h;
    if (w < 8 || h < 8) return false;
    const int p0 = a.ref_poc[0][cu.refi[0]], p1 = a.ref_poc[1][cu.refi[1]];
    const int d0 = a.poc - p0, d1 = a.poc - p1;
    if (!(d0 * d1 < 0 && abs(d0) == abs(d1))) return false;
    mv_clip(cu.x, cu.y, a.w, a.h, w, h, cu.mv[0][0], cu.mv[0][1], start[0][0], start[0][1]);
    mv_clip(cu.x, cu.y, a.w, a.h, w, h, cu.mv[1][0], cu.mv[1][1], start[1][0], start[1][1]);
    return !(p0 == p1 && start[0][0] == start[1][0] && start[0][1] == start[1][1]);
}

// Per-CU dispatch between the two inter kernels: a CU the throughput kernel (xb_recon2.cuh) has no code for
__device__ __forceinline__ bool cu_needs_generic(const XbFrameArgs &a, const XB200_CU &cu)
{
    if (a.ats && (cu.ats || (cu.flags & XB200_CUF_ATS_INTRA))) return true;        // DST-7 / DCT-8 lines, sub-block transform units
    if ((cu.flags & (XB200_CUF_LUMA | XB200_CUF_CHROMA)) != (XB200_CUF_LUMA | XB200_CUF_CHROMA)) return true;      // local dual tree: per-plane owners
    if (cu.mode == XB200_MODE_AFFINE) return true;
    int st[2][2];
    return dmvr_applies(a, cu, st);
}
// true when some CU of the CTU needs the generic kernel (all threads of the CTA call this)
__device__ __forceinline__ bool ctu_needs_generic(const XbFrameArgs &a, const XB200_CU *cus, int ncu, int tid, int nthreads)
{
    bool any = false;
    for (int i = tid; i < ncu; i += nthreads) any |= cu_needs_generic(a, cus[i]);
    return __syncthreads_or(any) != 0;
}

// one sample of xevdm_bl_mc_l at 1/16 position (gx, gy) + (j, i)
__device__ __forceinline__ int dmvr_bilinear(const pel *__restrict__ ref, int s, int gx, int gy, int i, int j, int bd)
{
    const int dx = gx & 15, dy = gy & 15, maxv = (1 << bd) - 1;
    const pel *p = ref + (ptrdiff_t)((gy >> 4) + i) * s + (gx >> 4) + j;
    const int cx0 = 64 - 4 * dx, cx1 = 4 * dx, cy0 = 64 - 4 * dy, cy1 = 4 * dy;
    if (!dx && !dy) return p[0];
    if (!dy) return xb_clip3(0, maxv, (cx0 * p[0] + cx1 * p[1]) >> 6);
    if (!dx) return xb_clip3(0, maxv, (cy0 * p[0] + cy1 * p[s]) >> 6);
    const int s1 = min(4, bd - 8), s2 = max(8, 20 - bd);
    const int t0 = (int16_t)((cx0 * p[0] + cx1 * p[1]) >> s1), t1 = (int16_t)((cx0 * p[s] + cx1 * p[s + 1]) >> s1);
    return xb_clip3(0, maxv, (cy0 * t0 + cy1 * t1 + (1 << (s2 - 1))) >> s2);
}

__device__ __forceinline__ int dmvr_div_q7(long long n, long long d)       // div_for_maxq7 (:1338-1375)
{
    const bool neg = n < 0;
    if (neg) n = -n;
    int q = 0;
    d <<= 3;
    if (n >= d) { n -= d; q++; }
    q <<= 1; d >>= 1;
    if (n >= d) { n -= d; q++; }
    q <<= 1;
    if (n >= (d >> 1)) q++;
    return neg ? -q : q;
}
__device__ __forceinline__ int dmvr_subpel(int c, int minus, int plus)     // one axis of xevd_SubPelErrorSrfc (:1376-1428)
{
    const long long num = (long long)((minus - plus) << 4), den = (long long)(minus + plus - (c << 1));
    if (den == 0) return 0;
    if (minus != c && plus != c) return dmvr_div_q7(num, den);
    return minus == c ? -8 : 8;
}
// mv_clip_only_one_ref_dmvr (:939-978)
__device__ __forceinline__ bool dmvr_clip_one(int x, int y, int pic_w, int pic_h, int w, int h, int mvx, int mvy, int &cx, int &cy)
{
    const int qx = x << 2, qy = y << 2, qw = w << 2, qh = h << 2;
    const int lo = -(128 << 2), hx = (pic_w - 1 + 128) << 2, hy = (pic_h - 1 + 128) << 2;
    bool f = false;
    cx = mvx; cy = mvy;
    if (qx + mvx < lo) { f = true; cx = lo - qx; }
    if (qy + mvy < lo) { f = true; cy = lo - qy; }
    if (qx + mvx + qw - 4 > hx) { f = true; cx = hx - qx - qw + 4; }
    if (qy + mvy + qh - 4 > hy) { f = true; cy = hy - qy - qh + 4; }
    cx = (int16_t)cx; cy = (int16_t)cy;
    return f;
}

// SAD of the dx x dy blocks at p0 and p1 (row stride bs) over the warp
__device__ __forceinline__ int dmvr_sad(const int16_t *p0, const int16_t *p1, int bs, int dx, int dy, int lane)
{
    int acc = 0;
    for (int idx = lane; idx < dx * dy; idx += 32) {
        const int r = idx / dx, c = idx - r * dx;
        acc += abs(p0[r * bs + c] - p1[r * bs + c]);
    }
    return __reduce_add_sync(0xffffffffu, acc);
}

// Refinement + prediction + reconstruction of one sub-PU (sx, sy, dx x dy) of a DMVR CU by one warp
__device__ void dmvr_sub_pu(const XbFrameArgs &a, const XB200_CU &cu, const int (&start)[2][2], int sx, int sy, int dx, int dy,
                            const int16_t *res_y, const int16_t *res_u, const int16_t *res_v, int rs_l, int rs_c, int ctu_x, int ctu_y,
                            int16_t *scr, int lane)
{
    const int w = 1 << cu.log2w;
    const pel *ry[2] = {a.ref_y[0][cu.refi[0]], a.ref_y[1][cu.refi[1]]};
    constexpr int IT = 2;
    const int bs = dx + 2 * IT;
    int16_t *bl0 = scr, *bl1 = scr + bs * (dy + 2 * IT);
    // bilinear search planes of this sub-PU: (dx + 4) x (dy + 4) around the start position of each list
    for (int l = 0; l < 2; l++) {
        const int gx = (((cu.x + sx) << 2) + start[l][0] - (IT << 2)) << 2, gy = (((cu.y + sy) << 2) + start[l][1] - (IT << 2)) << 2;
        int16_t *d = l ? bl1 : bl0;
        for (int idx = lane; idx < bs * (dy + 2 * IT); idx += 32) {
            const int i = idx / bs, j = idx - i * bs;
            d[idx] = (int16_t)dmvr_bilinear(ry[l], a.s_l, gx, gy, i, j, a.bd_l);
        }
    }
    __syncwarp();
    const int16_t *c0 = bl0 + IT * bs + IT, *c1 = bl1 + IT * bs + IT;
    int totx = 0, toty = 0, min_cost = 0x7fffffff, centre = 0x7fffffff, cost[5];
    bool not_zero = true;
    for (int it = 0; it < IT; it++) {
        const int16_t *a0 = c0 + totx + toty * bs, *a1 = c1 - (totx + toty * bs);
#pragma unroll
        for (int k = 0; k < 5; k++) cost[k] = 0x7fffffff;
        centre = 0x7fffffff;
        if (it == 0) min_cost = dmvr_sad(a0, a1, bs, dx, dy, lane);
        if ((it > 0 && min_cost == 0) || (it == 0 && min_cost < dx * dy)) { not_zero = false; break; }
        centre = min_cost;
        int ox[5] = {0, 0, 1, -1, 0}, oy[5] = {1, -1, 0, 0, 0}, bx = 0, by = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            cost[k] = dmvr_sad(a0 + ox[k] + oy[k] * bs, a1 - ox[k] - oy[k] * bs, bs, dx, dy, lane);
            if (k == 3) { oy[4] = cost[0] <= cost[1] ? 1 : -1; ox[4] = cost[2] <= cost[3] ? 1 : -1; }
            if (cost[k] < min_cost) { min_cost = cost[k]; bx = ox[k]; by = oy[k]; }
        }
        if (bx == 0 && by == 0) break;
        totx += bx; toty += by;
    }
    int ddx = totx << 4, ddy = toty << 4;
    if (not_zero && min_cost == centre) { ddx += dmvr_subpel(centre, cost[3], cost[2]); ddy += dmvr_subpel(centre, cost[1], cost[0]); }
    int refined[2][2];
    refined[0][0] = (start[0][0] << 2) + (int16_t)ddx; refined[0][1] = (start[0][1] << 2) + (int16_t)ddy;
    refined[1][0] = (start[1][0] << 2) - (int16_t)ddx; refined[1][1] = (start[1][1] << 2) - (int16_t)ddy;
    __syncwarp();
    // refined vectors of the sub-PU's SCUs -> map_mv (xevdm_set_dec_info publishes dmvr_mv for DMVR CUs)
    if (lane < (dx >> 2) * (dy >> 2)) {
        const int i = lane % (dx >> 2), j = lane / (dx >> 2);
        const int p = ((cu.y + sy) >> 2) * a.w_scu + ((cu.x + sx) >> 2) + j * a.w_scu + i;
        int16_t *o = a.map_mv + (size_t)p * 4;
        o[0] = (int16_t)(refined[0][0] >> 2); o[1] = (int16_t)(refined[0][1] >> 2);
        o[2] = (int16_t)(refined[1][0] >> 2); o[3] = (int16_t)(refined[1][1] >> 2);
    }
    // final prediction from the padded windows, plane by plane: both lists, average, residual, clip, store.  The loops stay rolled and
    // every interpolation routine has ONE call site: this kernel is instruction-cache bound (profiles/r1), code size is time.
    const int px = cu.x + sx, py = cu.y + sy;
    const int rpl_l = max(1, (dx * dy) >> 5), rpl_c = max(1, ((dx >> 1) * (dy >> 1)) >> 5);
    int gxv[2], gyv[2], dl[2][2], dc[2][2], wg[2][2];
#pragma unroll
    for (int l = 0; l < 2; l++) {
        int clx, cly;
        const bool clipped = dmvr_clip_one(px, py, a.w, a.h, dx, dy, refined[l][0] >> 2, refined[l][1] >> 2, clx, cly);
        wg[l][0] = ((px << 2) + start[l][0]) << 2; wg[l][1] = ((py << 2) + start[l][1]) << 2;       // window of the start vector
        if (clipped) {
            gxv[l] = (px << 4) + (clx << 2); gyv[l] = (py << 4) + (cly << 2);
            dl[l][0] = (clx >> 2) - (start[l][0] >> 2); dl[l][1] = (cly >> 2) - (start[l][1] >> 2);
            dc[l][0] = (clx >> 3) - (start[l][0] >> 3); dc[l][1] = (cly >> 3) - (start[l][1] >> 3);
        } else {
            gxv[l] = (px << 4) + refined[l][0]; gyv[l] = (py << 4) + refined[l][1];
            dl[l][0] = (refined[l][0] >> 4) - (start[l][0] >> 2); dl[l][1] = (refined[l][1] >> 4) - (start[l][1] >> 2);
            dc[l][0] = (refined[l][0] >> 5) - (start[l][0] >> 3); dc[l][1] = (refined[l][1] >> 5) - (start[l][1] >> 3);
        }
    }
    const int maxv = (1 << a.bd_l) - 1;
#pragma unroll 1
    for (int pl = 0; pl < 3; pl++) {
        const int sh = pl ? 1 : 0, bw = dx >> sh, bh = dy >> sh, rpl = pl ? rpl_c : rpl_l, s = pl ? a.s_c : a.s_l;
        const int ww = bw + (pl ? 3 : 7), wh = bh + (pl ? 3 : 7);
        int p0[8], pr[8];
#pragma unroll 1
        for (int l = 0; l < 2; l++) {
            const int ri = cu.refi[l];
            const pel *plane = pl == 0 ? a.ref_y[l][ri] : (pl == 1 ? a.ref_u[l][ri] : a.ref_v[l][ri]);
            const int wgx = l ? wg[1][0] : wg[0][0], wgy = l ? wg[1][1] : wg[0][1], gx = l ? gxv[1] : gxv[0], gy = l ? gyv[1] : gyv[0];
            const int wx = pl ? (wgx >> 5) - 1 : (wgx >> 4) - 3, wy = pl ? (wgy >> 5) - 1 : (wgy >> 4) - 3;
            const int ddxi = pl ? (l ? dc[1][0] : dc[0][0]) : (l ? dl[1][0] : dl[0][0]), ddyi = pl ? (l ? dc[1][1] : dc[0][1]) : (l ? dl[1][1] : dl[0][1]);
            auto ld = [&](int r, int c) { return (int)plane[(ptrdiff_t)(wy + min(max(ddyi + r, 0), wh - 1)) * s + wx + min(max(ddxi + c, 0), ww - 1)]; };
            if (l == 1) {
#pragma unroll
                for (int i = 0; i < 8; i++) p0[i] = pr[i];
            }
            if (pl == 0) mc_tile_ld<8>(ld, c_mc_l[a.main_tables][gx & 15], c_mc_l[a.main_tables][gy & 15], (gx & 15) != 0, (gy & 15) != 0, bw, bh, rpl, a.bd_l, scr, lane, pr);
            else mc_tile_ld<4>(ld, c_mc_c[a.main_tables][gx & 31], c_mc_c[a.main_tables][gy & 31], (gx & 31) != 0, (gy & 31) != 0, bw, bh, rpl, a.bd_c, scr, lane, pr);
        }
        // average, residual, clip, store (xevd_average_16b_no_clip + xevdm_recon)
        const int col = lane & (bw - 1), r0 = (lane >> (31 - __clz(bw))) * rpl;
        if (r0 >= bh) continue;
        const int lx = ((cu.x + sx - ctu_x) >> sh) + col, ly = ((cu.y + sy - ctu_y) >> sh) + r0;
        const int rs = pl ? rs_c : rs_l;
        const int16_t *res = (pl == 0 ? res_y : (pl == 1 ? res_u : res_v)) + ly * rs + lx;
        pel *dst = (pl == 0 ? a.cur.y : (pl == 1 ? a.cur.u : a.cur.v)) + (size_t)((ctu_y >> sh) + ly) * s + (ctu_x >> sh) + lx;
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (i < rpl) dst[(size_t)i * s] = (pel)xb_clip3(0, maxv, (int16_t)(((p0[i] + pr[i] + 1) >> 1) + res[i * rs]));
    }
    (void)w;
}

// ---- affine (Main, tool_affine) -------------------------------------------------------------------------------------------------
// xevdm_affine_mc (src_main/xevdm_mc.c:2606-2685): per list either ordinary interpolation with one vector for the whole CU (sub-blocks
// of at least 8x8; the reference evaluates the model at (sub_w/2, sub_h/2) for every sub-block, :2359-2360) or EIF - per-sample
// bilinear fetch at the model vector followed by a separable {-1, 10, -1} filter with s16 intermediates (:2425-2604).
struct AffineModel {
    int sub_w, sub_h;
    bool mem_ok;
};
__device__ __forceinline__ int aff_round(int v, int sh) { return (v + (sh > 0 ? 1 << (sh - 1) : 0) - (v >= 0)) >> sh; }
__device__ __forceinline__ void aff_gradients(const int16_t (*cp)[2], int lw, int lh, bool six, int (&dh)[2], int (&dv)[2])
{
    dh[0] = ((cp[1][0] - cp[0][0]) << 7) >> lw; dh[1] = ((cp[1][1] - cp[0][1]) << 7) >> lw;
    if (six) { dv[0] = ((cp[2][0] - cp[0][0]) << 7) >> lh; dv[1] = ((cp[2][1] - cp[0][1]) << 7) >> lh; }
    else { dv[0] = -dh[1]; dv[1] = dh[0]; }
}
// xevdm_derive_affine_subblock_size_bi + xevdm_check_eif_applicability_bi (src_main/xevdm_util.c:1870-2122)
__device__ __forceinline__ AffineModel aff_model(const XB200_CU &cu, const XB200_CU_EXT &ex)
{
    const int w = 1 << cu.log2w, h = 1 << cu.log2h;
    const bool six = (cu.flags & XB200_CUF_AFF6) != 0;
    AffineModel m;
    m.sub_w = w; m.sub_h = h; m.mem_ok = true;
    bool apply = true;
    for (int l = 0; l < 2; l++) {
        if (cu.refi[l] < 0) continue;
        int dh[2], dv[2];
        aff_gradients(ex.u.affine.cp[l], cu.log2w, cu.log2h, six, dh, dv);
        const int wx = max(abs(dh[0]), abs(dh[1])), wy = max(abs(dv[0]), abs(dv[1]));
        m.sub_w = min(m.sub_w, wx > 4 ? 4 : (wx == 0 ? w : (wx == 1 ? 32 : (wx == 2 ? 16 : 8))));
        m.sub_h = min(m.sub_h, wy > 4 ? 4 : (wy == 0 ? h : (wy == 1 ? 32 : (wy == 2 ? 16 : 8))));
    }
    for (int l = 0; l < 2 && apply; l++) {
        if (cu.refi[l] < 0) continue;
        int dh[2], dv[2];
        aff_gradients(ex.u.affine.cp[l], cu.log2w, cu.log2h, six, dh, dv);
        // fetch area of a 4x4 block (calculate_bounding_box_size) against MAX_MEMORY_ACCESS_BI = 72
        const int x1 = 5 * (dh[0] + 512), x2 = 5 * dv[0], y1 = 5 * dh[1], y2 = 5 * (dv[1] + 512);
        const int bw = ((max(max(0, x1), max(x2, x1 + x2)) - min(min(0, x1), min(x2, x1 + x2)) + 511) >> 9) + 2;
        const int bh = ((max(max(0, y1), max(y2, y1 + y2)) - min(min(0, y1), min(y2, y1 + y2)) + 511) >> 9) + 2;
        if (dv[1] < -512 || (max(0, dv[1]) + abs(dh[1])) * 5 > 512) apply = false;
        m.mem_ok = m.mem_ok && (bw * bh <= 72);
    }
    if (!apply) { m.sub_w = max(m.sub_w, 8); m.sub_h = max(m.sub_h, 8); }
    return m;
}
// vector xevdm_set_affine_mvf stores for SCU (sx, sy) of the CU, list l (src_main/xevdm_util.c:4095-4203)
__device__ __forceinline__ int aff_map_mv(const XB200_CU &cu, const XB200_CU_EXT &ex, const AffineModel &m, int l, int sx, int sy)
{
    const int16_t (*v)[2] = ex.u.affine.cp[l];
    const bool six = (cu.flags & XB200_CUF_AFF6) != 0;
    const int wc = 1 << (cu.log2w - 2), hc = 1 << (cu.log2h - 2), sws = m.sub_w >> 2, shs = m.sub_h >> 2;
    const int bx = sx - sx % sws, by = sy - sy % shs;                  // first SCU of the sub-block
    int mx, my;
    if (bx == 0 && by == 0) { mx = v[0][0]; my = v[0][1]; }
    else if (bx + sws == wc && by == 0) { mx = v[1][0]; my = v[1][1]; }
    else if (bx == 0 && by + shs == hc && six) { mx = v[2][0]; my = v[2][1]; }
    else {
        const int dhx = (v[1][0] - v[0][0]) << (7 - cu.log2w), dhy = (v[1][1] - v[0][1]) << (7 - cu.log2w);
        const int dvx = six ? (v[2][0] - v[0][0]) << (7 - cu.log2h) : -dhy, dvy = six ? (v[2][1] - v[0][1]) << (7 - cu.log2h) : dhx;
        const int px = (bx << 2) + (m.sub_w >> 1), py = (by << 2) + (m.sub_h >> 1);
        mx = xb_clip3(-(1 << 17), (1 << 17) - 1, aff_round((v[0][0] << 7) + dhx * px + dvx * py, 5)) >> 2;
        my = xb_clip3(-(1 << 17), (1 << 17) - 1, aff_round((v[0][1] << 7) + dhy * px + dvy * py, 5)) >> 2;
    }
    return (mx & 0xffff) | (my << 16);
}

// one <=16x16 luma tile (tx, ty) of an affine CU by one warp: both lists, three planes, average, residual, store
__device__ void affine_tile(const XbFrameArgs &a, const XB200_CU &cu, const XB200_CU_EXT &ex, const AffineModel &m, int tx, int ty, int tw, int th,
                            const int16_t *res_y, const int16_t *res_u, const int16_t *res_v, int rs_l, int rs_c, int ctu_x, int ctu_y,
                            int16_t *scr, int lane)
{
    const int w = 1 << cu.log2w, h = 1 << cu.log2h;
    const bool six = (cu.flags & XB200_CUF_AFF6) != 0, eif = m.sub_w < 8 || m.sub_h < 8;
    const int rpl_l = max(1, (tw * th) >> 5), rpl_c = max(1, ((tw >> 1) * (th >> 1)) >> 5);
    const int maxv = (1 << a.bd_l) - 1;
    // plane by plane, lists inside, loops rolled, one call site per interpolation routine (instruction-cache bound kernel: code size is time)
#pragma unroll 1
    for (int pl = 0; pl < 3; pl++) {
        const int sh = pl ? 1 : 0, bw = tw >> sh, bh = th >> sh, rpl = pl ? rpl_c : rpl_l, bd = pl ? a.bd_c : a.bd_l, s = pl ? a.s_c : a.s_l;
        int acc[8];
        int nl = 0;
#pragma unroll 1
        for (int l = 0; l < 2; l++) {
            if (cu.refi[l] < 0) continue;
            const int ri = cu.refi[l];
            int dh[2], dv[2];
            aff_gradients(ex.u.affine.cp[l], cu.log2w, cu.log2h, six, dh, dv);
            const int sc[2] = {ex.u.affine.cp[l][0][0] << 7, ex.u.affine.cp[l][0][1] << 7};
            const pel *plane = pl == 0 ? a.ref_y[l][ri] : (pl == 1 ? a.ref_u[l][ri] : a.ref_v[l][ri]);
            int pr[8];
            if (!eif) {
                int mvo[2], mvc[2];
#pragma unroll
                for (int c = 0; c < 2; c++) mvo[c] = xb_clip3(-(1 << 17), (1 << 17) - 1, aff_round(sc[c] + dh[c] * (m.sub_w >> 1) + dv[c] * (m.sub_h >> 1), 5));
                mvc[0] = min((a.w + 128 - cu.x - w) << 4, max((-128 - cu.x) << 4, mvo[0]));
                mvc[1] = min((a.h + 128 - cu.y - h) << 4, max((-128 - cu.y) << 4, mvo[1]));
                const int gx = ((cu.x + tx) << 4) + mvc[0], gy = ((cu.y + ty) << 4) + mvc[1];
                if (pl == 0) {
                    const pel *ref = plane + (ptrdiff_t)(gy >> 4) * s + (gx >> 4);
                    mc_tile<8>(ref, s, c_mc_l[a.main_tables][gx & 15], c_mc_l[a.main_tables][gy & 15], (mvo[0] & 15) != 0, (mvo[1] & 15) != 0, bw, bh, rpl, bd, scr, lane, pr);
                } else {
                    const pel *ref = plane + (ptrdiff_t)(gy >> 5) * s + (gx >> 5);
                    mc_tile<4>(ref, s, c_mc_c[a.main_tables][gx & 31], c_mc_c[a.main_tables][gy & 31], (mvo[0] & 31) != 0, (mvo[1] & 31) != 0, bw, bh, rpl, bd, scr, lane, pr);
                }
            } else {
                // eif_derive_mv_clip_range (xevdm_mc.c:2108-2150), 1/32 sample
                int mxv[2], mnv[2];
                const int pmx[2] = {(a.w + 128 - cu.x - w - 1) << 5, (a.h + 128 - cu.y - h - 1) << 5}, pmn[2] = {(-cu.x - 128) << 5, (-cu.y - 128) << 5};
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    if (m.mem_ok) { mxv[c] = pmx[c]; mnv[c] = pmn[c]; }
                    else {
                        const int centre = aff_round(sc[c] + dh[c] * (w >> 1) + dv[c] * (h >> 1), 4);
                        const int lg = (c == 0 ? cu.log2w : cu.log2h) - 3;
                        const int spread = lg == 0 ? 128 : (lg == 1 ? 256 : (lg == 2 ? 544 : (lg == 3 ? 1120 : 2272)));
                        mnv[c] = centre - spread; mxv[c] = centre + spread;
                        if (mnv[c] < pmn[c]) { mnv[c] = pmn[c]; mxv[c] = min(pmx[c], pmn[c] + 2 * spread); }
                        else if (mxv[c] > pmx[c]) { mxv[c] = pmx[c]; mnv[c] = max(pmn[c], pmx[c] - 2 * spread); }
                    }
                    mxv[c] = xb_clip3(-(1 << 17), (1 << 17) - 1, mxv[c]);
                    mnv[c] = xb_clip3(-(1 << 17), (1 << 17) - 1, mnv[c]);
                }
                // bilinear samples of the (bw + 2) x (bh + 2) neighbourhood of the tile, positions relative to the CU plane
                const int mv0x = sc[0] >> sh, mv0y = sc[1] >> sh, lim_x0 = mnv[0] >> sh, lim_x1 = mxv[0] >> sh, lim_y0 = mnv[1] >> sh, lim_y1 = mxv[1] >> sh;
                const int ox = (cu.x >> sh), oy = (cu.y >> sh), px0 = tx >> sh, py0 = ty >> sh;
                const int s1 = min(4, bd - 8), s2 = max(8, 20 - bd), sh_h = max(bd + 5 - 16, 0), sh_v = 6 - sh_h;
                const int st = bw + 2;
                int16_t *bb = scr, *hb = scr + 18 * 18;
                for (int idx = lane; idx < st * (bh + 2); idx += 32) {
                    const int jj = idx / st, j = jj - 1 + py0, i = idx - jj * st - 1 + px0;          // sample position inside the CU plane
                    const int vx = xb_clip3(lim_x0, lim_x1, (mv0x + i * dh[0] + j * dv[0]) >> 4), vy = xb_clip3(lim_y0, lim_y1, (mv0y + i * dh[1] + j * dv[1]) >> 4);
                    const pel *r = plane + (ptrdiff_t)(oy + j + (vy >> 5)) * s + ox + i + (vx >> 5);
                    const int fx = vx & 31, fy = vy & 31;
                    const int a0 = (int16_t)(((64 - 2 * fx) * r[0] + 2 * fx * r[1]) >> s1), a1 = (int16_t)(((64 - 2 * fx) * r[s] + 2 * fx * r[s + 1]) >> s1);
                    bb[idx] = (int16_t)(((64 - 2 * fy) * a0 + 2 * fy * a1 + (1 << (s2 - 1))) >> s2);
                }
                __syncwarp();
                const int lbw = 31 - __clz(bw);
                for (int idx = lane; idx < bw * (bh + 2); idx += 32) {
                    const int j = idx >> lbw, i = idx & (bw - 1);
                    hb[idx] = (int16_t)((-bb[j * st + i] + 10 * bb[j * st + i + 1] - bb[j * st + i + 2] + (sh_h ? 1 << (sh_h - 1) : 0)) >> sh_h);
                }
                __syncwarp();
                const int col = lane & (bw - 1), r0 = (lane >> lbw) * rpl;
#pragma unroll
                for (int i = 0; i < 8; i++)
                    if (i < rpl && r0 < bh) {
                        const int v = (int16_t)((-hb[(r0 + i) * bw + col] + 10 * hb[(r0 + i + 1) * bw + col] - hb[(r0 + i + 2) * bw + col] + (1 << (sh_v - 1))) >> sh_v);
                        pr[i] = xb_clip3(0, (1 << bd) - 1, v);
                    }
                __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = nl ? (acc[i] + pr[i] + 1) >> 1 : pr[i];
            nl++;
        }
        const int col = lane & (bw - 1), r0 = (lane >> (31 - __clz(bw))) * rpl;
        if (r0 >= bh) continue;
        const int lx = ((cu.x + tx - ctu_x) >> sh) + col, ly = ((cu.y + ty - ctu_y) >> sh) + r0;
        const int rs = pl ? rs_c : rs_l;
        const int16_t *res = (pl == 0 ? res_y : (pl == 1 ? res_u : res_v)) + ly * rs + lx;
        pel *dst = (pl == 0 ? a.cur.y : (pl == 1 ? a.cur.u : a.cur.v)) + (size_t)((ctu_y >> sh) + ly) * s + (ctu_x >> sh) + lx;
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (i < rpl) dst[(size_t)i * s] = (pel)xb_clip3(0, maxv, (int16_t)(acc[i] + res[i * rs]));
    }
}

// ---- residual phase helpers -------------------------------------------------------------------------------
// The coded transform block of plane `pl` covering SCU (xs, ys) (CTU-relative SCU coordinates) of a CU, or false when that SCU
// carries no coefficients.  Normal CUs: the CU plane cut into <= 64-sample (chroma 32) blocks gated by the nnz_sub bits
// (xevd_sub_block_itdq, src_base/xevd_itdq.c:544-621).  ats_inter CUs (Main): one sub-block TU at the side the syntax names
// (xevdm_get_tu_size / get_tu_pos_offset, src_main/xevdm_util.c:3585-3634); their coefficient blocks hold the TU only.
__device__ __forceinline__ int plane_coef_base(const XB200_CU &cu, int pl)
{
    const int n = 1 << (cu.log2w + cu.log2h);
    int off = cu.coef_off;
    if (pl >= 1 && (cu.cbf & 0x00f)) off += (n + 7) & ~7;
    if (pl == 2 && (cu.cbf & 0x0f0)) off += ((n >> 2) + 7) & ~7;
    return off;
}

struct TbInfo {
    int lw, lh;         // log2 size of the transform block (plane samples)
    int px0, py0;       // origin inside the CTU plane (plane samples)
    int sx0, sy0;       // first SCU column / row (CTU-relative)
    int coef;           // offset of the block's top-left coefficient inside the stream
    int cstride;        // coefficient row stride
    int qp;
    int ats;            // -1: DCT-2; else horizontal << 1 | vertical, 0 = DST-7, 1 = DCT-8
};

__device__ __forceinline__ int ats_inter_idx(const XB200_CU &cu) { return (cu.mode == XB200_MODE_INTRA || cu.mode == XB200_MODE_IBC) ? 0 : XB200_ATS_INTER_IDX(cu.ats); }

// log2 TU size and offset (luma samples) of an ats_inter CU
__device__ __forceinline__ void ats_inter_tu(const XB200_CU &cu, int idx, int &tlw, int &tlh, int &xo, int &yo)
{
    const int pos = XB200_ATS_INTER_POS(cu.ats), sh = (idx >= 3) ? 2 : 1;
    tlw = cu.log2w; tlh = cu.log2h; xo = 0; yo = 0;
    if (idx == 2 || idx == 4) { tlh -= sh; yo = pos ? (1 << cu.log2h) - (1 << tlh) : 0; }
    else { tlw -= sh; xo = pos ? (1 << cu.log2w) - (1 << tlw) : 0; }
}

__device__ __forceinline__ bool tb_at(const XbFrameArgs &a, const XB200_CU &cu, int pl, int ctu_x, int ctu_y, int xs, int ys, TbInfo &t)
{
    const int bits = (cu.cbf >> (4 * pl)) & 15;
    if (!bits) return false;
    const int sh = pl ? 1 : 0;
    const int cx_scu = (cu.x - ctu_x) >> 2, cy_scu = (cu.y - ctu_y) >> 2;
    const int rx = xs - cx_scu, ry = ys - cy_scu;            // SCU position inside the CU
    t.qp = pl == 0 ? cu.qp_y : (pl == 1 ? cu.qp_u : cu.qp_v);
    t.ats = -1;
    const int idx = a.ats ? ats_inter_idx(cu) : 0;
    if (idx) {
        int tlw, tlh, xo, yo;
        ats_inter_tu(cu, idx, tlw, tlh, xo, yo);
        if ((rx << 2) < xo || (rx << 2) >= xo + (1 << tlw) || (ry << 2) < yo || (ry << 2) >= yo + (1 << tlh)) return false;
        t.lw = tlw - sh; t.lh = tlh - sh;
        t.px0 = ((cu.x - ctu_x) + xo) >> sh; t.py0 = ((cu.y - ctu_y) + yo) >> sh;
        t.sx0 = cx_scu + (xo >> 2); t.sy0 = cy_scu + (yo >> 2);
        int off = cu.coef_off;
        const int n = 1 << (tlw + tlh);
        if (pl >= 1 && (cu.cbf & 0x00f)) off += (n + 7) & ~7;
        if (pl == 2 && (cu.cbf & 0x0f0)) off += ((n >> 2) + 7) & ~7;
        t.coef = off; t.cstride = 1 << t.lw;
        if (pl == 0 && cu.log2w <= 5 && cu.log2h <= 5) {       // xevdm_get_ats_inter_trs (xevdm_util.c:3636-3667)
            const int first = XB200_ATS_INTER_POS(cu.ats) == 0 ? 1 : 0;
            t.ats = (idx == 2 || idx == 4) ? first : (first << 1);
        }
        return true;
    }
    const int lws = min((int)cu.log2w, 6) - 2, lhs = min((int)cu.log2h, 6) - 2;     // log2 block size in SCUs
    const int sub_i = rx >> lws, sub_j = ry >> lhs;
    if (!((bits >> ((sub_j << 1) | sub_i)) & 1)) return false;
    t.lw = lws + 2 - sh; t.lh = lhs + 2 - sh;
    t.sx0 = cx_scu + (sub_i << lws); t.sy0 = cy_scu + (sub_j << lhs);
    t.px0 = (t.sx0 << 2) >> sh; t.py0 = (t.sy0 << 2) >> sh;
    const int pw = (1 << cu.log2w) >> sh;
    t.cstride = pw;
    t.coef = plane_coef_base(cu, pl) + (((sub_j << lhs) << 2) >> sh) * pw + (((sub_i << lws) << 2) >> sh);
    if (a.ats && pl == 0 && cu.mode == XB200_MODE_INTRA && (cu.flags & XB200_CUF_ATS_INTRA)) t.ats = cu.ats & 3;
    return true;
}

template <bool IQT>
__global__ void __launch_bounds__(kReconThreads, 2)
k_recon_inter(const __grid_constant__ XbFrameArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ReconSmem sm;
    sm.carve(smem_raw, a.log2_ctu);
    const int S = sm.S, Sc = sm.Sc, nscu = S >> 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctu = blockIdx.x;
    const int ctu_x = (ctu % a.w_ctu) << a.log2_ctu, ctu_y = (ctu / a.w_ctu + a.ctu_row0) << a.log2_ctu;
    const int cu0 = a.ctu_first[ctu], cu1 = a.ctu_first[ctu + 1];
    const XB200_CU *cus = a.cus + cu0;
    const int ncu = cu1 - cu0;
    if (a.dispatch && !ctu_needs_generic(a, cus, ncu, tid, kReconThreads)) {            // no ATS / DMVR / affine CU here: the throughput kernel does it all
        if (a.inter_done && tid == 0) { atomicExch(a.inter_done + ctu, 1); atomicAdd(a.inter_count, 1); }
        return;
    }

    // ---- SCU -> CU map, zero residual -------------------------------------------------------------------
    for (int i = tid; i < 2 * nscu * nscu; i += kReconThreads) sm.cu_of_scu[i] = 0xffff;          // both owner tables (contiguous)
    for (int i = tid; i < S * (S + 2) / 2; i += kReconThreads) ((int *)sm.res_y)[i] = 0;
    for (int i = tid; i < Sc * (Sc + 2); i += kReconThreads) ((int *)sm.res_u)[i] = 0;     // res_u and res_v are contiguous
    __syncthreads();
    for (int i = tid; i < ncu; i += kReconThreads) {
        const XB200_CU cu = cus[i];
        if (a.dispatch && !cu_needs_generic(a, cu)) continue;        // per-CU dispatch: the throughput kernel reconstructs this CU (and publishes its maps)
        const int sx = (cu.x - ctu_x) >> 2, sy = (cu.y - ctu_y) >> 2;
        const int nw = 1 << (cu.log2w - 2), nh = 1 << (cu.log2h - 2);
        for (int y = 0; y < nh; y++)
            for (int x = 0; x < nw; x++) {
                // TREE_L leaves own luma and the maps, the TREE_C CU that follows them the chroma of the whole node (xevdm.c:1828-1846)
                if (cu.flags & XB200_CUF_LUMA) sm.cu_of_scu[(sy + y) * nscu + sx + x] = (uint16_t)i;
                if (cu.flags & XB200_CUF_CHROMA) sm.cu_of_scu_c[(sy + y) * nscu + sx + x] = (uint16_t)i;
            }
    }
    __syncthreads();

    // ---- phase A1: column transforms -------------------------------------------------------------------------
    // slot = (plane column, SCU row); a slot is live when a coded transform block starts at that SCU row
    {
        const int sh1 = IQT ? 7 : 0;
        const int luma_slots = S * nscu, chroma_slots = Sc * nscu;
        for (int s = tid; s < luma_slots + 2 * chroma_slots; s += kReconThreads) {
            int pl, x, ys;
            if (s < luma_slots) { pl = 0; ys = s >> a.log2_ctu; x = s - ys * S; }
            else { int t = s - luma_slots; pl = 1 + (t >= chroma_slots); t -= (pl - 1) * chroma_slots; ys = t >> (a.log2_ctu - 1); x = t - ys * Sc; }
            const int xs = pl == 0 ? (x >> 2) : (x >> 1);
            const unsigned ci = (pl ? sm.cu_of_scu_c : sm.cu_of_scu)[ys * nscu + xs];
            if (ci == 0xffff) continue;
            const XB200_CU cu = cus[ci];
            TbInfo t;
            if (!tb_at(a, cu, pl, ctu_x, ctu_y, xs, ys, t) || t.sy0 != ys) continue;
            const int16_t *src = a.coef + t.coef + (x - t.px0);
            Dequant dq;
            dq.init(t.lw, t.lh, t.qp, a.bd_l, IQT);
            int *tmp = pl == 0 ? sm.tmp_y : (pl == 1 ? sm.tmp_u : sm.tmp_v);
            const int ts = (pl == 0 ? S : Sc) + 1, cs = t.cstride;
            int *dst = tmp + t.py0 * ts + x;
            if (t.ats >= 0) ats_line_dyn(t.lh, t.ats & 1, [&](int k) { return dq.apply(src[k * cs]); }, [&](int n, int v) { dst[n * ts] = v; }, 7);
            else itx_line_dyn<IQT>(t.lh, [&](int k) { return dq.apply(src[k * cs]); }, [&](int n, int v) { dst[n * ts] = v; }, sh1);
        }
    }
    __syncthreads();
    // ---- phase A2: row transforms ------------------------------------------------------------------------------
    {
        const int sh2 = IQT ? 12 - (a.bd_l - 8) : 19 - (a.bd_l - 8);
        const int luma_slots = S * nscu, chroma_slots = Sc * nscu;
        for (int s = tid; s < luma_slots + 2 * chroma_slots; s += kReconThreads) {
            int pl, y, xs;
            if (s < luma_slots) { pl = 0; xs = s >> a.log2_ctu; y = s - xs * S; }
            else { int t = s - luma_slots; pl = 1 + (t >= chroma_slots); t -= (pl - 1) * chroma_slots; xs = t >> (a.log2_ctu - 1); y = t - xs * Sc; }
            const int ys = pl == 0 ? (y >> 2) : (y >> 1);
            const unsigned ci = (pl ? sm.cu_of_scu_c : sm.cu_of_scu)[ys * nscu + xs];
            if (ci == 0xffff) continue;
            const XB200_CU cu = cus[ci];
            TbInfo t;
            if (!tb_at(a, cu, pl, ctu_x, ctu_y, xs, ys, t) || t.sx0 != xs) continue;
            const int *tmp = pl == 0 ? sm.tmp_y : (pl == 1 ? sm.tmp_u : sm.tmp_v);
            int16_t *res = pl == 0 ? sm.res_y : (pl == 1 ? sm.res_u : sm.res_v);
            const int ts = (pl == 0 ? S : Sc) + 1, rs = (pl == 0 ? S : Sc) + 2;
            const int *srow = tmp + y * ts + t.px0;
            int16_t *drow = res + y * rs + t.px0;
            if (t.ats >= 0) ats_line_dyn(t.lw, t.ats >> 1, [&](int k) { return srow[k]; }, [&](int n, int v) { drow[n] = (int16_t)v; }, 20 - a.bd_l);
            else itx_line_dyn<false>(t.lw, [&](int k) { return srow[k]; }, [&](int n, int v) { drow[n] = (int16_t)v; }, sh2);
        }
    }
    __syncthreads();

    // ---- phase B: prediction + reconstruction, warp per 16x16 luma tile ----------------------------------------------
    {
        int16_t *scr = sm.mc + warp * kMcScratchPerWarp;
        const int tiles = S >> 4;
        for (int t = warp; t < tiles * tiles; t += kReconWarps) {
            const int t_x = (t & (tiles - 1)) << 4, t_y = (t >> (a.log2_ctu - 4)) << 4;        // tile origin inside the CTU (luma)
            if (ctu_x + t_x >= a.w || ctu_y + t_y >= a.h) continue;
            // walk the (up to 16) CUs that start inside this tile, or the single CU that covers it
            for (int sy = 0; sy < 4; sy++) {
                for (int sx = 0; sx < 4; sx++) {
                    const unsigned ci = sm.cu_of_scu[((t_y >> 2) + sy) * nscu + (t_x >> 2) + sx];
                    if (ci == 0xffff) continue;
                    const XB200_CU cu = cus[ci];
                    if (xb_wavefront_mode(cu.mode)) continue;
                    if (a.dmvr) { int st[2][2]; if (dmvr_applies(a, cu, st)) continue; }      // refined CUs: phase B2
                    if (cu.mode == XB200_MODE_AFFINE) continue;                               // affine CUs: phase B3
                    const int cx = cu.x - ctu_x, cy = cu.y - ctu_y;
                    // piece of the CU inside this tile; handled when this SCU is the piece's top-left
                    const int px = max(cx, t_x), py = max(cy, t_y);
                    if (px != t_x + (sx << 2) || py != t_y + (sy << 2)) continue;
                    // the piece ends at the CU's or the tile's edge, whichever comes first: the middle part of a ternary split
                    // (size s at offset s/2) is not aligned to its own size, so a 16-wide CU can straddle two tiles
                    const int tw = min(cx + (1 << cu.log2w), t_x + 16) - px, th = min(cy + (1 << cu.log2h), t_y + 16) - py;
                    int pr[8];
                    // luma
                    {
                        const int rpl = max(1, (tw * th) >> 5);
                        pred_tile<8>(a, cu, 0, px - cx, py - cy, tw, th, rpl, scr, lane, pr);
                        const int col = lane & (tw - 1), r0 = (lane >> (31 - __clz(tw))) * rpl;
                        if (r0 < th) {
                            pel *dst = a.cur.y + (ctu_y + py + r0) * a.s_l + ctu_x + px + col;
                            const int16_t *res = sm.res_y + (py + r0) * (S + 2) + px + col;
                            const int maxv = (1 << a.bd_l) - 1;
#pragma unroll
                            for (int i = 0; i < 8; i++)
                                if (i < rpl) dst[i * a.s_l] = (pel)xb_clip3(0, maxv, (int16_t)(pr[i] + res[i * (S + 2)]));
                        }
                    }
                    // chroma (4:2:0): both planes
                    {
                        const int cw = tw >> 1, ch = th >> 1;
                        const int rpl = max(1, (cw * ch) >> 5);
#pragma unroll 1
                        for (int pl = 1; pl <= 2; pl++) {
                            pred_tile<4>(a, cu, pl, (px - cx) >> 1, (py - cy) >> 1, cw, ch, rpl, scr, lane, pr);
                            const int col = lane & (cw - 1), r0 = (lane >> (31 - __clz(cw))) * rpl;
                            if (r0 < ch) {
                                pel *dst = (pl == 1 ? a.cur.u : a.cur.v) + (((ctu_y + py) >> 1) + r0) * a.s_c + ((ctu_x + px) >> 1) + col;
                                const int16_t *res = (pl == 1 ? sm.res_u : sm.res_v) + ((py >> 1) + r0) * (Sc + 2) + (px >> 1) + col;
                                const int maxv = (1 << a.bd_l) - 1;      // the reference clips chroma with the luma depth (xevd_recon.c:70-91)
#pragma unroll
                                for (int i = 0; i < 8; i++)
                                    if (i < rpl) dst[i * a.s_c] = (pel)xb_clip3(0, maxv, (int16_t)(pr[i] + res[i * (Sc + 2)]));
                            }
                        }
                    }
                }
            }
        }
    }

    // ---- phase B1: intra / IBC CUs are predicted by the wavefront kernel; their residual (neighbour-independent, transformed above with
    //      everything else) is parked in the picture, where that kernel picks it up with its CTU preload
    bool any_wf = false;
    for (int i = tid; i < ncu; i += kReconThreads) any_wf |= xb_wavefront_mode(cus[i].mode) && (!a.dispatch || cu_needs_generic(a, cus[i]));
    any_wf = __syncthreads_or(any_wf);
    for (int i = tid; any_wf && i < S * S + 2 * Sc * Sc; i += kReconThreads) {
        int pl, x, y;
        if (i < S * S) { pl = 0; y = i >> a.log2_ctu; x = i - y * S; }
        else { int t = i - S * S; pl = 1 + (t >= Sc * Sc); t -= (pl - 1) * Sc * Sc; y = t >> (a.log2_ctu - 1); x = t - y * Sc; }
        const int sh = pl ? 1 : 0;
        const unsigned ci = (pl ? sm.cu_of_scu_c : sm.cu_of_scu)[((y << sh) >> 2) * nscu + ((x << sh) >> 2)];
        if (ci == 0xffff || !xb_wavefront_mode(cus[ci].mode)) continue;
        const int16_t *res = pl == 0 ? sm.res_y : (pl == 1 ? sm.res_u : sm.res_v);
        pel *dst = (pl == 0 ? a.cur.y : (pl == 1 ? a.cur.u : a.cur.v)) + (size_t)((ctu_y >> sh) + y) * (pl ? a.s_c : a.s_l) + (ctu_x >> sh) + x;
        *dst = res[y * ((pl ? Sc : S) + 2) + x];
    }

    // ---- phase B2: DMVR CUs, one warp per 16x16 sub-PU ---------------------------------------------------------------------------
    if (a.dmvr) {
        int16_t *scr = sm.mc + warp * kMcScratchPerWarp;
        int k = 0;
        for (int i = 0; i < ncu; i++) {
            const XB200_CU cu = cus[i];
            int st[2][2];
            if (!dmvr_applies(a, cu, st)) continue;
            const int w = 1 << cu.log2w, h = 1 << cu.log2h, dx = min(w, 16), dy = min(h, 16);
            for (int sy = 0; sy < h; sy += dy)
                for (int sx = 0; sx < w; sx += dx, k++)
                    if ((k & (kReconWarps - 1)) == warp)
                        dmvr_sub_pu(a, cu, st, sx, sy, dx, dy, sm.res_y, sm.res_u, sm.res_v, S + 2, Sc + 2, ctu_x, ctu_y, scr, lane);
        }
        __syncthreads();        // refined vectors are in map_mv before phase C decides what to publish
    }

    // ---- phase B3: affine CUs, one warp per 16x16 tile -----------------------------------------------------------------------------
    if (a.affine) {
        int16_t *scr = sm.mc + warp * kMcScratchPerWarp;
        int k = 0;
        for (int i = 0; i < ncu; i++) {
            const XB200_CU cu = cus[i];
            if (cu.mode != XB200_MODE_AFFINE) continue;
            uint32_t ei;
            memcpy(&ei, cu.mv[1], 4);
            const XB200_CU_EXT ex = a.ext[ei];
            const AffineModel m = aff_model(cu, ex);
            const int w = 1 << cu.log2w, h = 1 << cu.log2h, tw = min(w, 16), th = min(h, 16);
            for (int ty = 0; ty < h; ty += th)
                for (int tx = 0; tx < w; tx += tw, k++)
                    if ((k & (kReconWarps - 1)) == warp)
                        affine_tile(a, cu, ex, m, tx, ty, tw, th, sm.res_y, sm.res_u, sm.res_v, S + 2, Sc + 2, ctu_x, ctu_y, scr, lane);
        }
    }

    // ---- phase C: publish per-SCU maps (xevd_set_dec_info) ------------------------------------------------------------
    for (int i = tid; i < nscu * nscu; i += kReconThreads) {
        const unsigned ci = sm.cu_of_scu[i];
        if (ci == 0xffff) continue;
        const XB200_CU cu = cus[ci];
        const int gx = (ctu_x >> 2) + (i & (nscu - 1)), gy = (ctu_y >> 2) + (i >> (a.log2_ctu - 2));
        const int p = gy * a.w_scu + gx;
        const bool intra = cu.mode == XB200_MODE_INTRA, ibc = cu.mode == XB200_MODE_IBC;
        uint32_t m = ((uint32_t)(cu.qp_map & 0x7f) << 16) | (1u << 31) | (intra ? 1u << 15 : 0u) | (ibc ? 1u << 26 : 0u);     // MCU_SET_IBC
        bool cbfl = (cu.cbf & 1) != 0;
        const int aidx = a.ats ? ats_inter_idx(cu) : 0;
        if (aidx && cbfl) {          // xevdm_set_cu_cbf_flags (xevdm_util.c:3669-3714): luma cbf only on the SCUs of the sub-block TU
            int tlw, tlh, xo, yo;
            ats_inter_tu(cu, aidx, tlw, tlh, xo, yo);
            const int rx = (gx << 2) - cu.x, ry = (gy << 2) - cu.y;
            cbfl = rx >= xo && rx < xo + (1 << tlw) && ry >= yo && ry < yo + (1 << tlh);
        }
        if (cbfl) m |= 1u << 24;
        if (cu.flags & XB200_CUF_SKIP) m |= 1u << 23;
        bool refined = false;
        if (a.dmvr) { int st[2][2]; refined = dmvr_applies(a, cu, st); }
        if (refined) m |= 1u << 25;                                       // MCU_SET_DMVRF; map_mv already holds the refined vectors
        int2 mvw = intra ? make_int2(0, 0) : make_int2(((const int *)cu.mv)[0], ((const int *)cu.mv)[1]);
        int2 mvm = mvw;
        if (cu.mode == XB200_MODE_AFFINE) {
            // xevdm_set_dec_info publishes core->mv (extension record); xevdm_set_affine_mvf then overwrites map_mv of the lists in use
            uint32_t ei;
            memcpy(&ei, cu.mv[1], 4);
            const XB200_CU_EXT ex = a.ext[ei];
            const AffineModel am = aff_model(cu, ex);
            mvw = make_int2(((const int *)ex.u.affine.mv_unref)[0], ((const int *)ex.u.affine.mv_unref)[1]);
            mvm = mvw;
            const int sx = gx - (cu.x >> 2), sy = gy - (cu.y >> 2);
            if (cu.refi[0] >= 0) mvm.x = aff_map_mv(cu, ex, am, 0, sx, sy);
            if (cu.refi[1] >= 0) mvm.y = aff_map_mv(cu, ex, am, 1, sx, sy);
            m |= ((cu.flags & XB200_CUF_AFF6) ? 2u : 1u) << 8;             // MCU_SET_AFF
        }
        a.map_scu[p] = m;
        if (!refined) ((int2 *)a.map_mv)[p] = mvm;
        ((int2 *)a.map_unrefined_mv)[p] = mvw;
        ((int16_t *)a.map_refi)[p] = (intra || ibc) ? (int16_t)-1 : *(const int16_t *)cu.refi;
        // a TREE_L leaf: its edges are luma edges; the chroma outline of the node is restored below by the TREE_C CU
        const bool e_l = (((gx << 2) - cu.x) & 63) == 0, e_t = (((gy << 2) - cu.y) & 63) == 0, lonly = !(cu.flags & XB200_CUF_CHROMA);
        a.map_edge[p] = (uint8_t)((e_l ? XB200_EDGE_LEFT | (lonly ? XB200_EDGE_LEFT_NOC : 0) : 0) | (e_t ? XB200_EDGE_TOP | (lonly ? XB200_EDGE_TOP_NOC : 0) : 0) |
                                  (aidx ? XB200_EDGE_ATS : 0));
        if (a.map_order) a.map_order[p] = sm.cu_of_scu_c[i];        // decoding order of the CU whose visit filters the chroma edges here
    }
    // chroma-only CUs (deblock_tree visits the node once more as TREE_C, xevdm.c:1991-1998): chroma edges along their left column / top row
    bool any_c = false;
    for (int i = tid; i < ncu; i += kReconThreads) any_c |= (cus[i].flags & (XB200_CUF_LUMA | XB200_CUF_CHROMA)) == XB200_CUF_CHROMA;
    if (__syncthreads_or(any_c)) {          // the barrier also orders the map_edge stores above before the updates below
        for (int i = tid; i < ncu; i += kReconThreads) {
            const XB200_CU cu = cus[i];
            if ((cu.flags & (XB200_CUF_LUMA | XB200_CUF_CHROMA)) != XB200_CUF_CHROMA) continue;
            const int gx = cu.x >> 2, gy = cu.y >> 2, nw = 1 << (cu.log2w - 2), nh = 1 << (cu.log2h - 2);
            for (int y = 0; y < nh; y++) a.map_edge[(gy + y) * a.w_scu + gx] &= (uint8_t)~XB200_EDGE_LEFT_NOC;
            for (int x = 0; x < nw; x++) a.map_edge[gy * a.w_scu + gx + x] &= (uint8_t)~XB200_EDGE_TOP_NOC;
        }
    }
    if (a.inter_done) {         // samples and maps of this CTU are final as far as the inter kernels go: release the wavefront kernel's waiters
        __threadfence();
        __syncthreads();
        if (tid == 0) { atomicExch(a.inter_done + ctu, 1); atomicAdd(a.inter_count, 1); }
    }
}

}  // namespace xb
