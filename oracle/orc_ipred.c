/*
 * Intra prediction, Baseline profile (5 modes).  TEST INFRASTRUCTURE ONLY (orc_common.h).
 * Restates xevd_get_nbr_b (src_base/xevd_ipred.c:33-93) and xevd_ipred_b / xevd_ipred_uv_b with their five
 * predictors (src_base/xevd_ipred.c:95-160, 586-676).
 *
 * Neighbour availability in the reference is "already reconstructed in decoding order" (COD bit of map_scu, plus tile /
 * constrained-intra tests); it is carried here as the per-SCU bit masks of XB200_CU_EXT, which the producer
 * of the CU array derives.  Unavailable units read as 1 << (bit_depth - 1), where bit_depth is the LUMA depth for all
 * three planes (src_base/xevd.c:445-475).
 */
#include <string.h>
#include <stdlib.h>
#include "orc_common.h"

/* gather up[-1 .. w+h) and left[-1 .. h+w): unit = 4 luma / 2 chroma samples per SCU */
void orc_intra_neighbours(const pel *rec, int s, int w, int h, int unit, uint64_t up_mask, uint64_t left_mask, int up_left_avail,
                          int bit_depth, pel *up /* [-1 .. w+h) */, pel *left /* [-1 .. w+h) */)
{
    const pel dflt = (pel)(1 << (bit_depth - 1));
    const int n = (w + h) / unit;
    up[-1] = up_left_avail ? rec[-s - 1] : dflt;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < unit; j++) up[i * unit + j] = ((up_mask >> i) & 1) ? rec[-s + i * unit + j] : dflt;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < unit; j++) left[i * unit + j] = ((left_mask >> i) & 1) ? rec[(i * unit + j) * s - 1] : dflt;
    left[-1] = up[-1];
}

/* xevd_ipred_b: IPD_DC_B 0, IPD_HOR_B 1, IPD_VER_B 2, IPD_UL_B 3, IPD_UR_B 4 (xevd_def.h:332-344); chroma modes share them */
void orc_ipred_base(const pel *left, const pel *up, pel *dst, int mode, int w, int h)
{
    int lw = 0;
    while ((1 << lw) < w) lw++;
    switch (mode) {
    case 0: {                                    /* DC: divides by 2w whatever h is (xevd_ipred.c:146-160) */
        int dc = 0;
        for (int i = 0; i < h; i++) dc += left[i];
        for (int j = 0; j < w; j++) dc += up[j];
        dc = (dc + w) >> (lw + 1);
        for (int i = 0; i < w * h; i++) dst[i] = (pel)dc;
        break;
    }
    case 1:
        for (int i = 0; i < h; i++) for (int j = 0; j < w; j++) dst[i * w + j] = left[i];
        break;
    case 2:
        for (int i = 0; i < h; i++) for (int j = 0; j < w; j++) dst[i * w + j] = up[j];
        break;
    case 3:                                      /* down-right diagonal copy */
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) dst[i * w + j] = i > j ? left[i - j - 1] : (i == j ? up[-1] : up[j - i - 1]);
        break;
    default:                                     /* 4: average of the up-right and down-left diagonals */
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) dst[i * w + j] = (pel)((up[i + j + 1] + left[i + j + 1]) >> 1);
        break;
    }
}

/* ======================================================================================================================
 * Main profile (tool_eipd): 33 luma modes, 5 chroma modes, three reference arrays (left / up / right).
 * Restates xevdm_get_nbr (src_main/xevdm_ipred.c:39-150), xevdm_ipred / xevdm_ipred_uv (:241-305) and the shared
 * predictors xevd_ipred_vert, xevd_get_dc, xevd_ipred_plane, xevd_ipred_bi, ipred_ang (src_base/xevd_ipred.c:110-585).
 * ====================================================================================================================== */

/* Neighbour gather with "replicate the previous sample" substitution.  Arrays are indexed [-1 .. w+h).
 * Net effect of xevdm_get_nbr on the positions the predictors can read (they clamp to [-1, w+h-1]):
 *   up[-1]   = the up-left sample when available, else up[0] AFTER the up row has been filled (the reference's loop over
 *              the units left of the corner overwrites up[-1] with up[0] when the corner unit is unavailable, :85-104)
 *   up[i]    = sample above, or the last filled sample (starting from 1 << (bd-1) at up[-1])
 *   left[-1] = up[-1]; left[i] likewise downwards;  right[-1] = up[w]; right[i] likewise (column x = w) */
void orc_intra_neighbours_main(const pel *rec, int s, int w, int h, int unit, uint64_t up_mask, uint64_t left_mask, uint64_t right_mask,
                               int up_left_avail, int bit_depth, pel *up, pel *left, pel *right)
{
    const int n = (w + h) / unit;
    up[-1] = up_left_avail ? rec[-s - 1] : (pel)(1 << (bit_depth - 1));
    for (int i = 0; i < n; i++)
        for (int j = 0; j < unit; j++) up[i * unit + j] = ((up_mask >> i) & 1) ? rec[-s + i * unit + j] : up[i * unit - 1];
    if (!up_left_avail) up[-1] = up[0];
    left[-1] = up[-1];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < unit; j++) left[i * unit + j] = ((left_mask >> i) & 1) ? rec[(i * unit + j) * s - 1] : left[i * unit - 1];
    right[-1] = up[w];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < unit; j++) right[i * unit + j] = ((right_mask >> i) & 1) ? rec[(i * unit + j) * s + w] : right[i * unit - 1];
}

static int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }
static const int k_inv_size_plus1[8] = { 2048, 1365, 819, 455, 241, 124, 63, 32 };   /* ~ 4096 / (2^k + 1), xevd_ipred.c:108 */

/* horizontal (xevdm_ipred.c:153-196) */
static void pred_hor(const pel *le, const pel *ri, int lr, pel *dst, int w, int h)
{
    const int mul = k_inv_size_plus1[ilog2(w)];
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            dst[y * w + x] = lr == 3 ? (pel)(((le[y] * (w - x) + ri[y] * (x + 1) + (w >> 1)) * mul) >> 12) : (lr == 2 ? ri[y] : le[y]);
}

/* DC (xevdm_ipred.c:198-229 with xevd_get_dc, xevd_ipred.c:124-144) */
static void pred_dc(const pel *le, const pel *up, const pel *ri, int lr, pel *dst, int w, int h)
{
    int sum = 0, hh = h;
    for (int j = 0; j < w; j++) sum += up[j];
    if (lr == 3) { for (int i = 0; i < h; i++) sum += le[i] + ri[i]; sum += (w + h + h) >> 1; hh = h << 1; }
    else { const pel *sd = lr == 2 ? ri : le; for (int i = 0; i < h; i++) sum += sd[i]; sum += (w + h) >> 1; }
    const int lw = ilog2(w), lh = ilog2(hh);
    const int dc = (sum * k_inv_size_plus1[lw > lh ? lw - lh : lh - lw]) >> ((lw < lh ? lw : lh) + 12);
    for (int i = 0; i < w * h; i++) dst[i] = (pel)dc;
}

/* planar (xevd_ipred.c:163-249): gradients b (horizontal) and c (vertical) from the up row and one side column; with a right
 * column available the plane is anchored on the right side and built right-to-left */
static void pred_plane(const pel *le, const pel *up, const pel *ri, int lr, pel *dst, int w, int h, int bd)
{
    static const int mult[6] = { 13, 17, 5, 11, 23, 47 }, shft[6] = { 7, 10, 11, 15, 19, 23 };
    const int w2 = w >> 1, h2 = h >> 1, maxv = (1 << bd) - 1;
    const int iw = ilog2(w) < 2 ? 0 : ilog2(w) - 2, ih = ilog2(h) < 2 ? 0 : ilog2(h) - 2;
    const int from_right = lr == 2 || lr == 3;
    const pel *side = from_right ? ri : le;
    int ch = 0, cv = 0;
    for (int x = 1; x <= w2; x++) ch += from_right ? x * (up[w2 - x] - up[w2 + x]) : x * (up[w2 - 1 + x] - up[w2 - 1 - x]);
    for (int y = 1; y <= h2; y++) cv += y * (side[h2 - 1 + y] - side[h2 - 1 - y]);
    const int a = (side[h - 1] + (from_right ? up[0] : up[w - 1])) << 4;
    const int b = ((ch << 5) * mult[iw] + (1 << (shft[iw] - 1))) >> shft[iw];
    const int c = ((cv << 5) * mult[ih] + (1 << (shft[ih] - 1))) >> shft[ih];
    const int t0 = a - (h2 - 1) * c - (w2 - 1) * b + 16;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int k = from_right ? w - 1 - x : x;          /* steps of b away from the anchor side */
            dst[y * w + x] = (pel)orc_clip3(0, maxv, (t0 + y * c + k * b) >> 5);
        }
}

/* bilinear (xevd_ipred.c:251-372) */
static void pred_bi(const pel *le, const pel *up, const pel *ri, int lr, pel *dst, int w, int h, int bd)
{
    static const int tbl_wc[6] = { -1, 341, 205, 114, 60, 31 };
    const int lx = ilog2(w), ly = ilog2(h), lmin = lx < ly ? lx : ly, maxv = (1 << bd) - 1;
    if (lr == 3) {
        /* both sides: horizontal blend of left/right, vertical blend of up and the bottom row of the horizontal blend */
        const int mul = k_inv_size_plus1[lx];
        int *hb = (int *)malloc(sizeof(int) * w * h);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) hb[y * w + x] = (le[y] * (w - x) + ri[y] * (x + 1) + (w >> 1)) * mul >> 12;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const int vb = (up[x] * (h - 1 - y) + hb[(h - 1) * w + x] * (y + 1) + (h >> 1)) >> ly;
                dst[y * w + x] = (pel)((hb[y * w + x] + vb + 1) >> 1);
            }
        free(hb);
        return;
    }
    /* one side: corner samples a (far end of the up row) and b (far end of the side column), their weighted mean c */
    const int from_right = lr == 2;
    const pel *side = from_right ? ri : le;
    const int a = from_right ? up[-1] : up[w], b = side[h];
    const int wc = tbl_wc[lx > ly ? lx - ly : ly - lx];
    const int c = w == h ? (a + b + 1) >> 1 : (((a << lx) + (b << ly)) * wc + (1 << (lmin + 9))) >> (lmin + 10);
    const int wt = (c << 1) - a - b;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int k = from_right ? w - 1 - x : x;          /* distance index from the side column */
            const int px = (side[y] << lx) + (k + 1) * (a - side[y]);
            const int py = (up[x] << ly) + (y + 1) * (b - up[x]);
            const int v = ((px << ly) + (py << lx) + k * y * wt + (1 << (lx + ly))) >> (lx + ly + 1);
            dst[y * w + x] = (pel)orc_clip3(0, maxv, v);
        }
}

static const int k_dxdy[33][2] = {      /* xevd_tbl_ipred_dxdy (xevd_tbl.c:294-304): {dx/dy, dy/dx} in 1/1024 */
    { 0, 0 }, { 0, 0 }, { 0, 0 }, { 2816, 372 }, { 2048, 512 }, { 1408, 744 }, { 1024, 1024 }, { 744, 1408 }, { 512, 2048 }, { 372, 2816 },
    { 256, 4096 }, { 128, 8192 }, { 0, 0 }, { 128, 8192 }, { 256, 4096 }, { 372, 2816 }, { 512, 2048 }, { 744, 1408 }, { 1024, 1024 },
    { 1408, 744 }, { 2048, 512 }, { 2816, 372 }, { 4096, 256 }, { 8192, 128 }, { 0, 0 }, { 8192, 128 }, { 4096, 256 }, { 2816, 372 },
    { 2048, 512 }, { 1408, 744 }, { 1024, 1024 }, { 744, 1408 }, { 512, 2048 } };

/* one angular sample (ipred_ang_val, xevd_ipred.c:377-570).  The projected reference position is (whole, frac/32); the
 * 4-tap filter xevd_tbl_ipred_adi[frac] = {32-f, 64-f, 32+f, f} is applied along the reference array in the direction
 * `step`, positions clamped to [-1, w+h-1]. */
static pel pred_ang_px(const pel *up, const pel *le, const pel *ri, int lr, int ipm, int i, int j, int w, int h, int bd)
{
    const int mdx = k_dxdy[ipm][0], mdy = k_dxdy[ipm][1];
    const int right_ok = lr == 2 || lr == 3;
    const int dxy = (ipm > 24 || ipm < 12) ? -1 : 1;
    const pel *src;
    int pos, frac, step;
#define PROJ(m, d, whole) do { const int t_ = (d) * (m); (whole) = t_ >> 10; frac = (t_ >> 5) - ((whole) << 5); } while (0)
    int t;
    if (ipm < 12) {                               /* up-right family: project onto the up row (or the right column) */
        PROJ(mdx, j + 1, t);
        if (right_ok && i >= w - t) { int ty; PROJ(mdy, w - i, ty); src = ri; pos = j - ty; step = dxy > 0 ? 1 : -1; }
        else { src = up; pos = i + t; step = dxy < 0 ? 1 : -1; }
    } else if (ipm > 24) {                        /* down-left family */
        if (right_ok) {
            int ty; PROJ(mdy, w - i, ty);
            if (j < ty) { PROJ(mdx, w - i, t); src = up; pos = i + t; step = dxy < 0 ? 1 : -1; }
            else { src = ri; pos = j - ty; step = dxy > 0 ? 1 : -1; }
        } else { int ty; PROJ(mdy, i + 1, ty); src = le; pos = j + ty; step = dxy < 0 ? 1 : -1; }
    } else {                                      /* between vertical and horizontal: up row or left (right) column */
        int ty; PROJ(mdy, i + 1, ty);
        if (j < ty) { PROJ(mdx, j + 1, t); src = up; pos = i - t; step = dxy < 0 ? 1 : -1; }
        else if (lr == 2) { PROJ(mdy, w - i, ty); src = ri; pos = j + ty; step = dxy > 0 ? 1 : -1; }
        else { src = le; pos = j - ty; step = dxy < 0 ? 1 : -1; }
    }
#undef PROJ
    const int lo = -1, hi = w + h - 1;
    const int p0 = orc_clip3(lo, hi, pos - step), p1 = orc_clip3(lo, hi, pos), p2 = orc_clip3(lo, hi, pos + step), p3 = orc_clip3(lo, hi, pos + 2 * step);
    /* the reference narrows the filter result to pel (s16) before clipping */
    const pel v = (pel)((src[p0] * (32 - frac) + src[p1] * (64 - frac) + src[p2] * (32 + frac) + src[p3] * frac + 64) >> 7);
    return (pel)orc_clip3(0, (1 << bd) - 1, v);
}

/* xevdm_ipred: ipm 0 DC, 1 planar, 2 bilinear, 12 vertical, 24 horizontal, others angular */
void orc_ipred_main(const pel *left, const pel *up, const pel *right, int avail_lr, pel *dst, int ipm, int w, int h, int bit_depth)
{
    switch (ipm) {
    case 12: for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) dst[y * w + x] = up[x]; break;
    case 24: pred_hor(left, right, avail_lr, dst, w, h); break;
    case 0:  pred_dc(left, up, right, avail_lr, dst, w, h); break;
    case 1:  pred_plane(left, up, right, avail_lr, dst, w, h, bit_depth); break;
    case 2:  pred_bi(left, up, right, avail_lr, dst, w, h, bit_depth); break;
    default:
        for (int j = 0; j < h; j++)
            for (int i = 0; i < w; i++) dst[j * w + i] = pred_ang_px(up, left, right, avail_lr, ipm, i, j, w, h, bit_depth);
    }
}

/* xevdm_ipred_uv: chroma mode 0 = derived from luma (DM), 1 bilinear, 2 DC, 3 horizontal, 4 vertical */
void orc_ipred_uv_main(const pel *left, const pel *up, const pel *right, int avail_lr, pel *dst, int ipm_c, int ipm, int w, int h, int bit_depth)
{
    static const int to_luma[5] = { -1, 2, 0, 24, 12 };
    orc_ipred_main(left, up, right, avail_lr, dst, ipm_c == 0 ? ipm : to_luma[ipm_c], w, h, bit_depth);
}
