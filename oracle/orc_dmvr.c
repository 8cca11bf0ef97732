/*
 * Decoder-side motion vector refinement (Main profile, tool_dmvr).  TEST INFRASTRUCTURE ONLY (orc_common.h).
 * Restates the DMVR branch of xevdm_mc (src_main/xevdm_mc.c:1860-2038): processDMVR (:1638-1825), xevd_DMVR_cost /
 * xevd_DMVR_refine (:1270-1336), xevd_SubPelErrorSrfc + div_for_maxq7 (:1338-1428), prefetch_for_mc + padding (:1430-1544),
 * final_paddedMC_forDMVR (:1546-1636), the bilinear kernels xevdm_bl_mc_l_* (:356-476) and the 8/4-tap kernels
 * xevd_mc_dmvr_l_* / _c_* (:222-354, :482-627).
 */
#include <string.h>
#include <stdlib.h>
#include <limits.h>
#include "orc_common.h"

/* xevdm_bl_mc_l: 2-tap interpolation, taps {64 - 4f, 4f} for the 1/16 phase f (xevd_tbl_bl_mc_l_coeff, :93-119); variant picked by the
 * position's own fraction (xevdm_mc.h:70-73); same shift / rounding rules as the long filters */
static void bilinear(const pel *ref, int s_ref, int gx, int gy, pel *pred, int s_pred, int w, int h, int bd)
{
    const int dx = gx & 15, dy = gy & 15, maxv = (1 << bd) - 1;
    const int cx0 = 64 - 4 * dx, cx1 = 4 * dx, cy0 = 64 - 4 * dy, cy1 = 4 * dy;
    ref += (gy >> 4) * s_ref + (gx >> 4);
    if (!dx && !dy) {
        for (int i = 0; i < h; i++) memcpy(pred + i * s_pred, ref + i * s_ref, sizeof(pel) * w);
    } else if (dx && !dy) {
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) pred[i * s_pred + j] = (pel)orc_clip3(0, maxv, (cx0 * ref[i * s_ref + j] + cx1 * ref[i * s_ref + j + 1]) >> 6);
    } else if (!dx) {
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) pred[i * s_pred + j] = (pel)orc_clip3(0, maxv, (cy0 * ref[i * s_ref + j] + cy1 * ref[(i + 1) * s_ref + j]) >> 6);
    } else {
        const int s1 = orc_min(4, bd - 8), s2 = orc_max(8, 20 - bd);
        int16_t *t = (int16_t *)malloc(sizeof(int16_t) * w * (h + 1));
        for (int i = 0; i < h + 1; i++)
            for (int j = 0; j < w; j++) t[i * w + j] = (int16_t)((cx0 * ref[i * s_ref + j] + cx1 * ref[i * s_ref + j + 1]) >> s1);
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) pred[i * s_pred + j] = (pel)orc_clip3(0, maxv, (cy0 * t[i * w + j] + cy1 * t[(i + 1) * w + j] + (1 << (s2 - 1))) >> s2);
        free(t);
    }
}

static int sad(const pel *a, const pel *b, int s, int w, int h)
{
    int acc = 0;
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) acc += abs(a[i * s + j] - b[i * s + j]);
    return acc;
}

/* div_for_maxq7 (:1338-1375): 3-bit restoring division of |N| by D << 3, sign restored */
static int div_q7(int64_t n, int64_t d)
{
    int neg = n < 0, q = 0;
    if (neg) n = -n;
    d <<= 3;
    if (n >= d) { n -= d; q++; }
    q <<= 1; d >>= 1;
    if (n >= d) { n -= d; q++; }
    q <<= 1;
    if (n >= (d >> 1)) q++;
    return neg ? -q : q;
}

/* parabolic sub-sample offset in 1/16 from the centre cost c and the costs either side (xevd_SubPelErrorSrfc, :1376-1428) */
static int subpel_axis(int c, int minus, int plus)
{
    const int64_t num = (int64_t)((minus - plus) << 4), den = (int64_t)(minus + plus - (c << 1));
    if (den == 0) return 0;
    if (minus != c && plus != c) return div_q7(num, den);
    return minus == c ? -8 : 8;
}

/* clip of one vector: mv_clip_only_one_ref_dmvr (:939-978); returns whether anything was clipped */
static int clip_one(int x, int y, int pic_w, int pic_h, int w, int h, const int16_t mv[2], int16_t out[2])
{
    const int qx = x << 2, qy = y << 2, qw = w << 2, qh = h << 2;
    const int lo = -(128 << 2), hx = (pic_w - 1 + 128) << 2, hy = (pic_h - 1 + 128) << 2;
    int f = 0;
    out[0] = mv[0]; out[1] = mv[1];
    if (qx + mv[0] < lo) { f = 1; out[0] = (int16_t)(lo - qx); }
    if (qy + mv[1] < lo) { f = 1; out[1] = (int16_t)(lo - qy); }
    if (qx + mv[0] + qw - 4 > hx) { f = 1; out[0] = (int16_t)(hx - qx - qw + 4); }
    if (qy + mv[1] + qh - 4 > hy) { f = 1; out[1] = (int16_t)(hy - qy - qh + 4); }
    return f;
}

/* Final prediction of one sub-PU of one plane from the "padded window": the (bw + ntap - 1) x (bh + ntap - 1) samples the INITIAL
 * vector needs, extended by `pad` replicated samples on every side (prefetch_for_mc + padding); the refined vector moves inside it.
 *   wx, wy   : position of the window's first sample in the reference plane
 *   dxi, dyi : whole-sample displacement of the refined vector relative to the initial one
 *   g        : refined position in 1/16 (luma) or 1/32 (chroma) units (only its fraction is used) */
static void padded_mc(const pel *plane, int s, int wx, int wy, int bw, int bh, int ntap, int pad, int dxi, int dyi, int gx, int gy,
                      const int16_t *taps, int fmask, pel *pred, int s_pred, int bd)
{
    const int half = ntap / 2 - 1, ww = bw + ntap - 1, wh = bh + ntap - 1;
    const int ew = ww + 2 * pad, eh = wh + 2 * pad;
    pel *e = (pel *)malloc(sizeof(pel) * ew * eh);
    for (int i = 0; i < eh; i++)
        for (int j = 0; j < ew; j++)
            e[i * ew + j] = plane[(wy + orc_clip3(0, wh - 1, i - pad)) * s + wx + orc_clip3(0, ww - 1, j - pad)];
    const int fx = (gx & fmask) != 0, fy = (gy & fmask) != 0;
    orc_interp(e, ew, pad + half + dxi, pad + half + dyi, taps + ntap * (gx & fmask), taps + ntap * (gy & fmask), fx, fy, ntap, pred, s_pred, bw, bh, bd);
    free(e);
}

int orc_dmvr_pred(const XB200_PARAMS *prm, int x, int y, int w, int h, const int8_t refi[2], const int16_t mv[2][2],
                  const ORC_PIC *const *refs_l0, const ORC_PIC *const *refs_l1, pel *pred[2][3], int16_t *dmvr_mv)
{
    if (refi[0] < 0 || refi[1] < 0 || w < 8 || h < 8) return 0;
    const ORC_PIC *rp[2] = { refs_l0[refi[0]], refs_l1[refi[1]] };
    int16_t start[2][2];
    orc_mv_clip(x, y, prm->w, prm->h, w, h, refi, mv, start);
    /* xevdm_mc (:1899-1911): references on opposite sides at equal distance, not the identical-motion case */
    const int d0 = prm->poc - rp[0]->poc, d1 = prm->poc - rp[1]->poc;
    if (!(d0 * d1 < 0 && abs(d0) == abs(d1))) return 0;
    if (rp[0]->poc == rp[1]->poc && start[0][0] == start[1][0] && start[0][1] == start[1][1]) return 0;

    const int bdl = prm->bit_depth_luma, bdc = prm->bit_depth_chroma, iter = 2;
    const int bs = w + 2 * iter;                                   /* row stride of the bilinear search planes */
    pel *bl[2];
    for (int l = 0; l < 2; l++) {
        bl[l] = (pel *)malloc(sizeof(pel) * bs * (h + 2 * iter));
        const int qx = (x << 2) + start[l][0] - (iter << 2), qy = (y << 2) + start[l][1] - (iter << 2);
        bilinear(rp[l]->y, rp[l]->s_l, qx << 2, qy << 2, bl[l], bs, w + 2 * iter, h + 2 * iter, bdl);
    }
    const int16_t *tl = orc_mc_luma_taps(prm->tool_admvp), *tc = orc_mc_chroma_taps(prm->tool_admvp);
    const int dx = orc_min(w, 16), dy = orc_min(h, 16), scuw = w >> 2;
    for (int sy = 0; sy < h; sy += dy)
        for (int sx = 0; sx < w; sx += dx) {
            /* integer search: up to two rounds of {down, up, right, left, best diagonal}; list 1 moves the opposite way */
            const pel *c0 = bl[0] + (iter + sy) * bs + iter + sx, *c1 = bl[1] + (iter + sy) * bs + iter + sx;
            int tot[2] = { 0, 0 }, min_cost = INT_MAX, cost[5], centre = INT_MAX, not_zero = 1;
            for (int i = 0; i < iter; i++) {
                const pel *a0 = c0 + tot[0] + tot[1] * bs, *a1 = c1 - (tot[0] + tot[1] * bs);
                for (int k = 0; k < 5; k++) cost[k] = INT_MAX;
                centre = INT_MAX;
                if (i == 0) min_cost = sad(a0, a1, bs, dx, dy);
                if ((i > 0 && min_cost == 0) || (i == 0 && min_cost < dx * dy)) { not_zero = 0; break; }
                centre = min_cost;
                int ox[5] = { 0, 0, 1, -1, 0 }, oy[5] = { 1, -1, 0, 0, 0 }, best[2] = { 0, 0 };
                for (int k = 0; k < 5; k++) {                          /* SAD_BOTTOM, SAD_TOP, SAD_RIGHT, SAD_LEFT, diagonal */
                    cost[k] = sad(a0 + ox[k] + oy[k] * bs, a1 - ox[k] - oy[k] * bs, bs, dx, dy);
                    if (k == 3) { oy[4] = cost[0] <= cost[1] ? 1 : -1; ox[4] = cost[2] <= cost[3] ? 1 : -1; }
                    if (cost[k] < min_cost) { min_cost = cost[k]; best[0] = ox[k]; best[1] = oy[k]; }
                }
                if (best[0] == 0 && best[1] == 0) break;
                tot[0] += best[0]; tot[1] += best[1];
            }
            int delta[2] = { tot[0] << 4, tot[1] << 4 };
            if (not_zero && min_cost == centre) {                      /* the centre won: parabolic sub-sample step */
                delta[0] += subpel_axis(centre, cost[3], cost[2]);     /* left, right */
                delta[1] += subpel_axis(centre, cost[1], cost[0]);     /* top, bottom */
            }
            int refined[2][2];                                          /* 1/16 sample */
            for (int d = 0; d < 2; d++) {
                refined[0][d] = (start[0][d] << 2) + (int16_t)delta[d];
                refined[1][d] = (start[1][d] << 2) - (int16_t)delta[d];
            }
            for (int j = 0; j < dy >> 2; j++)
                for (int i = 0; i < dx >> 2; i++) {
                    int16_t *o = dmvr_mv + (((sy >> 2) + j) * scuw + (sx >> 2) + i) * 4;
                    o[0] = (int16_t)(refined[0][0] >> 2); o[1] = (int16_t)(refined[0][1] >> 2);
                    o[2] = (int16_t)(refined[1][0] >> 2); o[3] = (int16_t)(refined[1][1] >> 2);
                }
            /* final prediction of the sub-PU from the padded windows */
            const int px = x + sx, py = y + sy;
            for (int l = 0; l < 2; l++) {
                int16_t st_c[2], un[2] = { (int16_t)(refined[l][0] >> 2), (int16_t)(refined[l][1] >> 2) }, cl[2];
                clip_one(x, y, prm->w, prm->h, w, h, start[l], st_c);                  /* window position: CU-level clip of the start */
                const int wgx = ((px << 2) + st_c[0]) << 2, wgy = ((py << 2) + st_c[1]) << 2;
                const int clipped = clip_one(px, py, prm->w, prm->h, dx, dy, un, cl);  /* sub-PU-level clip of the refined vector */
                int gx, gy, dlx, dly, dcx, dcy;
                if (clipped) {
                    gx = (px << 4) + (cl[0] << 2); gy = (py << 4) + (cl[1] << 2);
                    dlx = (cl[0] >> 2) - (start[l][0] >> 2); dly = (cl[1] >> 2) - (start[l][1] >> 2);
                    dcx = (cl[0] >> 3) - (start[l][0] >> 3); dcy = (cl[1] >> 3) - (start[l][1] >> 3);
                } else {
                    gx = (px << 4) + refined[l][0]; gy = (py << 4) + refined[l][1];
                    dlx = (refined[l][0] >> 4) - (start[l][0] >> 2); dly = (refined[l][1] >> 4) - (start[l][1] >> 2);
                    dcx = (refined[l][0] >> 5) - (start[l][0] >> 3); dcy = (refined[l][1] >> 5) - (start[l][1] >> 3);
                }
                padded_mc(rp[l]->y, rp[l]->s_l, (wgx >> 4) - 3, (wgy >> 4) - 3, dx, dy, 8, 2, dlx, dly, gx, gy, tl, 15,
                          pred[l][0] + sy * w + sx, w, bdl);
                padded_mc(rp[l]->u, rp[l]->s_c, (wgx >> 5) - 1, (wgy >> 5) - 1, dx >> 1, dy >> 1, 4, 1, dcx, dcy, gx, gy, tc, 31,
                          pred[l][1] + (sy >> 1) * (w >> 1) + (sx >> 1), w >> 1, bdc);
                padded_mc(rp[l]->v, rp[l]->s_c, (wgx >> 5) - 1, (wgy >> 5) - 1, dx >> 1, dy >> 1, 4, 1, dcx, dcy, gx, gy, tc, 31,
                          pred[l][2] + (sy >> 1) * (w >> 1) + (sx >> 1), w >> 1, bdc);
            }
        }
    free(bl[0]); free(bl[1]);
    return 1;
}
