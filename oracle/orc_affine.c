/*
 * Affine motion compensation (Main profile, tool_affine).  TEST INFRASTRUCTURE ONLY (orc_common.h).
 * Restates xevdm_affine_mc / xevdm_affine_mc_lc / xevdm_eif_mc + helpers (src_main/xevdm_mc.c:2108-2685), the sub-block size and EIF
 * applicability rules (src_main/xevdm_util.c:1870-2122) and the per-SCU vectors xevdm_set_affine_mvf publishes (:4095-4203).
 *
 * Two prediction paths per reference list:
 *   - sub-blocks of at least 8x8: ordinary 8/4-tap interpolation with ONE vector for the whole CU - the reference evaluates the model
 *     at (sub_w/2, sub_h/2) for every sub-block (xevdm_mc.c:2359-2360 has no per-sub-block offset), kept as is;
 *   - otherwise EIF: per-sample bilinear fetch at the model's vector (1/32 sample, optionally clamped to a window around the centre
 *     vector), then a separable {-1, 10, -1} sharpening filter with s16 intermediates.
 */
#include <string.h>
#include <stdlib.h>
#include "orc_common.h"

static int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }
static int round_s32(int v, int sh) { return (v + (sh > 0 ? 1 << (sh - 1) : 0) - (v >= 0)) >> sh; }     /* xevdm_rounding_s32 */

/* model gradients in 1/4 * 2^-prec sample per sample: (:2195-2206); cp = control point vectors [vertex][xy] */
static void gradients(const int16_t cp[3][2], int w, int h, int six, int prec, int dh[2], int dv[2])
{
    for (int c = 0; c < 2; c++) dh[c] = ((cp[1][c] - cp[0][c]) << prec) >> ilog2(w);
    if (six) for (int c = 0; c < 2; c++) dv[c] = ((cp[2][c] - cp[0][c]) << prec) >> ilog2(h);
    else { dv[0] = -dh[1]; dv[1] = dh[0]; }
}

/* xevdm_check_eif_applicability_uni (xevdm_util.c:2073-2097) */
static int eif_applicable(const int16_t cp[3][2], int w, int h, int six, int *mem_ok)
{
    int dh[2], dv[2];
    const int p = 9;
    gradients(cp, w, h, six, 7, dh, dv);
    /* bounding box of a 4x4 block's fetch area (calculate_bounding_box_size) */
    int cx[4] = { 0, 5 * (dh[0] + (1 << p)), 5 * dv[0], 0 }, cy[4] = { 0, 5 * dh[1], 5 * (dv[1] + (1 << p)), 0 };
    cx[3] = cx[1] + cx[2]; cy[3] = cy[1] + cy[2];
    int mx[2] = { cx[0], cy[0] }, mn[2] = { cx[0], cy[0] };
    for (int i = 1; i < 4; i++) { mx[0] = orc_max(mx[0], cx[i]); mn[0] = orc_min(mn[0], cx[i]); mx[1] = orc_max(mx[1], cy[i]); mn[1] = orc_min(mn[1], cy[i]); }
    const int bw = ((mx[0] - mn[0] + (1 << p) - 1) >> p) + 2, bh = ((mx[1] - mn[1] + (1 << p) - 1) >> p) + 2;
    *mem_ok = bw * bh <= 72;
    if (dv[1] < -(1 << p)) return 0;
    if ((orc_max(0, dv[1]) + abs(dh[1])) * 5 > (1 << p)) return 0;
    return 1;
}

/* xevdm_derive_affine_subblock_size_bi (xevdm_util.c:1870-1945) */
void orc_affine_subblock(const int16_t cp[2][3][2], const int8_t refi[2], int w, int h, int six, int *sub_w, int *sub_h, int *mem_ok)
{
    static const int lut[4] = { 32, 16, 8, 8 };
    int sw = w, sh = h, apply = 1;
    *mem_ok = 1;
    for (int l = 0; l < 2; l++) {
        if (refi[l] < 0) continue;
        int dh[2], dv[2];
        gradients(cp[l], w, h, six, 7, dh, dv);
        const int wx = orc_max(abs(dh[0]), abs(dh[1])), wy = orc_max(abs(dv[0]), abs(dv[1]));
        sw = orc_min(sw, wx > 4 ? 4 : (wx == 0 ? w : lut[wx - 1]));
        sh = orc_min(sh, wy > 4 ? 4 : (wy == 0 ? h : lut[wy - 1]));
    }
    for (int l = 0; l < 2 && apply; l++) {
        if (refi[l] < 0) continue;
        int ok;
        if (!eif_applicable(cp[l], w, h, six, &ok)) apply = 0;       /* the reference returns at the first inapplicable list */
        *mem_ok &= ok;
    }
    if (!apply) { sw = orc_max(sw, 8); sh = orc_max(sh, 8); }
    *sub_w = sw; *sub_h = sh;
}

/* xevdm_eif_mc for one plane (xevdm_mc.c:2543-2604 with eif_bilinear_clip :2457-2497 and eif_filter :2425-2455) */
static void eif_plane(const pel *ref, int s, int x, int y, int bw, int bh, const int mv0_[2], const int dh[2], const int dv[2],
                      const int mx_[2], const int mn_[2], int chroma, pel *dst, int bd)
{
    int mv0[2] = { mv0_[0], mv0_[1] }, mx[2] = { mx_[0], mx_[1] }, mn[2] = { mn_[0], mn_[1] };
    if (chroma) { for (int c = 0; c < 2; c++) { mv0[c] >>= 1; mx[c] >>= 1; mn[c] >>= 1; } bw >>= 1; bh >>= 1; x >>= 1; y >>= 1; }
    const int s1 = orc_min(4, bd - 8), s2 = orc_max(8, 20 - bd), sh_h = orc_max(bd + 5 - 16, 0), sh_v = 6 - sh_h;
    const int st = bw + 2;
    pel *b = (pel *)malloc(sizeof(pel) * st * (bh + 2));
    for (int j = -1; j <= bh; j++)
        for (int i = -1; i <= bw; i++) {
            const int vx = orc_clip3(mn[0], mx[0], (mv0[0] + i * dh[0] + j * dv[0]) >> 4), vy = orc_clip3(mn[1], mx[1], (mv0[1] + i * dh[1] + j * dv[1]) >> 4);
            const pel *r = ref + (y + j + (vy >> 5)) * s + x + i + (vx >> 5);
            const int fx = vx & 31, fy = vy & 31;
            const pel a0 = (pel)(((64 - 2 * fx) * r[0] + 2 * fx * r[1]) >> s1), a1 = (pel)(((64 - 2 * fx) * r[s] + 2 * fx * r[s + 1]) >> s1);
            b[(j + 1) * st + i + 1] = (pel)(((64 - 2 * fy) * a0 + 2 * fy * a1 + (1 << (s2 - 1))) >> s2);
        }
    /* horizontal then vertical {-1, 10, -1}; the rounding term of the first pass is 1 << (shift - 1), which for shift 0 only touches
     * bits the s16 store drops */
    pel *hb = (pel *)malloc(sizeof(pel) * bw * (bh + 2));
    for (int j = 0; j < bh + 2; j++)
        for (int i = 0; i < bw; i++)
            hb[j * bw + i] = (pel)((-b[j * st + i] + 10 * b[j * st + i + 1] - b[j * st + i + 2] + (sh_h ? 1 << (sh_h - 1) : 0)) >> sh_h);
    for (int j = 0; j < bh; j++)
        for (int i = 0; i < bw; i++) {
            const pel v = (pel)((-hb[j * bw + i] + 10 * hb[(j + 1) * bw + i] - hb[(j + 2) * bw + i] + (1 << (sh_v - 1))) >> sh_v);
            dst[j * bw + i] = (pel)orc_clip3(0, (1 << bd) - 1, v);
        }
    free(b); free(hb);
}

/* prediction of all three planes from one list: xevdm_affine_mc_lc (xevdm_mc.c:2259-2392) */
static void affine_list(const XB200_PARAMS *prm, int x, int y, int w, int h, const int16_t cp[3][2], int six, const ORC_PIC *rp,
                        int sub_w, int sub_h, int mem_ok, pel *py, pel *pu, pel *pv)
{
    int dh[2], dv[2];
    gradients(cp, w, h, six, 7, dh, dv);
    const int sc[2] = { cp[0][0] << 7, cp[0][1] << 7 };
    const int bdl = prm->bit_depth_luma, bdc = prm->bit_depth_chroma;
    if (sub_w < 8 || sub_h < 8) {
        /* eif_derive_mv_clip_range (:2108-2150): 1/32-sample limits; without the memory-bandwidth guarantee the vectors are confined to
         * a window around the centre vector */
        static const int spread_tbl[5] = { 128, 256, 544, 1120, 2272 };
        const int pmx[2] = { (prm->w + 128 - x - w - 1) << 5, (prm->h + 128 - y - h - 1) << 5 }, pmn[2] = { (-x - 128) << 5, (-y - 128) << 5 };
        int mx[2], mn[2];
        for (int c = 0; c < 2; c++) {
            if (mem_ok) { mx[c] = pmx[c]; mn[c] = pmn[c]; }
            else {
                const int centre = round_s32(sc[c] + dh[c] * (w >> 1) + dv[c] * (h >> 1), 4);
                const int spread = spread_tbl[ilog2(c == 0 ? w : h) - 3];
                mn[c] = centre - spread; mx[c] = centre + spread;
                if (mn[c] < pmn[c]) { mn[c] = pmn[c]; mx[c] = orc_min(pmx[c], pmn[c] + 2 * spread); }
                else if (mx[c] > pmx[c]) { mx[c] = pmx[c]; mn[c] = orc_max(pmn[c], pmx[c] - 2 * spread); }
            }
            mx[c] = orc_clip3(-(1 << 17), (1 << 17) - 1, mx[c]);
            mn[c] = orc_clip3(-(1 << 17), (1 << 17) - 1, mn[c]);
        }
        eif_plane(rp->y, rp->s_l, x, y, w, h, sc, dh, dv, mx, mn, 0, py, bdl);
        eif_plane(rp->u, rp->s_c, x, y, w, h, sc, dh, dv, mx, mn, 1, pu, bdc);
        eif_plane(rp->v, rp->s_c, x, y, w, h, sc, dh, dv, mx, mn, 1, pv, bdc);
        return;
    }
    /* one vector (1/16 sample) for every sub-block; variant from the unclipped vector, phase from the clipped one */
    int mvo[2], mvc[2];
    for (int c = 0; c < 2; c++) mvo[c] = orc_clip3(-(1 << 17), (1 << 17) - 1, round_s32(sc[c] + dh[c] * (sub_w >> 1) + dv[c] * (sub_h >> 1), 5));
    mvc[0] = orc_min((prm->w + 128 - x - w) << 4, orc_max((-128 - x) << 4, mvo[0]));
    mvc[1] = orc_min((prm->h + 128 - y - h) << 4, orc_max((-128 - y) << 4, mvo[1]));
    for (int sy = 0; sy < h; sy += sub_h)
        for (int sx = 0; sx < w; sx += sub_w) {
            const int gx = ((x + sx) << 4) + mvc[0], gy = ((y + sy) << 4) + mvc[1];
            orc_mc_luma(rp->y, rp->s_l, gx, gy, mvo[0], mvo[1], py + sy * w + sx, w, sub_w, sub_h, bdl, prm->tool_admvp);
            orc_mc_chroma(rp->u, rp->s_c, gx, gy, mvo[0], mvo[1], pu + (sy >> 1) * (w >> 1) + (sx >> 1), w >> 1, sub_w >> 1, sub_h >> 1, bdc, prm->tool_admvp);
            orc_mc_chroma(rp->v, rp->s_c, gx, gy, mvo[0], mvo[1], pv + (sy >> 1) * (w >> 1) + (sx >> 1), w >> 1, sub_w >> 1, sub_h >> 1, bdc, prm->tool_admvp);
        }
}

/* xevdm_affine_mc (xevdm_mc.c:2606-2685); cp[list][vertex][xy] quarter-sample control point vectors */
void orc_affine_pred(const XB200_PARAMS *prm, int x, int y, int w, int h, const int8_t refi[2], const int16_t cp[2][3][2], int six,
                     const ORC_PIC *const *refs_l0, const ORC_PIC *const *refs_l1, pel *py, pel *pu, pel *pv)
{
    int sub_w, sub_h, mem_ok, n = 0;
    const int cw = w >> 1, ch = h >> 1;
    pel *t = (pel *)malloc(sizeof(pel) * (w * h + 2 * cw * ch));
    pel *out[2][3] = { { py, pu, pv }, { t, t + w * h, t + w * h + cw * ch } };
    orc_affine_subblock(cp, refi, w, h, six, &sub_w, &sub_h, &mem_ok);
    for (int l = 0; l < 2; l++) {
        if (refi[l] < 0) continue;
        affine_list(prm, x, y, w, h, cp[l], six, (l ? refs_l1 : refs_l0)[refi[l]], sub_w, sub_h, mem_ok, out[n][0], out[n][1], out[n][2]);
        n++;
    }
    if (n == 2) {
        for (int i = 0; i < w * h; i++) py[i] = (pel)((py[i] + t[i] + 1) >> 1);
        for (int i = 0; i < cw * ch; i++) { pu[i] = (pel)((pu[i] + out[1][1][i] + 1) >> 1); pv[i] = (pel)((pv[i] + out[1][2][i] + 1) >> 1); }
    }
    free(t);
}

/* vectors xevdm_set_affine_mvf leaves in map_mv for list l (xevdm_util.c:4095-4203): the control point vectors at three corner
 * sub-blocks, the rounded model vector (quarter sample) at the centre of every other sub-block; out[scu][xy], CU-raster SCU order */
void orc_affine_map_mv(const int16_t cp[2][3][2], const int8_t refi[2], int log2w, int log2h, int six, int l, int16_t *out)
{
    const int w = 1 << log2w, h = 1 << log2h, wc = w >> 2, hc = h >> 2;
    int sub_w, sub_h, mem_ok;
    orc_affine_subblock(cp, refi, w, h, six, &sub_w, &sub_h, &mem_ok);
    const int sws = sub_w >> 2, shs = sub_h >> 2;
    const int16_t (*v)[2] = cp[l];
    const int dhx = (v[1][0] - v[0][0]) << (7 - log2w), dhy = (v[1][1] - v[0][1]) << (7 - log2w);
    const int dvx = six ? (v[2][0] - v[0][0]) << (7 - log2h) : -dhy, dvy = six ? (v[2][1] - v[0][1]) << (7 - log2h) : dhx;
    for (int sy = 0; sy < hc; sy += shs)
        for (int sx = 0; sx < wc; sx += sws) {
            int mx, my;
            if (sx == 0 && sy == 0) { mx = v[0][0]; my = v[0][1]; }
            else if (sx + sws == wc && sy == 0) { mx = v[1][0]; my = v[1][1]; }
            else if (sx == 0 && sy + shs == hc && six) { mx = v[2][0]; my = v[2][1]; }
            else {
                const int px = (sx << 2) + (sub_w >> 1), py = (sy << 2) + (sub_h >> 1);
                mx = orc_clip3(-(1 << 17), (1 << 17) - 1, round_s32((v[0][0] << 7) + dhx * px + dvx * py, 5)) >> 2;
                my = orc_clip3(-(1 << 17), (1 << 17) - 1, round_s32((v[0][1] << 7) + dhy * px + dvy * py, 5)) >> 2;
            }
            for (int j = sy; j < sy + shs; j++)
                for (int i = sx; i < sx + sws; i++) { out[(j * wc + i) * 2] = (int16_t)mx; out[(j * wc + i) * 2 + 1] = (int16_t)my; }
        }
}
