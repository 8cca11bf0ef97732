/*
 * Inter prediction (motion compensation).  TEST INFRASTRUCTURE ONLY (see orc_common.h).
 * Restates src_base/xevd_mc.c:145-557 (and the table switch of src_main/xevdm_mc.c:1915-1924).
 */
#include <string.h>
#include <stdlib.h>
#include "orc_common.h"

/* "does this 1/16 (luma) or 1/32 (chroma) pel component have a fractional part" as decided by the
 * dispatch macros xevd_mc_l / xevd_mc_c (xevd_mc.h:66-74): OR of the low 4 (5) bits of the
 * UNCLIPPED vector component (T3). */
static int frac_l(int v) { return (v | (v >> 1) | (v >> 2) | (v >> 3)) & 1; }
static int frac_c(int v) { return (v | (v >> 1) | (v >> 2) | (v >> 3) | (v >> 4)) & 1; }

/*
 * Generic separable interpolation with `ntap` taps, shared by luma (8) and chroma (4).
 *   fx, fy : which directions are filtered (variant 00 / n0 / 0n / nn)
 *   1-D variants: (sum + 0) >> 6, clipped               (xevd_mc.c:186-236, T1)
 *   2-D variant : H pass (sum >> min(4, bd-8)) kept as s16, then
 *                 V pass (sum + 2^(s2-1)) >> s2, s2 = max(8, 20-bd), clipped   (xevd_mc.c:239-284)
 */
void orc_interp(const pel *ref, int s_ref, int ix, int iy, const int16_t *cx, const int16_t *cy,
                int fx, int fy, int ntap, pel *pred, int s_pred, int w, int h, int bd)
{
    const int half = ntap / 2 - 1;        /* taps start `half` samples before the integer position */
    const int maxv = (1 << bd) - 1;
    if (!fx && !fy) {
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) pred[i * s_pred + j] = ref[(iy + i) * s_ref + ix + j];
        return;
    }
    if (fx && !fy) {
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) {
                int32_t acc = 0;
                for (int t = 0; t < ntap; t++) acc += cx[t] * ref[(iy + i) * s_ref + ix + j + t - half];
                pred[i * s_pred + j] = (pel)orc_clip3(0, maxv, acc >> 6);
            }
        return;
    }
    if (!fx && fy) {
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) {
                int32_t acc = 0;
                for (int t = 0; t < ntap; t++) acc += cy[t] * ref[(iy + i + t - half) * s_ref + ix + j];
                pred[i * s_pred + j] = (pel)orc_clip3(0, maxv, acc >> 6);
            }
        return;
    }
    {
        const int s1 = orc_min(4, bd - 8), s2 = orc_max(8, 20 - bd), rnd2 = 1 << (s2 - 1);
        const int rows = h + ntap - 1;
        int16_t *tmp = (int16_t *)malloc((size_t)rows * w * sizeof(int16_t));
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < w; j++) {
                int32_t acc = 0;
                for (int t = 0; t < ntap; t++) acc += cx[t] * ref[(iy + i - half) * s_ref + ix + j + t - half];
                tmp[i * w + j] = (int16_t)(acc >> s1);          /* stored as s16 (xevd_mc.c:243,264) */
            }
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) {
                int32_t acc = 0;
                for (int t = 0; t < ntap; t++) acc += cy[t] * tmp[(i + t) * w + j];
                pred[i * s_pred + j] = (pel)orc_clip3(0, maxv, (acc + rnd2) >> s2);
            }
        free(tmp);
    }
}

/* xevd_mc_l dispatch + xevd_mc_l_{00,n0,0n,nn}: gmv in 1/16 pel, absolute (xevd_mc.c:169-284) */
void orc_mc_luma(const pel *ref, int s_ref, int gmv_x, int gmv_y, int ori_mv_x, int ori_mv_y,
                 pel *pred, int s_pred, int w, int h, int bit_depth, int main_tables)
{
    const int16_t *taps = orc_mc_luma_taps(main_tables);
    orc_interp(ref, s_ref, gmv_x >> 4, gmv_y >> 4, taps + 8 * (gmv_x & 15), taps + 8 * (gmv_y & 15),
           frac_l(ori_mv_x), frac_l(ori_mv_y), 8, pred, s_pred, w, h, bit_depth);
}

/* xevd_mc_c dispatch + xevd_mc_c_{00,n0,0n,nn}: gmv in 1/32 chroma pel (xevd_mc.c:290-408) */
void orc_mc_chroma(const pel *ref, int s_ref, int gmv_x, int gmv_y, int ori_mv_x, int ori_mv_y,
                   pel *pred, int s_pred, int w, int h, int bit_depth, int main_tables)
{
    const int16_t *taps = orc_mc_chroma_taps(main_tables);
    orc_interp(ref, s_ref, gmv_x >> 5, gmv_y >> 5, taps + 4 * (gmv_x & 31), taps + 4 * (gmv_y & 31),
           frac_c(ori_mv_x), frac_c(ori_mv_y), 4, pred, s_pred, w, h, bit_depth);
}

/* xevd_mv_clip (xevd_mc.c:435-467): keep the block within MAX_CU_SIZE(128) of the picture, q-pel units */
void orc_mv_clip(int x, int y, int pic_w, int pic_h, int w, int h, const int8_t refi[2],
                 const int16_t mv[2][2], int16_t mv_t[2][2])
{
    const int qx = x << 2, qy = y << 2, qw = w << 2, qh = h << 2;
    const int lo = -(128 << 2), hx = (pic_w - 1 + 128) << 2, hy = (pic_h - 1 + 128) << 2;
    for (int l = 0; l < 2; l++) {
        mv_t[l][0] = mv[l][0];
        mv_t[l][1] = mv[l][1];
        if (refi[l] < 0) continue;
        if (qx + mv[l][0] < lo) mv_t[l][0] = (int16_t)(lo - qx);
        if (qy + mv[l][1] < lo) mv_t[l][1] = (int16_t)(lo - qy);
        if (qx + mv[l][0] + qw - 4 > hx) mv_t[l][0] = (int16_t)(hx - qx - qw + 4);
        if (qy + mv[l][1] + qh - 4 > hy) mv_t[l][1] = (int16_t)(hy - qy - qh + 4);
    }
}

/* xevd_mc (xevd_mc.c:469-557) / xevdm_mc without DMVR (xevdm_mc.c:1860-2038) */
void orc_inter_pred(const XB200_PARAMS *prm, int x, int y, int w, int h, const int8_t refi[2],
                    const int16_t mv[2][2], const ORC_PIC *const *refs_l0, const ORC_PIC *const *refs_l1,
                    pel *pred_y, pel *pred_u, pel *pred_v)
{
    int16_t mv_t[2][2];
    const int cw = w >> 1, ch = h >> 1;
    const int bdl = prm->bit_depth_luma, bdc = prm->bit_depth_chroma, mt = prm->tool_admvp;
    pel *tmp = (pel *)malloc((size_t)(w * h + 2 * cw * ch) * sizeof(pel));
    pel *out[2][3] = {{pred_y, pred_u, pred_v}, {tmp, tmp + w * h, tmp + w * h + cw * ch}};
    int n = 0;

    orc_mv_clip(x, y, prm->w, prm->h, w, h, refi, mv, mv_t);
    for (int l = 0; l < 2; l++) {
        if (refi[l] < 0) continue;
        if (l == 1 && refi[0] >= 0) {
            /* identical motion: both lists point at the same picture with the same clipped vector
             * -> uni-prediction from list 0 (xevd_mc.c:513-519, T2) */
            if (refs_l0[refi[0]]->poc == refs_l1[refi[1]]->poc && mv_t[0][0] == mv_t[1][0] && mv_t[0][1] == mv_t[1][1]) break;
        }
        const ORC_PIC *rp = (l == 0 ? refs_l0 : refs_l1)[refi[l]];
        const int gx = ((x << 2) + mv_t[l][0]) << 2, gy = ((y << 2) + mv_t[l][1]) << 2;
        const int ox = mv[l][0] << 2, oy = mv[l][1] << 2;       /* unclipped, selects the variant (T3) */
        orc_mc_luma(rp->y, rp->s_l, gx, gy, ox, oy, out[n][0], w, w, h, bdl, mt);
        orc_mc_chroma(rp->u, rp->s_c, gx, gy, ox, oy, out[n][1], cw, cw, ch, bdc, mt);
        orc_mc_chroma(rp->v, rp->s_c, gx, gy, ox, oy, out[n][2], cw, cw, ch, bdc, mt);
        n++;
    }
    if (n == 2) {      /* xevd_average_16b_no_clip: (a + b + 1) >> 1 on the two clipped predictions */
        for (int i = 0; i < w * h; i++) pred_y[i] = (pel)((pred_y[i] + out[1][0][i] + 1) >> 1);
        for (int i = 0; i < cw * ch; i++) {
            pred_u[i] = (pel)((pred_u[i] + out[1][1][i] + 1) >> 1);
            pred_v[i] = (pel)((pred_v[i] + out[1][2][i] + 1) >> 1);
        }
    }
    free(tmp);
}
