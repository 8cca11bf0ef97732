/*
 * Adaptive loop filter (Main profile).  TEST INFRASTRUCTURE ONLY (orc_common.h).
 * Restates:
 *   per-CTU window with 3-sample margins  alf_process_tile            src_main/xevdm_alf.c:901-1165 (T9: mirrored margins)
 *   per-TILE copy with replicated border  alf_copy_and_extend_tile    :805-840 (a window never reads another tile)
 *   which sides of a CTU take the copy    tile_boundary_check         :844-881, called with the tile (flag 0) or with
 *                                                                     (0, width - 1, 0, height - 1) (flag 1), :989-999
 *   4x4 block classification              alf_derive_classification_blk :38-208
 *   7x7 diamond luma / 5x5 diamond chroma alf_filter_blk_7 / _5       :210-429
 * Every CTU filters from a copy of the pre-ALF picture, so CTUs are independent.
 */
#include <stdlib.h>
#include <string.h>
#include "orc_common.h"

typedef struct {
    const pel *p;        /* pre-ALF plane copy, sample (0,0) */
    int s, W, H;         /* stride, plane size                */
    int x0, y0, w, h;    /* CTU rectangle in this plane        */
    int aL, aR, aT, aB;  /* the side's margin comes from the (extended) tile copy; else it mirrors the CTU */
    int tx0, tx1, ty0, ty1;   /* the CTU's tile in this plane */
} Win;

static pel ext(const Win *k, int y, int x)          /* the tile's copy is extended by replication (alf_copy_and_extend_tile) */
{
    y = orc_clip3(k->ty0, k->ty1 - 1, y); x = orc_clip3(k->tx0, k->tx1 - 1, x);
    return k->p[y * k->s + x];
}

/* sample (r, c) of the CTU window, r in [-3, h+3), c in [-3, w+3) */
static pel win(const Win *k, int r, int c)
{
    if (r < 0) return k->aT ? ext(k, k->y0 + r, k->x0 + c) : win(k, -r, c);                          /* rows -3,-2,-1 <- 3,2,1   */
    if (r >= k->h) return k->aB ? ext(k, k->y0 + r, k->x0 + c) : win(k, 2 * k->h - r - 2, c);       /* h,h+1,h+2 <- h-2,h-3,h-4 */
    if (c < 0) return k->aL ? ext(k, k->y0 + r, k->x0 + c) : ext(k, k->y0 + r, k->x0 - c);
    if (c >= k->w) return k->aR ? ext(k, k->y0 + r, k->x0 + c) : ext(k, k->y0 + r, k->x0 + 2 * k->w - c - 2);
    return k->p[(k->y0 + r) * k->s + k->x0 + c];
}

/* class index (0..24) << 2 | transpose index (0..3) of the 4x4 block at (by, bx) of the window */
static int classify(const Win *k, int by, int bx, int bit_depth)
{
    static const int th[16] = {0, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4};
    static const int trans_tbl[8] = {0, 1, 0, 2, 2, 3, 1, 3};
    int sv = 0, sh = 0, sd0 = 0, sd1 = 0;
    for (int r = by - 2; r < by + 6; r++)
        for (int c = bx - 2; c < bx + 6; c++) {
            const int p2 = (int16_t)(win(k, r, c) << 1);
            sv += abs(p2 - win(k, r - 1, c) - win(k, r + 1, c));
            sh += abs(p2 - win(k, r, c - 1) - win(k, r, c + 1));
            sd0 += abs(p2 - win(k, r - 1, c - 1) - win(k, r + 1, c + 1));
            sd1 += abs(p2 - win(k, r + 1, c - 1) - win(k, r - 1, c + 1));
        }
    const int activity = (int16_t)orc_clip3(0, 15, (sv + sh) >> (bit_depth - 2));
    int cls = th[activity];
    int hv1, hv0, d1, d0, dir_hv, dir_d, hvd1, hvd0, main_dir, sec_dir;
    if (sv > sh) { hv1 = sv; hv0 = sh; dir_hv = 1; } else { hv1 = sh; hv0 = sv; dir_hv = 3; }
    if (sd0 > sd1) { d1 = sd0; d0 = sd1; dir_d = 0; } else { d1 = sd1; d0 = sd0; dir_d = 2; }
    if ((int)((unsigned)d1 * (unsigned)hv0) > (int)((unsigned)hv1 * (unsigned)d0))  /* int products, wrap as compiled (:165) */ { hvd1 = d1; hvd0 = d0; main_dir = dir_d; sec_dir = dir_hv; }
    else { hvd1 = hv1; hvd0 = hv0; main_dir = dir_hv; sec_dir = dir_d; }
    int strength = 0;
    if (hvd1 > 2 * hvd0) strength = 1;
    if (hvd1 * 2 > 9 * hvd0) strength = 2;
    if (strength) cls += (((main_dir & 1) << 1) + strength) * 5;
    return ((cls << 2) + trans_tbl[main_dir * 2 + (sec_dir >> 1)]) & 0xff;
}

static void filter_luma_blk(const Win *k, int by, int bx, int cl, const int16_t coef_final[25][13], pel *dst, int ds, int maxv)
{
    static const int perm[4][13] = {{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12}, {9, 4, 10, 8, 1, 5, 11, 7, 3, 0, 2, 6, 12},
                                    {0, 3, 2, 1, 8, 7, 6, 5, 4, 9, 10, 11, 12}, {9, 8, 10, 4, 3, 7, 11, 5, 1, 0, 2, 6, 12}};
    int16_t f[13];
    for (int i = 0; i < 13; i++) f[i] = coef_final[(cl >> 2) & 0x1f][perm[cl & 3][i]];
    for (int r = by; r < by + 4; r++)
        for (int c = bx; c < bx + 4; c++) {
#define P(dy, dx) win(k, r + (dy), c + (dx))
            int sum = f[0] * (P(3, 0) + P(-3, 0))
                    + f[1] * (P(2, 1) + P(-2, -1)) + f[2] * (P(2, 0) + P(-2, 0)) + f[3] * (P(2, -1) + P(-2, 1))
                    + f[4] * (P(1, 2) + P(-1, -2)) + f[5] * (P(1, 1) + P(-1, -1)) + f[6] * (P(1, 0) + P(-1, 0)) + f[7] * (P(1, -1) + P(-1, 1)) + f[8] * (P(1, -2) + P(-1, 2))
                    + f[9] * (P(0, 3) + P(0, -3)) + f[10] * (P(0, 2) + P(0, -2)) + f[11] * (P(0, 1) + P(0, -1)) + f[12] * P(0, 0);
#undef P
            dst[(k->y0 + r) * ds + k->x0 + c] = (pel)orc_clip3(0, maxv, (sum + 256) >> 9);
        }
}

static void filter_chroma_ctu(const Win *k, const int16_t f[7], pel *dst, int ds, int maxv)
{
    for (int r = 0; r < k->h; r++)
        for (int c = 0; c < k->w; c++) {
#define P(dy, dx) win(k, r + (dy), c + (dx))
            int sum = f[0] * (P(2, 0) + P(-2, 0)) + f[1] * (P(1, 1) + P(-1, -1)) + f[2] * (P(1, 0) + P(-1, 0)) + f[3] * (P(1, -1) + P(-1, 1))
                    + f[4] * (P(0, 2) + P(0, -2)) + f[5] * (P(0, 1) + P(0, -1)) + f[6] * P(0, 0);
#undef P
            dst[(k->y0 + r) * ds + k->x0 + c] = (pel)orc_clip3(0, maxv, (sum + 256) >> 9);
        }
}

int orc_alf_frame(const XB200_PARAMS *prm, ORC_PIC *pic, const XB200_ALF *alf, const uint8_t *ctb_flag_luma)
{
    if (!alf->enable[0] && !alf->enable[1] && !alf->enable[2]) return XB200_OK;
    const int ctu = 1 << prm->log2_ctu, W = pic->w_l, H = pic->h_l, maxv = (1 << prm->bit_depth_luma) - 1;
    const int wc = (W + ctu - 1) / ctu;
    pel *cy = (pel *)malloc(sizeof(pel) * W * H), *cu = (pel *)malloc(sizeof(pel) * (W / 2) * (H / 2)), *cv = (pel *)malloc(sizeof(pel) * (W / 2) * (H / 2));
    for (int y = 0; y < H; y++) memcpy(cy + y * W, pic->y + y * pic->s_l, sizeof(pel) * W);
    for (int y = 0; y < H / 2; y++) {
        memcpy(cu + y * (W / 2), pic->u + y * pic->s_c, sizeof(pel) * (W / 2));
        memcpy(cv + y * (W / 2), pic->v + y * pic->s_c, sizeof(pel) * (W / 2));
    }
    for (int y0 = 0, n = 0; y0 < H; y0 += ctu)
        for (int x0 = 0; x0 < W; x0 += ctu, n++) {
            const int w = orc_min(ctu, W - x0), h = orc_min(ctu, H - y0);
            const ORC_TILES *tl = orc_tiles();
            int tc = 0, tr = 0;
            while (tc + 1 < tl->n_cols && (x0 >> prm->log2_ctu) >= tl->col_bd[tc + 1]) tc++;
            while (tr + 1 < tl->n_rows && (y0 >> prm->log2_ctu) >= tl->row_bd[tr + 1]) tr++;
            const int tx0 = tl->col_bd[tc] << prm->log2_ctu, ty0 = tl->row_bd[tr] << prm->log2_ctu;
            const int tx1 = orc_min((int)tl->col_bd[tc + 1] << prm->log2_ctu, W), ty1 = orc_min((int)tl->row_bd[tr + 1] << prm->log2_ctu, H);
            Win k = {cy, W, W, H, x0, y0, w, h, 0, 0, 0, 0, tx0, tx1, ty0, ty1};
            if (tl->across) { k.aL = x0 != 0; k.aR = 1; k.aT = y0 != 0; k.aB = 1; }       /* x_pos + width is never width - 1 */
            else { k.aL = x0 != tx0; k.aR = x0 + w != tx1; k.aT = y0 != ty0; k.aB = y0 + h != ty1; }
            (void)wc;
            if (alf->enable[0] && (!ctb_flag_luma || ctb_flag_luma[n]))
                for (int by = 0; by < h; by += 4)
                    for (int bx = 0; bx < w; bx += 4)
                        filter_luma_blk(&k, by, bx, classify(&k, by, bx, prm->bit_depth_luma), alf->coef_luma, pic->y, pic->s_l, maxv);
            for (int c = 1; c < 3; c++) {
                if (!alf->enable[c]) continue;
                Win kc = {c == 1 ? cu : cv, W / 2, W / 2, H / 2, x0 / 2, y0 / 2, w / 2, h / 2, k.aL, k.aR, k.aT, k.aB, tx0 / 2, tx1 / 2, ty0 / 2, ty1 / 2};
                filter_chroma_ctu(&kc, alf->coef_chroma, c == 1 ? pic->u : pic->v, pic->s_c, maxv);
            }
        }
    free(cy); free(cu); free(cv);
    return XB200_OK;
}
