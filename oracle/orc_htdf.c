/*
 * Hadamard-transform domain filter (Main profile, tool_htdf).  TEST INFRASTRUCTURE ONLY (orc_common.h).
 * Restates xevdm_htdf / xevdm_htdf_skip_condition / filter_block_luma / xevdm_htdf_filter_block / read_table
 * (src_main/xevdm_recon.c:153-385).
 *
 * The reference slides a 2x2 window over the (w+2)x(h+2) block (the CU plus a one-sample ring), Hadamard-transforms it, shrinks
 * the three AC terms through a QP-dependent table, transforms back and accumulates the four outputs; every sample ends up as
 * the rounded average of the four windows that contain it.  It overwrites a sample only after its last window has been
 * processed, so all windows see unfiltered input: written here as a per-sample gather.
 */
#include <string.h>
#include <stdlib.h>
#include "orc_common.h"

static const uint8_t k_thr_log2[5] = { 6, 7, 7, 8, 8 };
static const uint8_t k_tbl[5][16] = {
    { 0, 0, 2,  6, 10, 14, 19, 23, 28, 32,  36,  41,  45,  49,  53,  57 },
    { 0, 0, 5, 12, 20, 29, 38, 47, 56, 65,  73,  82,  90,  98, 107, 115 },
    { 0, 0, 1,  4,  9, 16, 24, 32, 41, 50,  59,  68,  77,  86,  94, 103 },
    { 0, 0, 3,  9, 19, 32, 47, 64, 81, 99, 117, 135, 154, 179, 205, 230 },
    { 0, 0, 0,  2,  6, 11, 18, 27, 38, 51,  64,  96, 128, 160, 192, 224 },
};
const uint8_t *orc_htdf_table(int idx) { return k_tbl[idx]; }

/* read_table (:173-186): |z| below the threshold goes through the table, larger values pass */
static int shrink(int z, const uint8_t *tbl, int thr, int shift, int round)
{
    const int a = z < 0 ? -z : z;
    if (a >= thr) return z;
    const int v = tbl[((a + round) & thr) >> shift];
    return z < 0 ? -v : v;
}

/* the four outputs of the window whose top-left sample is t[0] (row stride s): out[k] for (0,0) (0,1) (1,0) (1,1) */
static void window(const pel *t, int s, const uint8_t *tbl, int thr, int shift, int round, int out[4])
{
    const int x0 = t[0], x1 = t[1], x2 = t[s], x3 = t[s + 1];
    const int y0 = x0 + x2, y1 = x1 + x3, y2 = x0 - x2, y3 = x1 - x3;
    const int z0 = y0 + y1;                                            /* DC is not filtered */
    const int z1 = shrink(y0 - y1, tbl, thr, shift, round), z2 = shrink(y2 + y3, tbl, thr, shift, round), z3 = shrink(y2 - y3, tbl, thr, shift, round);
    const int i0 = z0 + z2, i1 = z1 + z3, i2 = z0 - z2, i3 = z1 - z3;
    out[0] = (i0 + i1) >> 2; out[1] = (i0 - i1) >> 2; out[2] = (i2 + i3) >> 2; out[3] = (i2 - i3) >> 2;
}

/* avail: the AVAIL_* bits of xevd_get_avail_intra (src_base/xevd_util.c:689-745; bit numbers xevd_def.h:237-247) */
/* map_scu: the picture's map at the CU's first SCU, consulted when `constrained` (intra CU under pps.constrained_intra_pred_flag,
 * xevdm.c:1387): a left / right / upper ring sample comes from the picture only if that neighbour SCU is intra (MCU_GET_IF), else it
 * is replicated from the CU (xevdm_recon.c:313-368); the four corners do not take the test */
void orc_htdf(pel *rec, int s, int w, int h, int qp, int intra, int avail, int bit_depth, const uint32_t *map_scu, int w_scu, int constrained)
{
    /* xevdm_htdf_skip_condition (:271-297) */
    if (qp <= 17 || w * h < 64) return;
    const int mn = orc_min(w, h), mx = orc_max(w, h);
    if (mx >= 128) return;
    if (!intra) { if (mn >= 32) return; }
    else if (w == h && mn >= 32) qp -= 8;

    const int we = w + 2, he = h + 2;
    pel *t = (pel *)malloc(sizeof(pel) * we * he);
    const int le = (avail >> 1) & 1, up = avail & 1, ri = (avail >> 3) & 1;
    for (int i = 0; i < h; i++) {
        memcpy(t + (i + 1) * we + 1, rec + i * s, sizeof(pel) * w);
#define NB_INTRA(off) (!constrained || ((map_scu[off] >> 15) & 1))                 /* MCU_GET_IF */
        t[(i + 1) * we] = (le && NB_INTRA(-1 + (i >> 2) * w_scu)) ? rec[i * s - 1] : rec[i * s];
        t[(i + 1) * we + we - 1] = (ri && NB_INTRA((w >> 2) + (i >> 2) * w_scu)) ? rec[i * s + w] : rec[i * s + w - 1];
    }
    for (int j = 0; j < w; j++) {
        t[j + 1] = (up && NB_INTRA(-w_scu + (j >> 2))) ? rec[j - s] : rec[j];
#undef NB_INTRA
        t[(he - 1) * we + j + 1] = rec[(h - 1) * s + j];              /* the row below is never available */
    }
    t[0] = ((avail >> 5) & 1) ? rec[-1 - s] : rec[0];
    t[we - 1] = ((avail >> 6) & 1) ? rec[w - s] : rec[w - 1];
    t[we * (he - 1)] = ((avail >> 7) & 1) ? rec[-1 + h * s] : rec[(h - 1) * s];
    t[we - 1 + we * (he - 1)] = ((avail >> 8) & 1) ? rec[w + h * s] : rec[w - 1 + (h - 1) * s];

    int idx = (qp - 20 + 4) >> 3;
    idx = orc_clip3(0, 4, idx);
    const int lg = k_thr_log2[idx], shift = lg - 4, round = (1 << shift) >> 1, thr = (1 << lg) - (1 << shift);
    const int maxv = (1 << bit_depth) - 1;
    for (int i = 1; i <= h; i++)
        for (int j = 1; j <= w; j++) {
            int o[4];
            pel acc = 0;                                               /* the accumulator is a pel in the reference */
            window(t + (i - 1) * we + (j - 1), we, k_tbl[idx], thr, shift, round, o); acc = (pel)(acc + o[3]);
            window(t + (i - 1) * we + j, we, k_tbl[idx], thr, shift, round, o);       acc = (pel)(acc + o[2]);
            window(t + i * we + (j - 1), we, k_tbl[idx], thr, shift, round, o);       acc = (pel)(acc + o[1]);
            window(t + i * we + j, we, k_tbl[idx], thr, shift, round, o);             acc = (pel)(acc + o[0]);
            rec[(i - 1) * s + (j - 1)] = (pel)orc_clip3(0, maxv, (acc + 2) >> 2);
        }
    free(t);
}
