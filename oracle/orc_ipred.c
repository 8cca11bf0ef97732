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
