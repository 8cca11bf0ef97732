/*
 * Constant tables of the reconstruction path, regenerated from their definitions.
 * TEST INFRASTRUCTURE ONLY (see orc_common.h).
 *
 *  - DCT-2 kernels xevd_tbl_tm2..tm64 (src_base/xevd_tbl.c:89-241): the EVC integer DCT is the
 *    orthonormal DCT-II scaled by 64*sqrt(N) and rounded to nearest, so it is regenerated from
 *    cos() here; tests/test_oracle_vs_ref.py checks every entry against the reference's arrays.
 *  - DST-7 / DCT-8 inverse kernels xevd_tbl_inv_tr* (src_main/xevdm_itdq.c:121-159): same double
 *    expression as the reference (T8).
 *  - interpolation taps (src_base/xevd_mc.c:80-134, src_main/xevdm_mc.c:121-175) and dequant
 *    scales (src_base/xevd_tbl.c:255-256): constants of the standard, listed here.
 */
#include <math.h>
#include <string.h>
#include "orc_common.h"

static int8_t  g_dct2[7][64 * 64];
static int16_t g_ats[2][7][64 * 64];
static int     g_ready;

static void build_tables(void)
{
    if (g_ready) return;
    for (int lg = 1; lg <= 6; lg++) {
        int n = 1 << lg;
        for (int k = 0; k < n; k++)
            for (int i = 0; i < n; i++) {
                double v = (k == 0) ? 64.0 : 64.0 * sqrt(2.0) * cos(M_PI * (2 * i + 1) * k / (2.0 * n));
                g_dct2[lg][k * n + i] = (int8_t)lround(v);
            }
    }
    /* inverse ATS kernels: inv[n][k] = fwd[k][n] (xevdm_itdq.c:121-159) */
    for (int lg = 1; lg <= 6; lg++) {
        int c = 1 << lg;
        const double s = sqrt((double)c) * 64;
        for (int k = 0; k < c; k++)
            for (int n = 0; n < c; n++) {
                double v = cos(M_PI * (k + 0.5) * (n + 0.5) / (c + 0.5)) * sqrt(2.0 / (c + 0.5));
                g_ats[0][lg][n * c + k] = (int16_t)(s * v + (v > 0 ? 0.5 : -0.5));
                v = sin(M_PI * (k + 0.5) * (n + 1) / (c + 0.5)) * sqrt(2.0 / (c + 0.5));
                g_ats[1][lg][n * c + k] = (int16_t)(s * v + (v > 0 ? 0.5 : -0.5));
            }
    }
    g_ready = 1;
}

const int8_t *orc_dct2_matrix(int log2n) { build_tables(); return g_dct2[log2n]; }
const int16_t *orc_ats_matrix(int dst7, int log2n) { build_tables(); return g_ats[dst7 ? 1 : 0][log2n]; }

/* quarter-pel luma phases of the Baseline profile; other sixteenth phases are all-zero rows */
static const int16_t k_luma_base[16][8] = {
    [0]  = {0, 0, 0, 64, 0, 0, 0, 0},
    [4]  = {0, 1, -5, 52, 20, -5, 1, 0},
    [8]  = {0, 2, -10, 40, 40, -10, 2, 0},
    [12] = {0, 1, -5, 20, 52, -5, 1, 0},
};
static const int16_t k_luma_main[16][8] = {
    {0, 0, 0, 64, 0, 0, 0, 0},       {0, 1, -3, 63, 4, -2, 1, 0},     {-1, 2, -5, 62, 8, -3, 1, 0},
    {-1, 3, -8, 60, 13, -4, 1, 0},   {-1, 4, -10, 58, 17, -5, 1, 0},  {-1, 4, -11, 52, 26, -8, 3, -1},
    {-1, 3, -9, 47, 31, -10, 4, -1}, {-1, 4, -11, 45, 34, -10, 4, -1},{-1, 4, -11, 40, 40, -11, 4, -1},
    {-1, 4, -10, 34, 45, -11, 4, -1},{-1, 4, -10, 31, 47, -9, 3, -1}, {-1, 3, -8, 26, 52, -11, 4, -1},
    {0, 1, -5, 17, 58, -10, 4, -1},  {0, 1, -4, 13, 60, -8, 3, -1},   {0, 1, -3, 8, 62, -5, 2, -1},
    {0, 1, -2, 4, 63, -3, 1, 0},
};
/* Baseline chroma: eighth-pel phases (every 4th of 32) */
static const int16_t k_chroma_base[32][4] = {
    [0]  = {0, 64, 0, 0},   [4]  = {-2, 58, 10, -2}, [8]  = {-4, 52, 20, -4}, [12] = {-6, 46, 30, -6},
    [16] = {-8, 40, 40, -8},[20] = {-6, 30, 46, -6}, [24] = {-4, 20, 52, -4}, [28] = {-2, 10, 58, -2},
};
static const int16_t k_chroma_main[32][4] = {
    {0, 64, 0, 0},   {-1, 63, 2, 0},  {-2, 62, 4, 0},  {-2, 60, 7, -1}, {-2, 58, 10, -2}, {-3, 57, 12, -2},
    {-4, 56, 14, -2},{-4, 55, 15, -2},{-4, 54, 16, -2},{-5, 53, 18, -2},{-6, 52, 20, -2}, {-6, 49, 24, -3},
    {-6, 46, 28, -4},{-5, 44, 29, -4},{-4, 42, 30, -4},{-4, 39, 33, -4},{-4, 36, 36, -4}, {-4, 33, 39, -4},
    {-4, 30, 42, -4},{-4, 29, 44, -5},{-4, 28, 46, -6},{-3, 24, 49, -6},{-2, 20, 52, -6}, {-2, 18, 53, -5},
    {-2, 16, 54, -4},{-2, 15, 55, -4},{-2, 14, 56, -4},{-2, 12, 57, -3},{-2, 10, 58, -2}, {-1, 7, 60, -2},
    {0, 4, 62, -2},  {0, 2, 63, -1},
};

const int16_t *orc_mc_luma_taps(int main_tables)   { return main_tables ? &k_luma_main[0][0] : &k_luma_base[0][0]; }
const int16_t *orc_mc_chroma_taps(int main_tables) { return main_tables ? &k_chroma_main[0][0] : &k_chroma_base[0][0]; }

/* xevd_tbl_dq_scale / xevd_tbl_dq_scale_b (xevd_tbl.c:255-256); scale = tbl[qp%6] << (qp/6)
 * (xevd_itdq.c:588, xevdm_itdq.c:854-861) */
int orc_dq_scale(int qp, int iqt)
{
    static const int base[6] = {40, 45, 51, 57, 64, 71};
    int s = base[qp % 6];
    if (iqt && qp % 6 == 5) s = 72;
    return s << (qp / 6);
}
