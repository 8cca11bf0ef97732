/*
 * Inverse quantisation + inverse transform.  TEST INFRASTRUCTURE ONLY (see orc_common.h).
 * Restates src_base/xevd_itdq.c:472-621 (Baseline) and src_main/xevdm_itdq.c:698-887 (IQT, ATS).
 *
 * The reference evaluates the DCT-2 with partial butterflies; those are an exact factorisation of
 * the matrix product  out[n] = sum_k tm[k][n] * in[k], so the direct product below yields the same
 * integers (tests/test_oracle_vs_ref.py checks all 36 shapes against the compiled reference).
 *
 * Overflow convention: the second Baseline pass is accumulated in wrapping 32-bit
 * arithmetic, which is what the dispatched x86 path (AVX2/SSE `mullo_epi32`) does; the plain-C
 * reference uses 64-bit there.  They agree whenever the sum fits 32 bits, which holds for all
 * coefficients a conforming stream produces.
 */
#include <string.h>
#include <stdlib.h>
#include "orc_common.h"

static inline int16_t clip16(int64_t v) { return (int16_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

/* xevd_itdq prologue + xevd_dquant (xevd_itdq.c:480-517):
 *   shift = 6 - (15 - bd - ((log2w+log2h)>>1)) (+8 and x181 when log2w+log2h is odd, T5) */
void orc_dequant(int16_t *coef, int log2w, int log2h, int qp, int bit_depth, int iqt)
{
    const int scale = orc_dq_scale(qp, iqt);
    const int odd = (log2w + log2h) & 1;
    const int shift = 20 - 14 - (15 - bit_depth - ((log2w + log2h) >> 1)) + (odd ? 8 : 0);
    const int64_t offset = shift == 0 ? 0 : ((int64_t)1 << (shift - 1));
    const int64_t mul = (int64_t)scale * (odd ? 181 : 1);
    for (int i = 0; i < (1 << (log2w + log2h)); i++) coef[i] = clip16((coef[i] * mul + offset) >> shift);
}

/* xevd_itrans (xevd_itdq.c:472-477) / xevdm_itrans (xevdm_itdq.c:708-724).
 * Pass 1 transforms columns (length h), pass 2 rows (length w).
 *   Baseline: pass 1 shift 0 kept in s32; pass 2 shift 7 + 12-(bd-8), clipped to s16.
 *   IQT     : pass 1 shift 7 clipped to s16; pass 2 shift 12-(bd-8), clipped to s16. */
void orc_inv_dct2(int16_t *coef, int log2w, int log2h, int bit_depth, int iqt)
{
    const int w = 1 << log2w, h = 1 << log2h;
    const int8_t *tv = orc_dct2_matrix(log2h), *th = orc_dct2_matrix(log2w);
    int32_t *tmp = (int32_t *)malloc((size_t)w * h * sizeof(int32_t));
    const int sh1 = iqt ? 7 : 0;
    const int sh2 = iqt ? 12 - (bit_depth - 8) : 7 + 12 - (bit_depth - 8);
    for (int x = 0; x < w; x++)
        for (int y = 0; y < h; y++) {
            int32_t acc = 0;
            for (int k = 0; k < h; k++) acc += tv[k * h + y] * coef[k * w + x];
            if (iqt) acc = clip16(((int64_t)acc + (1 << (sh1 - 1))) >> sh1);
            tmp[y * w + x] = acc;
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint32_t acc = 0;                                  /* wrapping, see header comment */
            for (int k = 0; k < w; k++) acc += (uint32_t)(th[k * w + x] * tmp[y * w + k]);
            acc += 1u << (sh2 - 1);
            coef[y * w + x] = clip16((int32_t)acc >> sh2);
        }
    free(tmp);
}

void orc_itdq_block(int16_t *coef, int log2w, int log2h, int qp, int bit_depth, int iqt)
{
    orc_dequant(coef, log2w, log2h, qp, bit_depth, iqt);
    orc_inv_dct2(coef, log2w, log2h, bit_depth, iqt);
}

/* xevdm_it_MxN_ats_intra (xevdm_itdq.c:404-421) with the kernels xevdm_itrans_ats_intra_{DST7,DCT8}_B{4..32}
 * (:163-402): full matrix products, columns first (shift 7), rows second (shift 20 - bit_depth), each clipped to s16.
 * ats_mode = horizontal << 1 | vertical, 0 = DST-7, 1 = DCT-8 (xevd_tbl_tr_subset_intra, xevdm_tbl.c:51).  The
 * reference's skip_w / skip_h arguments only skip all-zero inputs and do not change any value. */
void orc_inv_ats(int16_t *coef, int log2w, int log2h, int bit_depth, int ats_mode)
{
    const int w = 1 << log2w, h = 1 << log2h;
    const int16_t *mv = orc_ats_matrix(!(ats_mode & 1), log2h), *mh = orc_ats_matrix(!(ats_mode >> 1), log2w);
    int16_t *tmp = (int16_t *)malloc((size_t)w * h * sizeof(int16_t));
    const int sh2 = 20 - bit_depth;
    for (int x = 0; x < w; x++)
        for (int y = 0; y < h; y++) {
            int acc = 0;
            for (int k = 0; k < h; k++) acc += mv[y * h + k] * coef[k * w + x];
            tmp[y * w + x] = clip16((acc + 64) >> 7);
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int acc = 0;
            for (int k = 0; k < w; k++) acc += mh[x * w + k] * tmp[y * w + k];
            coef[y * w + x] = clip16((acc + (1 << (sh2 - 1))) >> sh2);
        }
    free(tmp);
}

/* TU geometry of a CU whose residual is a sub-block transform (ats_inter): xevdm_get_tu_size / get_tu_pos_offset
 * (xevdm_util.c:3585-3634).  idx 1/3: vertical split, TU = left or right half/quarter; idx 2/4: horizontal split. */
void orc_ats_inter_tu(int ats, int log2w, int log2h, int *tlw, int *tlh, int *xoff, int *yoff)
{
    const int idx = XB200_ATS_INTER_IDX(ats), pos = XB200_ATS_INTER_POS(ats);
    *tlw = log2w; *tlh = log2h; *xoff = 0; *yoff = 0;
    if (idx == 0) return;
    const int quad = idx == 3 || idx == 4;
    if (idx == 2 || idx == 4) { *tlh = log2h - (quad ? 2 : 1); *yoff = pos ? (1 << log2h) - (1 << *tlh) : 0; }
    else { *tlw = log2w - (quad ? 2 : 1); *xoff = pos ? (1 << log2w) - (1 << *tlw) : 0; }
}

/* one plane of xevd_sub_block_itdq (xevd_itdq.c:544-621): CUs wider/taller than 64 luma samples are
 * cut into 64-sample (chroma: 32-sample) transform blocks, each gated by its own nnz_sub bit
 * (bit (j<<1)|i).  lmax = 6 for luma, 5 for 4:2:0 chroma. */
static void itdq_plane(int16_t *c, int log2w, int log2h, int lmax, int qp, int bits, int bit_depth, int iqt)
{
    const int lw = orc_min(log2w, lmax), lh = orc_min(log2h, lmax);
    const int nx = 1 << (log2w - lw), ny = 1 << (log2h - lh), stride = 1 << log2w;
    if (nx == 1 && ny == 1) {
        if (bits & 1) orc_itdq_block(c, lw, lh, qp, bit_depth, iqt);
        return;
    }
    int16_t *blk = (int16_t *)malloc(sizeof(int16_t) << (lw + lh));
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            if (!((bits >> ((j << 1) | i)) & 1)) continue;
            int16_t *src = c + (j << lh) * stride + (i << lw);
            for (int r = 0; r < (1 << lh); r++) memcpy(blk + (r << lw), src + r * stride, sizeof(int16_t) << lw);
            orc_itdq_block(blk, lw, lh, qp, bit_depth, iqt);
            for (int r = 0; r < (1 << lh); r++) memcpy(src + r * stride, blk + (r << lw), sizeof(int16_t) << lw);
        }
    free(blk);
}

/* note: every plane is dequantised/transformed with the LUMA bit depth (src_base/xevd.c:441-442) */
void orc_itdq_cu(const XB200_PARAMS *prm, const XB200_CU *cu, int16_t *cy, int16_t *cu_, int16_t *cv)
{
    const int bd = prm->bit_depth_luma, iqt = prm->tool_iqt;
    if (prm->tool_ats && cu->mode != XB200_MODE_IBC) {
        /* xevdm_sub_block_itdq with ATS (xevdm_itdq.c:790-887): the luma block of an ats_intra CU uses DST-7 / DCT-8 per
         * ats_mode; an ats_inter CU transforms only its sub-block TU - luma with the position-dependent DST-7 / DCT-8 pair
         * when the CU is at most 32x32 (xevdm_get_ats_inter_trs, xevdm_util.c:3636-3667), DCT-2 otherwise; chroma always DCT-2 */
        const int inter_idx = cu->mode == XB200_MODE_INTRA ? 0 : XB200_ATS_INTER_IDX(cu->ats);
        int ats_on = cu->mode == XB200_MODE_INTRA && (cu->flags & XB200_CUF_ATS_INTRA), ats_mode = cu->ats & 3;
        int tlw = cu->log2w, tlh = cu->log2h, xo, yo;
        if (inter_idx) {
            const int pos = XB200_ATS_INTER_POS(cu->ats);
            orc_ats_inter_tu(cu->ats, cu->log2w, cu->log2h, &tlw, &tlh, &xo, &yo);
            if (cu->log2w <= 5 && cu->log2h <= 5) {
                ats_on = 1;
                ats_mode = (inter_idx == 2 || inter_idx == 4) ? (pos == 0 ? 1 : 0) : ((pos == 0 ? 1 : 0) << 1);
            }
        }
        if (ats_on || inter_idx) {
            if (cu->cbf & 0x00f) {
                if (ats_on) { orc_dequant(cy, tlw, tlh, cu->qp_y, bd, iqt); orc_inv_ats(cy, tlw, tlh, bd, ats_mode); }
                else orc_itdq_block(cy, tlw, tlh, cu->qp_y, bd, iqt);
            }
            if (cu->cbf & 0x0f0) orc_itdq_block(cu_, tlw - 1, tlh - 1, cu->qp_u, bd, iqt);
            if (cu->cbf & 0xf00) orc_itdq_block(cv, tlw - 1, tlh - 1, cu->qp_v, bd, iqt);
            return;
        }
    }
    if (cu->cbf & 0x00f) itdq_plane(cy, cu->log2w, cu->log2h, 6, cu->qp_y, cu->cbf & 15, bd, iqt);
    if (cu->cbf & 0x0f0) itdq_plane(cu_, cu->log2w - 1, cu->log2h - 1, 5, cu->qp_u, (cu->cbf >> 4) & 15, bd, iqt);
    if (cu->cbf & 0xf00) itdq_plane(cv, cu->log2w - 1, cu->log2h - 1, 5, cu->qp_v, (cu->cbf >> 8) & 15, bd, iqt);
}
