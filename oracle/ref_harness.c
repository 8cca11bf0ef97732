/*
 * ref_harness.c -- thin ctypes-friendly entry points into the UNMODIFIED reference decoder library
 * (compiled from /root/reference by oracle/Makefile into oracle/_ref/libxevd_ref.so).
 * TEST INFRASTRUCTURE ONLY: it pins the oracle restatement (oracle/orc_*.c) and provides the
 * "reference" CPU baseline (the dispatched AVX2/SSE path).
 *
 * This file contains no reference code: it only calls the reference's own non-static symbols
 * through the reference's own headers.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "xevdm_def.h"
#include "xevd_mc.h"
#include "xevdm_mc.h"
#include "xevd_itdq.h"
#include "xevdm_itdq.h"
#include "xevd_recon.h"
#include "xevdm_recon.h"
#include "xevd_tbl.h"
#include "xevdm_tbl.h"
#include "xevd_mc_sse.h"
#include "xevd_mc_avx.h"
#include "xevd_itdq_sse.h"
#include "xevd_itdq_avx.h"
#include "xevdm_itdq_avx.h"
#include "xevdm_itdq_sse.h"
#include "xevdm_mc_sse.h"
#include "xevd_recon_avx.h"
#include "xevd_recon_sse.h"
#include "xevd_dbk_sse.h"
#include "../include/xevd_b200.h"
#include "orc_common.h"

static XEVD_CTX *g_ctx;      /* zeroed context carrying only the function tables */
static int       g_impl = 2; /* 0 = plain C, 1 = SSE, 2 = AVX2 (what xevdm_platform_init picks on this box) */

/* mirrors the table wiring of xevdm_platform_init (src_main/xevdm.c:3388-3479) for a chosen ISA level */
int ref_set_impl(int impl)
{
    if (!g_ctx) {
        g_ctx = (XEVD_CTX *)calloc(1, sizeof(XEVDM_CTX));
        xevdm_init_multi_tbl();
        xevd_init_multi_inv_tbl();
    }
    g_impl = impl;
    if (impl == 2) {
        xevd_func_itrans = xevdm_itrans_map_tbl_sse;  xevdm_fn_itx = &xevdm_tbl_itx_avx;
        xevdm_func_dmvr_mc_l = xevdm_tbl_dmvr_mc_l_sse; xevdm_func_dmvr_mc_c = xevdm_tbl_dmvr_mc_c_sse;
        xevdm_func_bl_mc_l = xevdm_tbl_bl_mc_l_sse;
        xevd_func_mc_l = xevd_tbl_mc_l_avx;  xevd_func_mc_c = xevd_tbl_mc_c_avx;
        xevd_func_average_no_clip = xevd_average_16b_no_clip_sse;
        g_ctx->fn_itxb = &xevd_tbl_itxb_avx;  g_ctx->fn_dbk = &xevd_tbl_dbk_sse;  g_ctx->fn_dbk_chroma = &xevd_tbl_dbk_chroma_sse;
        g_ctx->fn_recon = xevd_recon_avx;
    } else if (impl == 1) {
        xevd_func_itrans = xevdm_itrans_map_tbl_sse;  xevdm_fn_itx = &xevdm_tbl_itx;
        xevdm_func_dmvr_mc_l = xevdm_tbl_dmvr_mc_l_sse; xevdm_func_dmvr_mc_c = xevdm_tbl_dmvr_mc_c_sse;
        xevdm_func_bl_mc_l = xevdm_tbl_bl_mc_l_sse;
        xevd_func_mc_l = xevd_tbl_mc_l_sse;  xevd_func_mc_c = xevd_tbl_mc_c_sse;
        xevd_func_average_no_clip = xevd_average_16b_no_clip_sse;
        g_ctx->fn_itxb = &xevd_tbl_itxb_sse;  g_ctx->fn_dbk = &xevd_tbl_dbk_sse;  g_ctx->fn_dbk_chroma = &xevd_tbl_dbk_chroma_sse;
        g_ctx->fn_recon = xevd_recon_sse;
    } else {
        xevd_func_itrans = xevdm_itrans_map_tbl;  xevdm_fn_itx = &xevdm_tbl_itx;
        xevdm_func_dmvr_mc_l = xevdm_tbl_dmvr_mc_l; xevdm_func_dmvr_mc_c = xevdm_tbl_dmvr_mc_c;
        xevdm_func_bl_mc_l = xevdm_tbl_bl_mc_l;
        xevd_func_mc_l = xevd_tbl_mc_l;  xevd_func_mc_c = xevd_tbl_mc_c;
        xevd_func_average_no_clip = xevd_average_16b_no_clip;
        g_ctx->fn_itxb = &xevd_tbl_itxb;  g_ctx->fn_dbk = &xevd_tbl_dbk;  g_ctx->fn_dbk_chroma = &xevd_tbl_dbk_chroma;
        g_ctx->fn_recon = xevd_recon;
    }
    return 0;
}

static void ensure_init(void) { if (!g_ctx) ref_set_impl(2); }

static void select_mc_tables(int main_tables)
{
    /* the reference flips these process-global pointers per call (xevdm_mc.c:1915-1924, T13) */
    tbl_mc_l_coeff = main_tables ? tbl_mc_l_coeff_main : xevd_tbl_mc_l_coeff;
    tbl_mc_c_coeff = main_tables ? tbl_mc_c_coeff_main : xevd_tbl_mc_c_coeff;
}

/* ---- tables ----------------------------------------------------------------------------------- */
int ref_get_dct2(int log2n, int8_t *out)
{
    const s8 *t = NULL;
    switch (log2n) {
    case 1: t = &xevd_tbl_tm2[0][0]; break;   case 2: t = &xevd_tbl_tm4[0][0]; break;
    case 3: t = &xevd_tbl_tm8[0][0]; break;   case 4: t = &xevd_tbl_tm16[0][0]; break;
    case 5: t = &xevd_tbl_tm32[0][0]; break;  case 6: t = &xevd_tbl_tm64[0][0]; break;
    default: return -1;
    }
    memcpy(out, t, (size_t)1 << (2 * log2n));
    return 0;
}

int ref_get_inv_ats(int dst7, int log2n, int16_t *out)
{
    const s16 *t = NULL;
    ensure_init();
    switch (log2n) {
    case 2: t = xevd_tbl_inv_tr4[dst7 ? DST7 : DCT8][0]; break;
    case 3: t = xevd_tbl_inv_tr8[dst7 ? DST7 : DCT8][0]; break;
    case 4: t = xevd_tbl_inv_tr16[dst7 ? DST7 : DCT8][0]; break;
    case 5: t = xevd_tbl_inv_tr32[dst7 ? DST7 : DCT8][0]; break;
    default: return -1;
    }
    memcpy(out, t, sizeof(s16) << (2 * log2n));
    return 0;
}

int ref_get_mc_taps(int main_tables, int16_t *luma /*16*8*/, int16_t *chroma /*32*4*/)
{
    memcpy(luma, main_tables ? tbl_mc_l_coeff_main : xevd_tbl_mc_l_coeff, sizeof(s16) * 16 * 8);
    memcpy(chroma, main_tables ? tbl_mc_c_coeff_main : xevd_tbl_mc_c_coeff, sizeof(s16) * 32 * 4);
    return 0;
}

/* ---- leaf kernels ------------------------------------------------------------------------------- */
void ref_mc_luma(const pel *ref, int s_ref, int gmv_x, int gmv_y, int ori_mv_x, int ori_mv_y,
                 pel *pred, int s_pred, int w, int h, int bit_depth, int main_tables)
{
    ensure_init();
    select_mc_tables(main_tables);
    xevd_mc_l(ori_mv_x, ori_mv_y, (pel *)ref, gmv_x, gmv_y, s_ref, s_pred, pred, w, h, bit_depth);
}

void ref_mc_chroma(const pel *ref, int s_ref, int gmv_x, int gmv_y, int ori_mv_x, int ori_mv_y,
                   pel *pred, int s_pred, int w, int h, int bit_depth, int main_tables)
{
    ensure_init();
    select_mc_tables(main_tables);
    xevd_mc_c(ori_mv_x, ori_mv_y, (pel *)ref, gmv_x, gmv_y, s_ref, s_pred, pred, w, h, bit_depth);
}

/* xevdm_itdq on one transform block (dequant + inverse transform), coefficient buffer 32-byte aligned inside */
void ref_itdq_block(int16_t *coef, int log2w, int log2h, int qp, int bit_depth, int iqt)
{
    s16 *buf;
    int n = 1 << (log2w + log2h);
    int scale;
    ensure_init();
    buf = (s16 *)aligned_alloc(64, sizeof(s16) * (n < 32 ? 32 : n));
    memcpy(buf, coef, sizeof(s16) * n);
    scale = (iqt ? xevd_tbl_dq_scale : xevd_tbl_dq_scale_b)[qp % 6] << (qp / 6);
    xevdm_itdq(g_ctx, buf, log2w, log2h, scale, iqt, 0, 0, bit_depth);
    memcpy(coef, buf, sizeof(s16) * n);
    free(buf);
}

/* ---- CU-level picture reconstruction: the reference's own per-CU calls --------- */
static void wrap_pic(const ORC_PIC *o, XEVD_PIC *p)
{
    memset(p, 0, sizeof(*p));
    p->y = o->y; p->u = o->u; p->v = o->v;
    p->s_l = o->s_l; p->s_c = o->s_c;
    p->w_l = o->w_l; p->h_l = o->h_l; p->w_c = o->w_c; p->h_c = o->h_c;
    p->pad_l = o->pad_l; p->pad_c = o->pad_c;
    p->poc = o->poc;
}

typedef struct { pel pred[REFP_NUM][N_C][MAX_CU_DIM]; s16 coef[N_C][MAX_CU_DIM]; } REF_SCRATCH;

int ref_recon_frame(const XB200_PARAMS *prm, ORC_PIC *cur,
                    const ORC_PIC *const *refs_l0, int n_l0, const ORC_PIC *const *refs_l1, int n_l1,
                    const XB200_CU *cus, int n_cu, const XB200_CU_EXT *ext, const int16_t *coef)
{
    XEVD_PIC  rp[2][XEVD_MAX_NUM_REF_PICS];
    XEVD_REFP refp[XEVD_MAX_NUM_REF_PICS][REFP_NUM];
    REF_SCRATCH *s;
    int n, l, i;
    (void)ext;
    ensure_init();
    s = (REF_SCRATCH *)aligned_alloc(64, (sizeof(REF_SCRATCH) + 63) & ~(size_t)63);
    memset(refp, 0, sizeof(refp));
    for (i = 0; i < n_l0; i++) { wrap_pic(refs_l0[i], &rp[0][i]); refp[i][REFP_0].pic = &rp[0][i]; refp[i][REFP_0].poc = rp[0][i].poc; }
    for (i = 0; i < n_l1; i++) { wrap_pic(refs_l1[i], &rp[1][i]); refp[i][REFP_1].pic = &rp[1][i]; refp[i][REFP_1].poc = rp[1][i].poc; }

    for (n = 0; n < n_cu; n++) {
        const XB200_CU *cu = &cus[n];
        const int w = 1 << cu->log2w, h = 1 << cu->log2h, cw = w >> 1, ch = h >> 1;
        const int16_t *c = coef + cu->coef_off;
        int is_coef[N_C], nnz_sub[N_C][MAX_SUB_TB_NUM];
        s8  refi[REFP_NUM] = { cu->refi[0], cu->refi[1] };
        s16 mv[REFP_NUM][MV_D] = { { cu->mv[0][0], cu->mv[0][1] }, { cu->mv[1][0], cu->mv[1][1] } };
        for (l = 0; l < N_C; l++) {
            int bits = (cu->cbf >> (4 * l)) & 15;
            is_coef[l] = bits != 0;
            for (i = 0; i < MAX_SUB_TB_NUM; i++) nnz_sub[l][i] = (bits >> i) & 1;
        }
        if (is_coef[Y_C]) { memcpy(s->coef[Y_C], c, sizeof(s16) * w * h); c += (w * h + 7) & ~7; }
        if (is_coef[U_C]) { memcpy(s->coef[U_C], c, sizeof(s16) * cw * ch); c += (cw * ch + 7) & ~7; }
        if (is_coef[V_C]) { memcpy(s->coef[V_C], c, sizeof(s16) * cw * ch); }
        if (cu->cbf)
            xevdm_sub_block_itdq(g_ctx, s->coef, cu->log2w, cu->log2h, cu->qp_y, cu->qp_u, cu->qp_v, is_coef, nnz_sub,
                                 prm->tool_iqt, 0, 0, 0, prm->bit_depth_luma, prm->chroma_format_idc);
        if (cu->mode != XB200_MODE_INTER) { free(s); return XB200_ERR_UNSUPPORTED; }
        if (prm->tool_admvp) {
            select_mc_tables(1);
        } else {
            select_mc_tables(0);
        }
        /* Baseline xevd_mc (src_base/xevd_mc.c:469); identical to xevdm_mc with DMVR off apart from the table switch */
        xevd_mc(cu->x, cu->y, prm->w, prm->h, w, h, refi, mv, refp, s->pred, prm->poc,
                prm->bit_depth_luma, prm->bit_depth_chroma, prm->chroma_format_idc);
        /* xevd_recon_yuv (src_base/xevd_recon.c:70-91) */
        g_ctx->fn_recon(s->coef[Y_C], s->pred[0][Y_C], is_coef[Y_C], w, h, cur->s_l, cur->y + cu->y * cur->s_l + cu->x, prm->bit_depth_luma);
        g_ctx->fn_recon(s->coef[U_C], s->pred[0][U_C], is_coef[U_C], cw, ch, cur->s_c, cur->u + (cu->y >> 1) * cur->s_c + (cu->x >> 1), prm->bit_depth_luma);
        g_ctx->fn_recon(s->coef[V_C], s->pred[0][V_C], is_coef[V_C], cw, ch, cur->s_c, cur->v + (cu->y >> 1) * cur->s_c + (cu->x >> 1), prm->bit_depth_luma);
    }
    free(s);
    return XB200_OK;
}

void ref_pad(ORC_PIC *pic)
{
    XEVD_PIC p;
    wrap_pic(pic, &p);
    xevd_picbuf_lc_expand(&p, pic->pad_l, pic->pad_c);
}
