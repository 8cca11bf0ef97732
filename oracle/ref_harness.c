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
#include "xevd_df.h"
#include "xevdm_alf.h"
#include "xevd_ipred.h"
#include "xevdm_ipred.h"
#include "xevd_util.h"
#include "xevdm_df.h"
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

/* batched forms of the two leaf calls: the CPU baseline of BASELINE config 5 (bench.py), no per-block allocation or Python overhead inside
 * the timed loop.  coef: n contiguous blocks, transformed in place.  mv: int32[n][4] = {gmv_x, gmv_y, ori_mv_x, ori_mv_y} as xb200_mc_blocks */
void ref_itdq_blocks(int16_t *coef, int n, int log2w, int log2h, int qp, int bit_depth, int iqt)
{
    const int sz = 1 << (log2w + log2h);
    int i, scale;
    s16 *buf;
    ensure_init();
    buf = (s16 *)aligned_alloc(64, sizeof(s16) * (sz < 32 ? 32 : sz));
    scale = (iqt ? xevd_tbl_dq_scale : xevd_tbl_dq_scale_b)[qp % 6] << (qp / 6);
    for (i = 0; i < n; i++) {
        memcpy(buf, coef + (size_t)i * sz, sizeof(s16) * sz);
        xevdm_itdq(g_ctx, buf, log2w, log2h, scale, iqt, 0, 0, bit_depth);
        memcpy(coef + (size_t)i * sz, buf, sizeof(s16) * sz);
    }
    free(buf);
}
void ref_mc_blocks(const pel *ref, int s_ref, int chroma, const int *mv, pel *out, int n, int w, int h, int bit_depth, int main_tables)
{
    int i;
    ensure_init();
    select_mc_tables(main_tables);
    for (i = 0; i < n; i++) {
        const int *m = mv + 4 * i;
        if (chroma) xevd_mc_c(m[2], m[3], (pel *)ref, m[0], m[1], s_ref, w, out + (size_t)i * w * h, w, h, bit_depth);
        else        xevd_mc_l(m[2], m[3], (pel *)ref, m[0], m[1], s_ref, w, out + (size_t)i * w * h, w, h, bit_depth);
    }
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

typedef struct { pel pred[REFP_NUM][N_C][MAX_CU_DIM]; s16 coef[N_C][MAX_CU_DIM]; pel nb[N_C][N_REF][MAX_CU_SIZE * 3]; } REF_SCRATCH;

/* ---- per-SCU map publication through the reference's OWN xevdm_set_dec_info (src_main/xevdm_util.c:4205-4389) ---------------
 * A zeroed XEVDM_CTX / XEVDM_CORE pair carries exactly the fields that function reads; the core is filled from the work item the way
 * cu_init (src_main/xevdm.c:1022) and the motion derivation leave it.  Nothing here writes a map entry itself. */
typedef struct {
    XEVDM_CTX *m; XEVDM_CORE *k; XEVD_SPS sps;
    s8 *map_ipm; u32 *map_affine, *map_cu_mode; u8 *map_ats_inter; s16 *own_unref;
} REF_INFO;
static void info_open(REF_INFO *I, const XB200_PARAMS *prm, ORC_PIC *cur, u32 *map_scu)
{
    const int f_scu = cur->w_scu * cur->h_scu;
    memset(I, 0, sizeof(*I));
    I->m = (XEVDM_CTX *)calloc(1, sizeof(XEVDM_CTX));
    I->k = (XEVDM_CORE *)calloc(1, sizeof(XEVDM_CORE));
    I->map_ipm = (s8 *)calloc(f_scu, 1); I->map_affine = (u32 *)calloc(f_scu, 4); I->map_cu_mode = (u32 *)calloc(f_scu, 4);
    I->map_ats_inter = (u8 *)calloc(f_scu, 1);
    I->sps.chroma_format_idc = prm->chroma_format_idc;
    XEVD_CTX *c = &I->m->bctx;
    c->sps = &I->sps;
    c->w_scu = cur->w_scu; c->h_scu = cur->h_scu;
    c->map_scu = map_scu;
    c->map_refi = (s8(*)[REFP_NUM])cur->map_refi;
    c->map_mv = (s16(*)[REFP_NUM][MV_D])cur->map_mv;
    if (!cur->map_unrefined_mv) I->own_unref = (s16 *)calloc(f_scu, 8);
    I->m->map_unrefined_mv = (s16(*)[REFP_NUM][MV_D])(cur->map_unrefined_mv ? cur->map_unrefined_mv : I->own_unref);
    c->map_ipm = I->map_ipm; c->map_cu_mode = I->map_cu_mode;
    I->m->map_affine = I->map_affine; I->m->map_ats_inter = I->map_ats_inter;
    c->pps.cu_qp_delta_enabled_flag = 1;          /* map_scu takes core->qp (xevdm_util.c:4303-4306) = XB200_CU.qp_map */
    c->slice_num = 0;
}
static void info_close(REF_INFO *I)
{
    free(I->m); free(I->k); free(I->map_ipm); free(I->map_affine); free(I->map_cu_mode); free(I->map_ats_inter); free(I->own_unref);
}
/* dmvr_mv / dmvr_flag: what xevdm_mc returned for this CU (NULL / 0 otherwise) */
static void info_publish(REF_INFO *I, const XB200_PARAMS *prm, ORC_PIC *cur, const XB200_CU *cu, const XB200_CU_EXT *ext,
                         int dmvr_flag, s16 (*dmvr_mv)[REFP_NUM][MV_D])
{
    XEVD_CORE *k = &I->k->core;
    XEVDM_CORE *mk = I->k;
    const int do_l = (cu->flags & XB200_CUF_LUMA) != 0, do_c = (cu->flags & XB200_CUF_CHROMA) != 0;
    int l, i;
    if (!cur->map_mv || !cur->map_refi) return;
    mk->tree_cons = (TREE_CONS){ FALSE, do_l && do_c ? TREE_LC : (do_l ? TREE_L : TREE_C), do_l && do_c ? eAll : eOnlyIntra };
    k->scup = (cu->y >> 2) * cur->w_scu + (cu->x >> 2);
    k->log2_cuw = cu->log2w; k->log2_cuh = cu->log2h;
    k->qp = cu->qp_map;
    mk->affine_flag = 0; mk->ibc_flag = 0; mk->mmvd_flag = 0; mk->dmvr_flag = 0; mk->ats_inter_info = 0;
    memset(k->mv, 0, sizeof(k->mv));
    k->refi[0] = k->refi[1] = -1;
    k->ipm[0] = k->ipm[1] = 0;
    for (l = 0; l < N_C; l++) {
        const int bits = (cu->cbf >> (4 * l)) & 15;
        k->is_coef[l] = bits != 0;
        for (i = 0; i < MAX_SUB_TB_NUM; i++) k->is_coef_sub[l][i] = (bits >> i) & 1;
    }
    if (cu->mode == XB200_MODE_INTRA) {
        k->pred_mode = MODE_INTRA;
        k->ipm[0] = cu->refi[0]; k->ipm[1] = cu->refi[1];
    } else if (cu->mode == XB200_MODE_IBC) {
        k->pred_mode = MODE_IBC; mk->ibc_flag = 1;
        k->mv[0][0] = cu->mv[0][0]; k->mv[0][1] = cu->mv[0][1];
    } else {
        /* DMVR is only ever enabled for skip and direct-mode CUs (xevdm.c:1273-1288) */
        k->pred_mode = (cu->flags & XB200_CUF_SKIP) ? MODE_SKIP : ((cu->flags & XB200_CUF_DMVR) ? MODE_DIR : MODE_INTER);
        k->refi[0] = cu->refi[0]; k->refi[1] = cu->refi[1];
        if (prm->tool_ats) mk->ats_inter_info = get_ats_inter_info(XB200_ATS_INTER_IDX(cu->ats), XB200_ATS_INTER_POS(cu->ats));
        if (cu->mode == XB200_MODE_AFFINE) {
            uint32_t ei; int v;
            memcpy(&ei, cu->mv[1], 4);
            mk->affine_flag = (cu->flags & XB200_CUF_AFF6) ? 2 : 1;
            memset(mk->affine_mv, 0, sizeof(mk->affine_mv));
            for (l = 0; l < 2; l++) {
                for (v = 0; v < 3; v++) { mk->affine_mv[l][v][0] = ext[ei].u.affine.cp[l][v][0]; mk->affine_mv[l][v][1] = ext[ei].u.affine.cp[l][v][1]; }
                k->mv[l][0] = ext[ei].u.affine.mv_unref[l][0]; k->mv[l][1] = ext[ei].u.affine.mv_unref[l][1];
            }
        } else {
            for (l = 0; l < 2; l++) { k->mv[l][0] = cu->mv[l][0]; k->mv[l][1] = cu->mv[l][1]; }
            if (dmvr_flag) {
                mk->dmvr_flag = 1;
                memcpy(mk->dmvr_mv, dmvr_mv, sizeof(mk->dmvr_mv));
            }
        }
    }
    xevdm_set_dec_info(&I->m->bctx, k);
}

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
    /* decoding-order state the intra path reads: COD bits of map_scu (set as each CU is reconstructed), one tile */
    const int f_scu = cur->w_scu * cur->h_scu;
    u32 *map_scu = (u32 *)calloc(f_scu, sizeof(u32));
    u8 *map_tidx = (u8 *)calloc(f_scu, 1);

    /* the intra flags are in map_scu from the parsing pass of the CTU (xevdm_set_dec_info in xevd_entropy_dec_unit), i.e. before any CU
     * of it is reconstructed; only luma-carrying CUs publish (xevdm_util.c:4241).  Read under pps.constrained_intra_pred_flag. */
    REF_INFO info;
    info_open(&info, prm, cur, map_scu);
    for (n = 0; n < n_cu; n++)          /* the parsing pass: xevdm_set_dec_info of every CU (intra flags, QP, modes); inter CUs again after MC below */
        info_publish(&info, prm, cur, &cus[n], ext, 0, NULL);
    for (n = 0; n < n_cu; n++) {
        const XB200_CU *cu = &cus[n];
        const int w = 1 << cu->log2w, h = 1 << cu->log2h, cw = w >> 1, ch = h >> 1;
        const int16_t *c = coef + cu->coef_off;
        const int cip = cu->mode == XB200_MODE_INTRA && prm->constrained_intra_pred;       /* constrained_intra_flag (xevdm.c:609,1387) */
        int is_coef[N_C], nnz_sub[N_C][MAX_SUB_TB_NUM];
        s8  refi[REFP_NUM] = { cu->refi[0], cu->refi[1] };
        s16 mv[REFP_NUM][MV_D] = { { cu->mv[0][0], cu->mv[0][1] }, { cu->mv[1][0], cu->mv[1][1] } };
        for (l = 0; l < N_C; l++) {
            int bits = (cu->cbf >> (4 * l)) & 15;
            is_coef[l] = bits != 0;
            for (i = 0; i < MAX_SUB_TB_NUM; i++) nnz_sub[l][i] = (bits >> i) & 1;
        }
        /* ATS syntax elements as xevdm_itdq_main passes them (src_main/xevdm.c:599-603) */
        const int ats_on = prm->tool_ats && cu->mode != XB200_MODE_IBC;
        const u8 ats_intra_cu = (ats_on && cu->mode == XB200_MODE_INTRA && (cu->flags & XB200_CUF_ATS_INTRA)) ? 1 : 0;
        const u8 ats_mode = ats_intra_cu ? (cu->ats & 3) : 0;
        const u8 ats_inter_info = (ats_on && cu->mode != XB200_MODE_INTRA) ? get_ats_inter_info(XB200_ATS_INTER_IDX(cu->ats), XB200_ATS_INTER_POS(cu->ats)) : 0;
        int tlw = cu->log2w, tlh = cu->log2h;
        if (ats_inter_info) xevdm_get_tu_size(ats_inter_info, cu->log2w, cu->log2h, &tlw, &tlh);
        const int tn = 1 << (tlw + tlh);
        if (is_coef[Y_C]) { memcpy(s->coef[Y_C], c, sizeof(s16) * tn); c += (tn + 7) & ~7; }
        if (is_coef[U_C]) { memcpy(s->coef[U_C], c, sizeof(s16) * (tn / 4)); c += (tn / 4 + 7) & ~7; }
        if (is_coef[V_C]) { memcpy(s->coef[V_C], c, sizeof(s16) * (tn / 4)); }
        if (cu->cbf)
            xevdm_sub_block_itdq(g_ctx, s->coef, cu->log2w, cu->log2h, cu->qp_y, cu->qp_u, cu->qp_v, is_coef, nnz_sub,
                                 prm->tool_iqt, ats_intra_cu, ats_mode, ats_inter_info, prm->bit_depth_luma, prm->chroma_format_idc);
        const int scup = (cu->y >> 2) * cur->w_scu + (cu->x >> 2);
        /* local dual tree (src_main/xevdm.c:1828-1846,1908-1927): TREE_L CUs carry luma only, the TREE_C CU that follows them chroma only */
        const int do_l = (cu->flags & XB200_CUF_LUMA) != 0, do_c = (cu->flags & XB200_CUF_CHROMA) != 0;
        const TREE_CONS tcu = { FALSE, do_l && do_c ? TREE_LC : (do_l ? TREE_L : TREE_C), do_l && do_c ? eAll : eOnlyIntra };
        if (!do_l && !do_c) { free(s); free(map_scu); free(map_tidx); return XB200_ERR_INVALID_ARGUMENT; }
        if (cu->mode == XB200_MODE_INTRA && !prm->tool_eipd) {
            /* xevd_recon_unit intra branch (src_base/xevd.c:732-741) with the reference's own availability logic */
            const u16 avail_cu = xevd_get_avail_intra(cu->x >> 2, cu->y >> 2, cur->w_scu, cur->h_scu, scup, cu->log2w, cu->log2h, map_scu, map_tidx);
            const int bdl = prm->bit_depth_luma;
            /* get_nbr_yuv and the prediction calls are gated by xevd_check_luma / xevd_check_chroma (src_main/xevdm.c:611-640,1362-1376) */
            if (do_l) xevd_get_nbr_b(cu->x, cu->y, w, h, cur->y + cu->y * cur->s_l + cu->x, cur->s_l, avail_cu, s->nb, scup, map_scu, cur->w_scu, cur->h_scu,
                           Y_C, cip, map_tidx, bdl, 1);
            if (do_c) xevd_get_nbr_b(cu->x >> 1, cu->y >> 1, cw, ch, cur->u + (cu->y >> 1) * cur->s_c + (cu->x >> 1), cur->s_c, avail_cu, s->nb, scup, map_scu,
                           cur->w_scu, cur->h_scu, U_C, cip, map_tidx, bdl, 1);
            if (do_c) xevd_get_nbr_b(cu->x >> 1, cu->y >> 1, cw, ch, cur->v + (cu->y >> 1) * cur->s_c + (cu->x >> 1), cur->s_c, avail_cu, s->nb, scup, map_scu,
                           cur->w_scu, cur->h_scu, V_C, cip, map_tidx, bdl, 1);
            if (do_l) xevd_ipred_b(s->nb[0][0] + 2, s->nb[0][1] + h, s->nb[0][2] + 2, 0, s->pred[0][Y_C], cu->refi[0], w, h);
            if (do_c) xevd_ipred_uv_b(s->nb[1][0] + 2, s->nb[1][1] + ch, s->nb[1][2] + 2, 0, s->pred[0][U_C], cu->refi[1], cu->refi[0], cw, ch);
            if (do_c) xevd_ipred_uv_b(s->nb[2][0] + 2, s->nb[2][1] + ch, s->nb[2][2] + 2, 0, s->pred[0][V_C], cu->refi[1], cu->refi[0], cw, ch);
        } else if (cu->mode == XB200_MODE_IBC) {
            XEVD_PIC cp;
            wrap_pic(cur, &cp);
            xevdm_IBC_mc(cu->x, cu->y, cu->log2w, cu->log2h, mv[0], &cp, s->pred[0], tcu, prm->chroma_format_idc);
        } else if (cu->mode == XB200_MODE_INTRA) {
            /* Main-profile intra branch of xevd_recon_unit (src_main/xevdm.c:1344-1361) with the reference's own availability logic */
            const u16 avail_cu = xevd_get_avail_intra(cu->x >> 2, cu->y >> 2, cur->w_scu, cur->h_scu, scup, cu->log2w, cu->log2h, map_scu, map_tidx);
            const u16 avail_lr = xevd_check_nev_avail(cu->x >> 2, cu->y >> 2, w, h, cur->w_scu, cur->h_scu, map_scu, map_tidx);
            const int bdl = prm->bit_depth_luma, bdc = prm->bit_depth_chroma;
            if (do_l) xevdm_get_nbr(cu->x, cu->y, w, h, cur->y + cu->y * cur->s_l + cu->x, cur->s_l, avail_cu, s->nb, scup, map_scu, cur->w_scu, cur->h_scu,
                          Y_C, cip, map_tidx, bdl, 1);
            if (do_c) xevdm_get_nbr(cu->x >> 1, cu->y >> 1, cw, ch, cur->u + (cu->y >> 1) * cur->s_c + (cu->x >> 1), cur->s_c, avail_cu, s->nb, scup, map_scu,
                          cur->w_scu, cur->h_scu, U_C, cip, map_tidx, bdl, 1);
            if (do_c) xevdm_get_nbr(cu->x >> 1, cu->y >> 1, cw, ch, cur->v + (cu->y >> 1) * cur->s_c + (cu->x >> 1), cur->s_c, avail_cu, s->nb, scup, map_scu,
                          cur->w_scu, cur->h_scu, V_C, cip, map_tidx, bdl, 1);
            if (do_l) xevdm_ipred(s->nb[0][0] + 2, s->nb[0][1] + h, s->nb[0][2] + 2, avail_lr, s->pred[0][Y_C], cu->refi[0], w, h, bdl);
            if (do_c) xevdm_ipred_uv(s->nb[1][0] + 2, s->nb[1][1] + ch, s->nb[1][2] + 2, avail_lr, s->pred[0][U_C], cu->refi[1], cu->refi[0], cw, ch, bdc);
            if (do_c) xevdm_ipred_uv(s->nb[2][0] + 2, s->nb[2][1] + ch, s->nb[2][2] + 2, avail_lr, s->pred[0][V_C], cu->refi[1], cu->refi[0], cw, ch, bdc);
        } else if (cu->mode == XB200_MODE_AFFINE) {
            /* xevdm_affine_mc (src_main/xevdm_mc.c:2606) + the per-SCU vectors of xevdm_set_affine_mvf (src_main/xevdm_util.c:4095) */
            static pel eif_tmp[(MAX_CU_SIZE + 2) * (MAX_CU_SIZE + 2)];
            uint32_t ei;
            s16 amv[REFP_NUM][VER_NUM][MV_D];
            int v;
            memcpy(&ei, cu->mv[1], 4);
            memset(amv, 0, sizeof(amv));
            for (l = 0; l < 2; l++) for (v = 0; v < 3; v++) { amv[l][v][0] = ext[ei].u.affine.cp[l][v][0]; amv[l][v][1] = ext[ei].u.affine.cp[l][v][1]; }
            select_mc_tables(prm->tool_admvp ? 1 : 0);
            xevdm_affine_mc(cu->x, cu->y, prm->w, prm->h, w, h, refi, amv, refp, s->pred, (cu->flags & XB200_CUF_AFF6) ? 3 : 2, eif_tmp,
                            prm->bit_depth_luma, prm->bit_depth_chroma, prm->chroma_format_idc);
            info_publish(&info, prm, cur, cu, ext, 0, NULL);          /* xevdm_set_dec_info incl. xevdm_set_affine_mvf (src_main/xevdm.c:1333) */
        } else if (cu->mode == XB200_MODE_INTER && prm->tool_dmvr) {
            /* Main xevdm_mc with the DMVR scratch buffers of XEVDM_CORE (src_main/xevdm_def.h:505-530); it leaves the averaged
             * prediction in pred[0], the refined vectors per SCU in dmvr_mv and restores mv[] */
            static pel tmpl[MAX_CU_DIM];
            static pel interp[REFP_NUM][(MAX_CU_SIZE + ((DMVR_NEW_VERSION_ITER_COUNT + 1) * REF_PRED_EXTENTION_PEL_COUNT)) * (MAX_CU_SIZE + ((DMVR_NEW_VERSION_ITER_COUNT + 1) * REF_PRED_EXTENTION_PEL_COUNT))];
            static pel halfp[REFP_NUM][(MAX_CU_SIZE + 1) * (MAX_CU_SIZE + 1)];
            static pel padbuf[REFP_NUM][N_C][PAD_BUFFER_STRIDE * PAD_BUFFER_STRIDE];
            static s16 dmvr_mv[MAX_CU_CNT_IN_LCU][REFP_NUM][MV_D];
            u8 dmvr_flag = 0;
            xevdm_mc(cu->x, cu->y, prm->w, prm->h, w, h, refi, mv, refp, s->pred, prm->poc, tmpl, interp, halfp, (cu->flags & XB200_CUF_DMVR) ? 1 : 0,
                     padbuf, &dmvr_flag, dmvr_mv, prm->tool_admvp, prm->bit_depth_luma, prm->bit_depth_chroma, prm->chroma_format_idc);
            info_publish(&info, prm, cur, cu, ext, dmvr_flag, dmvr_mv);
        } else if (cu->mode == XB200_MODE_INTER) {
            select_mc_tables(prm->tool_admvp ? 1 : 0);
            /* Baseline xevd_mc (src_base/xevd_mc.c:469); identical to xevdm_mc with DMVR off apart from the table switch */
            xevd_mc(cu->x, cu->y, prm->w, prm->h, w, h, refi, mv, refp, s->pred, prm->poc,
                    prm->bit_depth_luma, prm->bit_depth_chroma, prm->chroma_format_idc);
            info_publish(&info, prm, cur, cu, ext, 0, NULL);
        } else { free(s); free(map_scu); free(map_tidx); return XB200_ERR_UNSUPPORTED; }
        for (l = 0; l < (h >> 2); l++)
            for (i = 0; i < (w >> 2); i++) MCU_SET_COD(map_scu[scup + l * cur->w_scu + i]);
        if (prm->tool_ats || prm->tool_htdf) {
            /* Main profile: xevdm_recon_yuv (src_main/xevdm_recon.c:128-151), which places the ats_inter TU */
            if (do_l) xevdm_recon(s->coef[Y_C], s->pred[0][Y_C], is_coef[Y_C], w, h, cur->s_l, cur->y + cu->y * cur->s_l + cu->x, ats_inter_info, prm->bit_depth_luma);
            if (do_c) xevdm_recon(s->coef[U_C], s->pred[0][U_C], is_coef[U_C], cw, ch, cur->s_c, cur->u + (cu->y >> 1) * cur->s_c + (cu->x >> 1), ats_inter_info, prm->bit_depth_luma);
            if (do_c) xevdm_recon(s->coef[V_C], s->pred[0][V_C], is_coef[V_C], cw, ch, cur->s_c, cur->v + (cu->y >> 1) * cur->s_c + (cu->x >> 1), ats_inter_info, prm->bit_depth_luma);
            if (cu->mode != XB200_MODE_IBC && prm->tool_htdf && (is_coef[Y_C] || cu->mode == XB200_MODE_INTRA) && do_l) {
                /* src_main/xevdm.c:1381-1391; the COD bits of this CU are cleared around the call as they are in the decoder */
                u16 av;
                for (l = 0; l < (h >> 2); l++) for (i = 0; i < (w >> 2); i++) MCU_CLR_COD(map_scu[scup + l * cur->w_scu + i]);
                av = xevd_get_avail_intra(cu->x >> 2, cu->y >> 2, cur->w_scu, cur->h_scu, scup, cu->log2w, cu->log2h, map_scu, map_tidx);
                xevdm_htdf(cur->y + cu->y * cur->s_l + cu->x, prm->slice_qp, w, h, cur->s_l, cu->mode == XB200_MODE_INTRA,
                           cur->y + cu->y * cur->s_l + cu->x, cur->s_l, av, scup, cur->w_scu, cur->h_scu, map_scu, cip, prm->bit_depth_luma);
                for (l = 0; l < (h >> 2); l++) for (i = 0; i < (w >> 2); i++) MCU_SET_COD(map_scu[scup + l * cur->w_scu + i]);
            }
            continue;
        }
        /* xevd_recon_yuv (src_base/xevd_recon.c:70-91) */
        if (do_l) g_ctx->fn_recon(s->coef[Y_C], s->pred[0][Y_C], is_coef[Y_C], w, h, cur->s_l, cur->y + cu->y * cur->s_l + cu->x, prm->bit_depth_luma);
        if (do_c) g_ctx->fn_recon(s->coef[U_C], s->pred[0][U_C], is_coef[U_C], cw, ch, cur->s_c, cur->u + (cu->y >> 1) * cur->s_c + (cu->x >> 1), prm->bit_depth_luma);
        if (do_c) g_ctx->fn_recon(s->coef[V_C], s->pred[0][V_C], is_coef[V_C], cw, ch, cur->s_c, cur->v + (cu->y >> 1) * cur->s_c + (cu->x >> 1), prm->bit_depth_luma);
    }
    if (cur->map_scu) memcpy(cur->map_scu, map_scu, sizeof(u32) * f_scu);
    info_close(&info);
    free(s); free(map_scu); free(map_tidx);
    return XB200_OK;
}

void ref_pad(ORC_PIC *pic)
{
    XEVD_PIC p;
    wrap_pic(pic, &p);
    xevd_picbuf_lc_expand(&p, pic->pad_l, pic->pad_c);
}

/* ---- deblocking, Baseline filter through the Main library's CU walkers (tool_addb == 0) ----------------------------------
 * Mirrors xevdm_deblock + deblock_tree's leaves (src_main/xevdm.c:1935-2103) for one tile / one slice: COD bits cleared,
 * every CU visited in decoding order by xevdm_deblock_cu_ver (pass 1) then xevdm_deblock_cu_hor (pass 2), CUs larger than
 * MAX_TR_SIZE visited as two halves. */
/* PPS tile grid for ref_deblock_frame / ref_alf_frame (same arguments as xb200_set_tiles / orc_set_tiles) */
static struct { int n_cols, n_rows, across; uint16_t col_bd[XB200_MAX_TILE_COLS + 1], row_bd[XB200_MAX_TILE_ROWS + 1]; } g_ref_tiles = {1, 1, 0, {0, 0xffff}, {0, 0xffff}};
void ref_set_tiles(int n_cols, const uint16_t *col_bd, int n_rows, const uint16_t *row_bd, int across)
{
    int i;
    g_ref_tiles.n_cols = n_cols; g_ref_tiles.n_rows = n_rows; g_ref_tiles.across = across != 0;
    for (i = 0; i <= n_cols; i++) g_ref_tiles.col_bd[i] = col_bd[i];
    for (i = 0; i <= n_rows; i++) g_ref_tiles.row_bd[i] = row_bd[i];
}
/* tile index of the CTU at (cx, cy) in CTU units, as set_tile_info numbers them (src_main/xevdm.c:2275-2327) */
static int ref_tile_of(int cx, int cy)
{
    int tc = 0, tr = 0;
    while (tc + 1 < g_ref_tiles.n_cols && cx >= g_ref_tiles.col_bd[tc + 1]) tc++;
    while (tr + 1 < g_ref_tiles.n_rows && cy >= g_ref_tiles.row_bd[tr + 1]) tr++;
    return tr * g_ref_tiles.n_cols + tc;
}

int ref_deblock_frame(const XB200_PARAMS *prm, ORC_PIC *pic, const XB200_CU *cus, int n_cu, const int *chroma_qp_tbl, int tool_addb,
                      const int *ref_id_l0, int n_l0, const int *ref_id_l1, int n_l1)
{
    /* get_bs compares XEVD_PIC pointers: pictures with the same id share one dummy XEVD_PIC */
    static XEVD_PIC dummy[64];
    static XEVD_SPS sps;
    XEVDM_CTX *m = (XEVDM_CTX *)calloc(1, sizeof(XEVDM_CTX));
    XEVD_CTX *ctx = &m->bctx;
    XEVD_PIC xp;
    const int f_scu = pic->w_scu * pic->h_scu;
    u32 *map_scu = (u32 *)malloc(sizeof(u32) * f_scu);
    TREE_CONS tc = { FALSE, TREE_LC, eAll };
    int pass, n, i;
    ensure_init();
    memset(&sps, 0, sizeof(sps));
    sps.bit_depth_luma_minus8 = prm->bit_depth_luma - 8;
    sps.bit_depth_chroma_minus8 = prm->bit_depth_chroma - 8;
    sps.chroma_format_idc = prm->chroma_format_idc;
    sps.tool_addb = tool_addb;
    ctx->sps = &sps;
    memcpy(map_scu, pic->map_scu, sizeof(u32) * f_scu);
    ctx->map_scu = map_scu;
    ctx->map_refi = (s8(*)[REFP_NUM])pic->map_refi;
    ctx->map_mv = (s16(*)[REFP_NUM][MV_D])pic->map_mv;
    /* xevdm_deblock's first loop (src_main/xevdm.c:2077-2090): map_mv replaces map_unrefined_mv wherever the DMVR flag is not set.  Done on
     * a copy: the caller's maps are inputs here */
    s16 (*umv)[REFP_NUM][MV_D] = (s16(*)[REFP_NUM][MV_D])malloc(sizeof(s16) * REFP_NUM * MV_D * f_scu);
    memcpy(umv, pic->map_unrefined_mv ? pic->map_unrefined_mv : pic->map_mv, sizeof(s16) * REFP_NUM * MV_D * f_scu);
    for (i = 0; i < f_scu; i++)
        if (!MCU_GET_DMVRF(pic->map_scu[i])) memcpy(umv[i], ctx->map_mv[i], sizeof(umv[i]));
    m->map_unrefined_mv = umv;
    ctx->w_scu = pic->w_scu; ctx->h_scu = pic->h_scu; ctx->w = pic->w_l; ctx->h = pic->h_l;
    ctx->log2_max_cuwh = prm->log2_ctu;
    ctx->map_tidx = (u8 *)calloc(f_scu, 1);
    for (i = 0; i < f_scu; i++)        /* what set_tile_info writes (src_main/xevdm.c:2298-2320) */
        ctx->map_tidx[i] = (u8)ref_tile_of(((i % pic->w_scu) << 2) >> prm->log2_ctu, ((i / pic->w_scu) << 2) >> prm->log2_ctu);
    ctx->map_cu_mode = (u32 *)calloc(f_scu, sizeof(u32));
    m->map_ats_inter = (u8 *)calloc(f_scu, 1);
    if (prm->tool_ats)          /* what xevdm_set_dec_info leaves in map_ats_inter (src_main/xevdm_util.c:4307-4311) */
        for (n = 0; n < n_cu; n++)
            if (cus[n].mode != XB200_MODE_INTRA && cus[n].mode != XB200_MODE_IBC && XB200_ATS_INTER_IDX(cus[n].ats)) {
                int j;
                for (j = 0; j < (1 << (cus[n].log2h - 2)); j++)
                    memset(m->map_ats_inter + ((cus[n].y >> 2) + j) * pic->w_scu + (cus[n].x >> 2),
                           get_ats_inter_info(XB200_ATS_INTER_IDX(cus[n].ats), XB200_ATS_INTER_POS(cus[n].ats)), 1 << (cus[n].log2w - 2));
            }
    ctx->fn_dbk = g_ctx->fn_dbk; ctx->fn_dbk_chroma = g_ctx->fn_dbk_chroma;
    for (i = 0; i < n_l0 && ref_id_l0; i++) ctx->refp[i][REFP_0].pic = &dummy[ref_id_l0[i] & 63];
    for (i = 0; i < n_l1 && ref_id_l1; i++) ctx->refp[i][REFP_1].pic = &dummy[ref_id_l1[i] & 63];
    wrap_pic(pic, &xp);
    xp.pic_qp_u_offset = prm->qp_u_offset; xp.pic_qp_v_offset = prm->qp_v_offset;
    xp.pic_deblock_alpha_offset = prm->deblock_alpha_offset; xp.pic_deblock_beta_offset = prm->deblock_beta_offset;
    /* chroma QP mapping: same set-up as sequence_init (src_main/xevdm.c:471-486) with the caller's table */
    xevd_set_chroma_qp_tbl_loc(prm->bit_depth_chroma);
    memcpy(xevd_qp_chroma_dynamic[0], chroma_qp_tbl, sizeof(int) * XEVD_MAX_QP_TABLE_SIZE);
    memcpy(xevd_qp_chroma_dynamic[1], chroma_qp_tbl + XEVD_MAX_QP_TABLE_SIZE, sizeof(int) * XEVD_MAX_QP_TABLE_SIZE);

    for (pass = 0; pass < 2; pass++) {
        for (i = 0; i < f_scu; i++) MCU_CLR_COD(map_scu[i]);
        for (n = 0; n < n_cu; n++) {
            const int x = cus[n].x, y = cus[n].y, w = 1 << cus[n].log2w, h = 1 << cus[n].log2h;
            /* deblock_tree visits the TREE_L leaves, then the parent block once more as TREE_C (src_main/xevdm.c:1991-1998) */
            const int do_l = (cus[n].flags & XB200_CUF_LUMA) != 0, do_c = (cus[n].flags & XB200_CUF_CHROMA) != 0;
            tc.tree_type = do_l && do_c ? TREE_LC : (do_l ? TREE_L : TREE_C);
            tc.mode_cons = do_l && do_c ? eAll : eOnlyIntra;
            if (pass == 0) {
                const int parts = w > MAX_TR_SIZE ? 2 : 1;
                for (i = 0; i < parts; i++)
                    xevdm_deblock_cu_ver(ctx, &xp, x + i * MAX_TR_SIZE, y, w / parts, h, ctx->map_scu, ctx->map_refi, m->map_unrefined_mv, ctx->w_scu,
                                         ctx->log2_max_cuwh, ctx->map_cu_mode, ctx->refp, 0, tc, ctx->map_tidx, g_ref_tiles.across, tool_addb, m->map_ats_inter,
                                         prm->bit_depth_luma, prm->bit_depth_chroma, prm->chroma_format_idc);
            } else {
                const int parts = h > MAX_TR_SIZE ? 2 : 1;
                for (i = 0; i < parts; i++)
                    xevdm_deblock_cu_hor(ctx, &xp, x, y + i * MAX_TR_SIZE, w, h / parts, ctx->map_scu, ctx->map_refi, m->map_unrefined_mv, ctx->w_scu,
                                         ctx->log2_max_cuwh, ctx->refp, 0, tc, ctx->map_tidx, g_ref_tiles.across, tool_addb, m->map_ats_inter,
                                         prm->bit_depth_luma, prm->bit_depth_chroma, prm->chroma_format_idc);
            }
        }
    }
    free(ctx->map_tidx); free(ctx->map_cu_mode); free(m->map_ats_inter); free(map_scu); free(umv); free(m);
    return XB200_OK;
}

/* ---- DRA on pull: the reference's plane functions (src_main/xevdm_dra.c:272-354) in the order xevd_apply_filter calls them --------- */
#include "xevdm_dra.h"
void ref_dra_apply(ORC_PIC *pic, const XB200_DRA *d)
{
    static DRA_CONTROL dc;
    XEVD_IMGB im;
    memset(&im, 0, sizeof(im));
    memset(&dc, 0, sizeof(dc));
    memcpy(dc.luma_inv_scale_lut, d->luma_inv_scale_lut, sizeof(dc.luma_inv_scale_lut));
    memcpy(dc.int_chroma_inv_scale_lut, d->chroma_inv_scale_lut, sizeof(dc.int_chroma_inv_scale_lut));
    im.np = 3;
    im.a[0] = pic->y; im.a[1] = pic->u; im.a[2] = pic->v;
    im.w[0] = pic->w_l; im.h[0] = pic->h_l; im.w[1] = im.w[2] = pic->w_c; im.h[1] = im.h[2] = pic->h_c;
    im.s[0] = pic->s_l * 2; im.s[1] = im.s[2] = pic->s_c * 2;
    xevd_apply_dra_chroma_plane(&im, &im, &dc, 1, TRUE);
    xevd_apply_dra_chroma_plane(&im, &im, &dc, 2, TRUE);
    xevd_apply_dra_luma_plane(&im, &im, &dc, 0, TRUE);
}

/* ---- adaptive loop filter: the reference's own per-tile driver alf_process_tile (src_main/xevdm_alf.c:901) ----------------
 * with the final coefficients installed directly (alf->coef_final / chroma_coef), one tile covering the picture. */
int alf_process_tile(void *arg);
typedef struct { ADAPTIVE_LOOP_FILTER *alf; CODING_STRUCTURE *cs; ALF_SLICE_PARAM *alf_slice_param; int tile_idx; int tsk_num; } REF_ALF_TMP;

int ref_alf_frame(const XB200_PARAMS *prm, ORC_PIC *pic, const XB200_ALF *ap, const uint8_t *ctb_flag_luma)
{
    static XEVD_SPS sps;
    XEVDM_CTX *m = (XEVDM_CTX *)calloc(1, sizeof(XEVDM_CTX));
    XEVD_CTX *ctx = &m->bctx;
    XEVD_PIC xp;
    CODING_STRUCTURE cs;
    ALF_SLICE_PARAM *sp = (ALF_SLICE_PARAM *)calloc(1, sizeof(ALF_SLICE_PARAM));
    ADAPTIVE_LOOP_FILTER *alf;
    REF_ALF_TMP tmp;
    const int ctu = 1 << prm->log2_ctu;
    int i, c;
    ensure_init();
    if (!ap->enable[0] && !ap->enable[1] && !ap->enable[2]) { free(m); free(sp); return XB200_OK; }
    memset(&sps, 0, sizeof(sps));
    sps.chroma_format_idc = 1;
    sps.pic_width_in_luma_samples = pic->w_l; sps.pic_height_in_luma_samples = pic->h_l;
    ctx->sps = &sps;
    ctx->w = pic->w_l; ctx->h = pic->h_l; ctx->w_scu = pic->w_scu; ctx->h_scu = pic->h_scu;
    ctx->log2_max_cuwh = prm->log2_ctu; ctx->max_cuwh = ctu;
    ctx->w_lcu = (pic->w_l + ctu - 1) / ctu; ctx->h_lcu = (pic->h_l + ctu - 1) / ctu; ctx->f_lcu = ctx->w_lcu * ctx->h_lcu;
    {   /* the tile array as set_tile_info builds it (src_main/xevdm.c:2275-2296); a grid that was never set is one tile */
        const int nc = g_ref_tiles.n_cols, nr = g_ref_tiles.n_rows;
        int tx, ty;
        ctx->w_tile = nc; ctx->h_tile = nr;
        ctx->tile = (XEVD_TILE *)calloc((size_t)nc * nr, sizeof(XEVD_TILE));
        for (ty = 0; ty < nr; ty++)
            for (tx = 0; tx < nc; tx++) {
                XEVD_TILE *t = &ctx->tile[ty * nc + tx];
                const int c0 = g_ref_tiles.col_bd[tx], c1 = tx + 1 == nc ? ctx->w_lcu : g_ref_tiles.col_bd[tx + 1];
                const int r0 = g_ref_tiles.row_bd[ty], r1 = ty + 1 == nr ? ctx->h_lcu : g_ref_tiles.row_bd[ty + 1];
                t->ctba_rs_first = r0 * ctx->w_lcu + c0; t->w_ctb = (u16)(c1 - c0); t->h_ctb = (u16)(r1 - r0); t->f_ctb = t->w_ctb * t->h_ctb;
            }
        ctx->pps.num_tile_columns_minus1 = nc - 1; ctx->pps.num_tile_rows_minus1 = nr - 1;
        ctx->pps.loop_filter_across_tiles_enabled_flag = g_ref_tiles.across;
    }
    wrap_pic(pic, &xp);
    cs.ctx = ctx; cs.pic = &xp;
    alf = new_alf(prm->bit_depth_luma);
    xevd_alf_create(alf, pic->w_l, pic->h_l, ctu, ctu, 5, 1, prm->bit_depth_luma);
    for (c = 0; c < 3; c++) sp->enable_flag[c] = ap->enable[c];
    memcpy(sp->chroma_coef, ap->coef_chroma, sizeof(short) * 7);
    memcpy(alf->coef_final, ap->coef_luma, sizeof(short) * 25 * 13);
    sp->alf_ctb_flag = (u8 *)malloc(3 * ctx->f_lcu);
    for (i = 0; i < ctx->f_lcu; i++) {
        sp->alf_ctb_flag[i] = (u8)((ap->enable[0] && (!ctb_flag_luma || ctb_flag_luma[i])) ? 1 : 0);
        sp->alf_ctb_flag[ctx->f_lcu + i] = ap->enable[1];
        sp->alf_ctb_flag[2 * ctx->f_lcu + i] = ap->enable[2];
    }
    for (c = 0; c < 3; c++) alf->ctu_enable_flag[c] = sp->alf_ctb_flag + ctx->f_lcu * c;
    for (i = 0; i < ctx->w_tile * ctx->h_tile; i++) {       /* alf_process (:1203-1247) runs the tiles one after the other */
        tmp.alf = alf; tmp.cs = &cs; tmp.alf_slice_param = sp; tmp.tile_idx = i; tmp.tsk_num = 0;
        alf_process_tile(&tmp);
    }
    xevd_alf_destroy(alf); delete_alf(alf);
    free(sp->alf_ctb_flag); free(sp); free(ctx->tile); free(m);
    return XB200_OK;
}
