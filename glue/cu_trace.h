/* cu_trace.h -- debugging aid shared by the drop-in library and by libxevd_reftrace.so (the unmodified reference with one logging
 * wrapper): one text line per coding unit with the state xevd_recon_unit has reached when it calls xevdm_recon_yuv, written to the
 * file named by XEVD_CU_TRACE.  tools/cu_trace_diff.py compares the two logs: the first differing line is the first CU whose motion
 * derivation / syntax state differs between the host halves of the two decoders. */
#ifndef XB200_CU_TRACE_H
#define XB200_CU_TRACE_H
#include <stdio.h>
#include <stdlib.h>
#include "xevdm_def.h"

static FILE *cu_trace_fp(void)
{
    static FILE *fp; static int tried;
    if (!tried) { const char *p = getenv("XEVD_CU_TRACE"); tried = 1; if (p) fp = fopen(p, "w"); }
    return fp;
}

static void cu_trace(XEVD_CTX *ctx, XEVD_CORE *core, int x, int y, int cuw, int cuh, int tree_type, int ats_inter_info)
{
    FILE *fp = cu_trace_fp();
    if (!fp) return;
    XEVDM_CORE *m = (XEVDM_CORE *)core;
    const u32 scu = ctx->map_scu[core->scup];
    fprintf(fp, "poc %d cu %d %d %dx%d tree %d mode %d aff %d refi %d %d ipm %d %d qp %d %d %d cbf %d %d %d ats %d %d %d avail_lr %d",
            ctx->poc.poc_val, x, y, cuw, cuh, tree_type, core->pred_mode, core->pred_mode == MODE_INTRA || core->pred_mode == MODE_IBC ? 0 : m->affine_flag,
            core->refi[0], core->refi[1], core->ipm[0], core->ipm[1], core->qp_y, core->qp_u, core->qp_v, core->is_coef[0], core->is_coef[1], core->is_coef[2],
            m->ats_intra_cu, (m->ats_intra_mode_h << 1) | m->ats_intra_mode_v, ats_inter_info, core->avail_lr);
    if (core->pred_mode != MODE_INTRA) {
        if (core->pred_mode != MODE_IBC && m->affine_flag) {
            fprintf(fp, " cp");
            for (int l = 0; l < 2; l++) for (int v = 0; v < 3; v++) fprintf(fp, " %d,%d", m->affine_mv[l][v][0], m->affine_mv[l][v][1]);
        }
        /* what xevdm_set_dec_info left in the maps for later CUs' candidates (unrefined vectors, refi) */
        XEVDM_CTX *mctx = (XEVDM_CTX *)ctx;
        fprintf(fp, " umv %d,%d %d,%d maprefi %d %d scu %08x", mctx->map_unrefined_mv[core->scup][0][0], mctx->map_unrefined_mv[core->scup][0][1],
                mctx->map_unrefined_mv[core->scup][1][0], mctx->map_unrefined_mv[core->scup][1][1], ctx->map_refi[core->scup][0], ctx->map_refi[core->scup][1],
                scu & ~((1u << 25) | (1u << 31)));
    }
    fprintf(fp, "\n");
    fflush(fp);
}
#endif
