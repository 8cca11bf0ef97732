/* ref_trace.c -- libxevd_reftrace.so: the unmodified reference decoder with ONE wrapper, around xevdm_recon_yuv, that logs the coding
 * unit (glue/cu_trace.h) and then calls the real function.  Debugging aid for the drop-in library (test infrastructure). */
#include "cu_trace.h"
void xevdm_recon_yuv(int x, int y, int cuw, int cuh, s16 coef[N_C][MAX_CU_DIM], pel pred[N_C][MAX_CU_DIM], int nnz[N_C], XEVD_PIC *pic,
                     u8 ats_inter_info, TREE_CONS tree_cons, int bit_depth, int chroma_format_idc);
void trace_recon_yuv(int x, int y, int cuw, int cuh, s16 coef[N_C][MAX_CU_DIM], pel pred[N_C][MAX_CU_DIM], int nnz[N_C], XEVD_PIC *pic,
                     u8 ats_inter_info, TREE_CONS tree_cons, int bit_depth, int chroma_format_idc)
{
    XEVD_CORE *core = (XEVD_CORE *)((char *)coef - offsetof(XEVD_CORE, coef));
    cu_trace(core->ctx, core, x, y, cuw, cuh, tree_cons.tree_type, ats_inter_info);
    xevdm_recon_yuv(x, y, cuw, cuh, coef, pred, nnz, pic, ats_inter_info, tree_cons, bit_depth, chroma_format_idc);
}
