/*
 * xevd_b200_glue.c -- the reference decoder with its reconstruction half on the GPU: libxevd_gpu.so
 *
 * What this file is: the reference-side binding of include/xevd_b200.h.  It is compiled by gcc against the reference's
 * own headers and linked with the reference's own objects (glue/Makefile), and the resulting library exports exactly the
 * six entry points of inc/xevd.h:369-374 (xevd_create / xevd_delete / xevd_decode / xevd_pull / xevd_config / xevd_info),
 * so an application that links libxevd (app/xevd_app.c, FFmpeg's libxevd wrapper) links this instead and changes nothing.
 *
 * How the seam is cut (nothing of the reference is copied or edited; glue/Makefile renames symbols at compile time):
 *
 *   - src_main/xevdm.c is compiled with its public entry points renamed to xevdref_* (-D on the command line), and with every
 *     PIXEL function xevd_recon_unit calls (src_main/xevdm.c:1230-1405) renamed to a glue_* hook defined here:
 *       xevdm_sub_block_itdq, xevdm_mc, xevdm_affine_mc, xevdm_IBC_mc, xevdm_get_nbr, xevd_get_nbr_b, xevdm_ipred(_uv),
 *       xevd_ipred(_uv)_b, xevdm_recon_yuv, xevdm_htdf.
 *     Everything else of the per-CU path runs as the reference wrote it: the split-tree walk xevd_recon_tree (:1854, SUCO order,
 *     local dual tree), cu_init (:1022), coef_rect_to_series (:1185), the motion derivation (xevd_get_skip_motion /
 *     xevd_get_inter_motion / xevd_get_direct_motion / xevdm_get_mmvd_motion / xevd_get_affine_motion, :800-1020), HMVP,
 *     xevdm_set_dec_info (src_main/xevdm_util.c:4205) on the HOST maps that later CUs' derivation reads.
 *     The hooks do no arithmetic on samples: they record what the reference was about to compute as one XB200_CU work item
 *     (+ coefficient blocks + the neighbour-availability masks xevdm_get_nbr derives from the COD bits).
 *   - ctx->fn_dec_slice (xevdm_dec_slice, :2608) is wrapped: entropy decode + the hooked walk fill the work lists of the slice,
 *     then ONE xb200_recon_frame call reconstructs it on the device.
 *   - ctx->fn_deblock (xevdm_deblock, :2048; called per tile and per pass, :3152-3202) -> one xb200_deblock per picture.
 *   - ALF: src_main/xevdm_alf.c keeps alf_process (:1167, APS -> alf_recon_coef -> coef_final on the host); its per-tile worker
 *     alf_process_tile (:901) is the hook -> xb200_alf.
 *   - ctx->fn_picbuf_expand (xevd_picbuf_expand) -> xb200_pad, then an asynchronous copy of the picture into the host XEVD_IMGB
 *     the application will receive from xevd_pull (and, with tool_dmvr, of the refined map_mv the NEXT pictures' temporal
 *     candidates read).
 *   - PICBUF_ALLOCATOR (ctx->pa.fn_alloc = xevdm_picbuf_alloc, :457): every XEVD_PIC gets a device twin (xb200_pic).
 *
 * There is no CPU fallback: xevd_create fails when no CUDA device is usable.
 * Limit (refused loudly with XEVD_ERR_UNSUPPORTED_COLORSPACE): chroma formats other than 4:2:0.
 * Pictures of several tiles: the chunks of a slice are re-sorted into raster order and the tile grid goes to the loop filters
 * (xb200_set_tiles).
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "xevdm_def.h"
#include "xevdm_alf.h"
#include "xevd_b200.h"
#include "cu_trace.h"

/* the reference's entry points, as src_main/xevdm.c defines them under the names glue/Makefile gives them */
XEVD xevdref_create(XEVD_CDSC *cdsc, int *err);
void xevdref_delete(XEVD id);
int  xevdref_decode(XEVD id, XEVD_BITB *bitb, XEVD_STAT *stat);
int  xevdref_pull(XEVD id, XEVD_IMGB **imgb);
int  xevdref_config(XEVD id, int cfg, void *buf, int *size);

/* reference functions the hooks forward to (unrenamed in their own translation units) */
XEVD_PIC *xevdm_picbuf_alloc(PICBUF_ALLOCATOR *pa, int *ret, int bitdepth);
void      xevdm_picbuf_free(PICBUF_ALLOCATOR *pa, XEVD_PIC *pic);
XEVD_IMGB *xevd_imgb_generate(int w, int h, int padl, int padc, int idc, int bit_depth);
void      xevdm_get_tu_size(u8 ats_inter_info, int log2_cuw, int log2_cuh, int *log2_tuw, int *log2_tuh);

#define GLUE_MAX_PICS 64

typedef struct GLUE_PIC {
    XEVD_PIC  *host;
    xb200_pic *dev;
    void      *registered[3];       /* page-locked plane buffers of the imgb (NULL where registration failed: that copy falls back to staged DMA) */
} GLUE_PIC;

typedef struct GLUE_CHUNK {         /* the CUs of one CTU, in the order the walk reached it */
    int ctu, cu0, cu1;
    size_t coef0, coef1;
} GLUE_CHUNK;

typedef struct GLUE {
    XEVD_CTX  *ctx;
    xb200_ctx *dev;
    int      (*ref_dec_slice)(XEVD_CTX *ctx, XEVD_CORE *core);
    /* work lists of the slice being decoded */
    XB200_CU     *cus;   int n_cu, cap_cu;
    XB200_CU_EXT *ext;   int n_ext, cap_ext;
    int16_t      *coef;  size_t n_coef, cap_coef;
    GLUE_CHUNK   *chunk; int n_chunk, cap_chunk;
    uint32_t     *ctu_first; int cap_ctu;
    XB200_CU     *cus2;  int16_t *coef2; int cap_cu2; size_t cap_coef2;     /* raster reorder (several tiles) */
    /* what the hooks saw of the CU that is being reconstructed, until xevdm_recon_yuv emits it */
    struct {
        int have_nbr; uint64_t up, left, right; int ul;
        int have_aff; s16 mv_unref[REFP_NUM][MV_D];
        int dmvr;
    } pend;
    int last_cu;                    /* index of the CU emitted last (xevdm_htdf, called right after, adds its avail_cu) */
    GLUE_PIC pics[GLUE_MAX_PICS];
    int dbk_done, alf_done;         /* per picture: the per-tile / per-pass callbacks launch once */
    int maps_pending;
    int err;                        /* first device error of the current access unit (returned by xevd_decode) */
    XB200_PARAMS prm;
    xb200_pic *l0[XEVD_MAX_NUM_REF_PICS], *l1[XEVD_MAX_NUM_REF_PICS];
    int n0, n1;
    long long n_pictures, n_cus_total;
    double t_host, t_recon, t_dbk, t_alf, t_out, t_wait, t_alloc;       /* XEVD_B200_STATS: seconds spent per stage (host side of each call) */
} GLUE;

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static GLUE *g_inst[8];

static GLUE *glue_of(const XEVD_CTX *ctx)
{
    for (int i = 0; i < 8; i++) if (g_inst[i] && g_inst[i]->ctx == ctx) return g_inst[i];
    return NULL;
}
static GLUE_PIC *pic_of(GLUE *g, const XEVD_PIC *p)
{
    for (int i = 0; i < GLUE_MAX_PICS; i++) if (g->pics[i].host == p) return &g->pics[i];
    return NULL;
}
static void dev_fail(GLUE *g, int code, const char *what)
{
    if (!g->err) {
        g->err = code;
        fprintf(stderr, "[xevd-b200] %s failed: %d (%s)\n", what, code, xb200_last_error(g->dev));
    }
}
#define CORE_FROM(member, ptr) ((XEVD_CORE *)((char *)(ptr) - offsetof(XEVD_CORE, member)))

static int grow(void **p, int *cap, int need, size_t elem)
{
    if (need <= *cap) return 0;
    int n = *cap ? *cap : 1024;
    while (n < need) n *= 2;
    void *q = realloc(*p, (size_t)n * elem);
    if (!q) return -1;
    *p = q; *cap = n;
    return 0;
}

/* debugging aid (XEVD_B200_DUMP=<directory>): the device picture after a stage, compact planes, for tools/gpu_replay.py */
static int g_dump_serial;
static void dump_stage(GLUE *g, GLUE_PIC *p, const char *stage)
{
    const char *dir = getenv("XEVD_B200_DUMP");
    if (!dir || !p || getenv("XEVD_B200_DUMP_NOPICS")) return;          /* work lists only: pictures of large streams are big */
    const int w = g->ctx->w, h = g->ctx->h;
    int16_t *buf = (int16_t *)malloc((size_t)w * h * 3 / 2 * sizeof(int16_t) + 64);
    if (!buf) return;
    int16_t *y = buf, *u = buf + (size_t)w * h, *v = u + (size_t)(w / 2) * (h / 2);
    if (xb200_pic_download(g->dev, p->dev, y, w, u, w / 2, v, w / 2) >= 0 && xb200_sync(g->dev) >= 0) {
        char fn[512];
        snprintf(fn, sizeof(fn), "%s/pic_%04d_%s.bin", dir, g_dump_serial - 1, stage);
        FILE *fp = fopen(fn, "wb");
        if (fp) { int32_t hdr[2] = { w, h }; fwrite(hdr, sizeof(hdr), 1, fp); fwrite(buf, 2, (size_t)w * h * 3 / 2, fp); fclose(fp); }
    }
    free(buf);
    if (!strcmp(stage, "recon")) {          /* what deblocking is about to read: per-SCU maps, edge map, parameters */
        const size_t n = (size_t)g->ctx->f_scu;
        uint8_t *m = (uint8_t *)calloc(n, 8 + 2 + 4 + 8 + 1);
        if (!m) return;
        int16_t *mv = (int16_t *)m; int8_t *refi = (int8_t *)(m + n * 8); uint32_t *scu = (uint32_t *)(m + n * 10);
        int16_t *umv = (int16_t *)(m + n * 14); uint8_t *edge = m + n * 22;
        xb200_pic_download_maps(g->dev, p->dev, mv, refi, scu);
        xb200_pic_download_unrefined_mv(g->dev, p->dev, umv);
        xb200_pic_download_edge_map(g->dev, p->dev, edge);
        xb200_sync(g->dev);
        char fn[512];
        snprintf(fn, sizeof(fn), "%s/pic_%04d_maps.bin", dir, g_dump_serial - 1);
        FILE *fp = fopen(fn, "wb");
        if (fp) { int32_t hdr[2] = { (int32_t)n, 0 }; fwrite(hdr, sizeof(hdr), 1, fp); fwrite(m, 23, n, fp); fclose(fp); }
        free(m);
    }
}

/* ---- hooks: the pixel functions of xevd_recon_unit --------------------------------------------------------------------- */

/* xevdm_sub_block_itdq (src_main/xevdm_itdq.c:790): dequantisation + inverse transform happen on the device; the
 * coefficients stay untouched in core->coef until the CU is emitted */
void glue_sub_block_itdq(XEVD_CTX *ctx, s16 coef[N_C][MAX_CU_DIM], int log2_cuw, int log2_cuh, u8 qp_y, u8 qp_u, u8 qp_v, int flag[N_C],
                         int nnz_sub[N_C][MAX_SUB_TB_NUM], int iqt_flag, u8 ats_intra_cu, u8 ats_mode, u8 ats_inter_info, int bit_depth,
                         int chroma_format_idc)
{
    (void)ctx; (void)coef; (void)log2_cuw; (void)log2_cuh; (void)qp_y; (void)qp_u; (void)qp_v; (void)flag; (void)nnz_sub; (void)iqt_flag;
    (void)ats_intra_cu; (void)ats_mode; (void)ats_inter_info; (void)bit_depth; (void)chroma_format_idc;
}

/* xevdm_mc (src_main/xevdm_mc.c:1860).  The DMVR decision (POC distances, block size, identical motion) is taken on the device;
 * here only the enable the reference computed (xevdm.c:1273-1288) is kept.  *cu_dmvr_flag stays 0, so xevdm_set_dec_info publishes
 * the unrefined vectors to the host maps - which is what every same-picture reader wants (spatial candidates of DMVR neighbours read
 * map_unrefined_mv, xevdm_util.c:791-1215); the refined map comes back from the device once per picture. */
void glue_mc(int x, int y, int pic_w, int pic_h, int w, int h, s8 refi[REFP_NUM], s16 (*mv)[MV_D], XEVD_REFP (*refp)[REFP_NUM],
             pel pred[REFP_NUM][N_C][MAX_CU_DIM], int poc_c, pel *dmvr_current_template, void *dmvr_ref_pred_interpolated,
             void *dmvr_half_pred_interpolated, BOOL apply_DMVR, void *dmvr_padding_buf, u8 *cu_dmvr_flag, void *dmvr_mv,
             int sps_admvp_flag, int bit_depth_luma, int bit_depth_chroma, int chroma_format_idc)
{
    XEVD_CORE *core = CORE_FROM(pred, pred);
    GLUE *g = glue_of(core->ctx);
    (void)x; (void)y; (void)pic_w; (void)pic_h; (void)w; (void)h; (void)refi; (void)mv; (void)refp; (void)poc_c; (void)dmvr_current_template;
    (void)dmvr_ref_pred_interpolated; (void)dmvr_half_pred_interpolated; (void)dmvr_padding_buf; (void)cu_dmvr_flag; (void)dmvr_mv;
    (void)sps_admvp_flag; (void)bit_depth_luma; (void)bit_depth_chroma; (void)chroma_format_idc;
    if (g) g->pend.dmvr = apply_DMVR ? 1 : 0;
}

/* xevdm_affine_mc (src_main/xevdm_mc.c:2606): the control-point vectors are read from mcore->affine_mv when the CU is emitted;
 * core->mv is saved now because xevdm_set_dec_info overwrites it with the first sub-block's vector (xevdm_util.c:4383) */
void glue_affine_mc(int x, int y, int pic_w, int pic_h, int w, int h, s8 refi[REFP_NUM], s16 mv[REFP_NUM][VER_NUM][MV_D],
                    XEVD_REFP (*refp)[REFP_NUM], pel pred[2][N_C][MAX_CU_DIM], int vertex_num, pel *tmp_buffer, int bit_depth_luma,
                    int bit_depth_chroma, int chroma_format_idc)
{
    XEVD_CORE *core = CORE_FROM(pred, pred);
    GLUE *g = glue_of(core->ctx);
    (void)x; (void)y; (void)pic_w; (void)pic_h; (void)w; (void)h; (void)refi; (void)mv; (void)refp; (void)vertex_num; (void)tmp_buffer;
    (void)bit_depth_luma; (void)bit_depth_chroma; (void)chroma_format_idc;
    if (g) { g->pend.have_aff = 1; memcpy(g->pend.mv_unref, core->mv, sizeof(g->pend.mv_unref)); }
}

/* xevdm_IBC_mc (src_main/xevdm_mc.c:2040): a copy inside the current picture -> wavefront kernel */
void glue_IBC_mc(int x, int y, int log2_cuw, int log2_cuh, s16 mv[MV_D], XEVD_PIC *ref_pic, pel (*pred)[MAX_CU_DIM], TREE_CONS tree_cons,
                 int chroma_format_idc)
{
    (void)x; (void)y; (void)log2_cuw; (void)log2_cuh; (void)mv; (void)ref_pic; (void)pred; (void)tree_cons; (void)chroma_format_idc;
}

/* Neighbour availability, one bit per SCU, exactly the tests of xevdm_get_nbr (src_main/xevdm_ipred.c:54-145) / xevd_get_nbr_b
 * (src_base/xevd_ipred.c:49-91) on the host map_scu / map_tidx at this point of the decoding order.  The first call of a CU
 * (luma, or Cb for a chroma-only CU) fixes the masks; the chroma calls see the same SCUs. */
static void record_nbr(int eipd, int x, int y, int cuw, int cuh, u16 avail_cu, pel nb[N_C][N_REF][MAX_CU_SIZE * 3], int scup, u32 *map_scu,
                       int w_scu, int h_scu, int ch_type, int cip, u8 *map_tidx)
{
    XEVD_CORE *core = CORE_FROM(nb, nb);
    GLUE *g = glue_of(core->ctx);
    if (!g || g->pend.have_nbr) return;
    const int sh = ch_type == Y_C ? 0 : 1;                    /* 4:2:0 only (checked at sequence level) */
    const int scuw = (cuw << sh) >> MIN_CU_LOG2, scuh = (cuh << sh) >> MIN_CU_LOG2;
    const int x_scu = PEL2SCU(x << sh), y_scu = PEL2SCU(y << sh);
    uint64_t up = 0, left = 0, right = 0;
    int ul;
#define NBR_OK(p) (MCU_GET_COD(map_scu[p]) && (!cip || MCU_GET_IF(map_scu[p])) && map_tidx[scup] == map_tidx[p])
    for (int i = 0; i < scuw + scuh; i++) {
        if (y_scu > 0 && x_scu + i < w_scu && NBR_OK(scup - w_scu + i)) up |= 1ull << i;
        if (x_scu > 0 && y_scu + i < h_scu && NBR_OK(scup - 1 + i * w_scu)) left |= 1ull << i;
        if (eipd && x_scu + scuw < w_scu && y_scu + i < h_scu && NBR_OK(scup + scuw + i * w_scu)) right |= 1ull << i;
    }
    if (eipd)       /* the loop over the units left of the corner decides what up[-1] ends up as (xevdm_ipred.c:85-104) */
        ul = x_scu > 0 && scup > 0 && y_scu > 0 && NBR_OK(scup - w_scu - 1);
    else            /* xevd_ipred.c:49-57 */
        ul = IS_AVAIL(avail_cu, AVAIL_UP_LE) && (!cip || MCU_GET_IF(map_scu[scup - w_scu - 1])) && map_tidx[scup] == map_tidx[scup - w_scu - 1];
#undef NBR_OK
    g->pend.have_nbr = 1; g->pend.up = up; g->pend.left = left; g->pend.right = right; g->pend.ul = ul;
}
void glue_get_nbr(int x, int y, int cuw, int cuh, pel *src, int s_src, u16 avail_cu, pel nb[N_C][N_REF][MAX_CU_SIZE * 3], int scup, u32 *map_scu,
                  int w_scu, int h_scu, int ch_type, int constrained_intra_pred, u8 *map_tidx, int bit_depth, int chroma_format_idc)
{
    (void)src; (void)s_src; (void)bit_depth; (void)chroma_format_idc;
    record_nbr(1, x, y, cuw, cuh, avail_cu, nb, scup, map_scu, w_scu, h_scu, ch_type, constrained_intra_pred, map_tidx);
}
void glue_get_nbr_b(int x, int y, int cuw, int cuh, pel *src, int s_src, u16 avail_cu, pel nb[N_C][N_REF][MAX_CU_SIZE * 3], int scup, u32 *map_scu,
                    int w_scu, int h_scu, int ch_type, int constrained_intra_pred, u8 *map_tidx, int bit_depth, int chroma_format_idc)
{
    (void)src; (void)s_src; (void)bit_depth; (void)chroma_format_idc;
    record_nbr(0, x, y, cuw, cuh, avail_cu, nb, scup, map_scu, w_scu, h_scu, ch_type, constrained_intra_pred, map_tidx);
}
/* xevdm_ipred / xevdm_ipred_uv / xevd_ipred_b / xevd_ipred_uv_b: intra prediction runs in the wavefront kernel */
void glue_ipred(pel *src_le, pel *src_up, pel *src_ri, u16 avail_lr, pel *dst, int ipm, int w, int h, int bit_depth)
{ (void)src_le; (void)src_up; (void)src_ri; (void)avail_lr; (void)dst; (void)ipm; (void)w; (void)h; (void)bit_depth; }
void glue_ipred_uv(pel *src_le, pel *src_up, pel *src_ri, u16 avail_lr, pel *dst, int ipm_c, int ipm, int w, int h, int bit_depth)
{ (void)src_le; (void)src_up; (void)src_ri; (void)avail_lr; (void)dst; (void)ipm_c; (void)ipm; (void)w; (void)h; (void)bit_depth; }
void glue_ipred_b(pel *src_le, pel *src_up, pel *src_ri, u16 avail_lr, pel *dst, int ipm, int w, int h)
{ (void)src_le; (void)src_up; (void)src_ri; (void)avail_lr; (void)dst; (void)ipm; (void)w; (void)h; }
void glue_ipred_uv_b(pel *src_le, pel *src_up, pel *src_ri, u16 avail_lr, pel *dst, int ipm_c, int ipm, int w, int h)
{ (void)src_le; (void)src_up; (void)src_ri; (void)avail_lr; (void)dst; (void)ipm_c; (void)ipm; (void)w; (void)h; }

/* xevdm_recon_yuv (src_main/xevdm_recon.c:128): the last pixel call of every CU - the work item is complete here */
void glue_recon_yuv(int x, int y, int cuw, int cuh, s16 coef[N_C][MAX_CU_DIM], pel pred[N_C][MAX_CU_DIM], int nnz[N_C], XEVD_PIC *pic,
                    u8 ats_inter_info, TREE_CONS tree_cons, int bit_depth, int chroma_format_idc)
{
    XEVD_CORE *core = CORE_FROM(coef, coef);
    XEVDM_CORE *mcore = (XEVDM_CORE *)core;
    XEVD_CTX *ctx = core->ctx;
    GLUE *g = glue_of(ctx);
    (void)pred; (void)nnz; (void)pic; (void)bit_depth; (void)chroma_format_idc;
    if (!g || g->err) return;
    cu_trace(ctx, core, x, y, cuw, cuh, tree_cons.tree_type, ats_inter_info);
    const int log2_ctu = ctx->log2_max_cuwh;
    const int ctu = (y >> log2_ctu) * ctx->w_lcu + (x >> log2_ctu);
    if (grow((void **)&g->cus, &g->cap_cu, g->n_cu + 1, sizeof(XB200_CU)) || grow((void **)&g->ext, &g->cap_ext, g->n_ext + 2, sizeof(XB200_CU_EXT))) {
        g->err = XEVD_ERR_OUT_OF_MEMORY; return;
    }
    if (g->n_ext == 0) memset(&g->ext[g->n_ext++], 0, sizeof(XB200_CU_EXT));        /* record 0 is never referenced */
    if (g->n_chunk == 0 || g->chunk[g->n_chunk - 1].ctu != ctu) {
        if (grow((void **)&g->chunk, &g->cap_chunk, g->n_chunk + 1, sizeof(GLUE_CHUNK))) { g->err = XEVD_ERR_OUT_OF_MEMORY; return; }
        GLUE_CHUNK *k = &g->chunk[g->n_chunk++];
        k->ctu = ctu; k->cu0 = k->cu1 = g->n_cu; k->coef0 = k->coef1 = g->n_coef;
    }
    const int do_l = tree_cons.tree_type != TREE_C, do_c = tree_cons.tree_type != TREE_L;
    const int log2w = XEVD_CONV_LOG2(cuw), log2h = XEVD_CONV_LOG2(cuh);
    XB200_CU *c = &g->cus[g->n_cu];
    memset(c, 0, sizeof(*c));
    c->x = (uint16_t)x; c->y = (uint16_t)y; c->log2w = (uint8_t)log2w; c->log2h = (uint8_t)log2h;
    c->flags = (uint8_t)((do_l ? XB200_CUF_LUMA : 0) | (do_c ? XB200_CUF_CHROMA : 0));
    c->qp_y = core->qp_y; c->qp_u = core->qp_u; c->qp_v = core->qp_v;
    /* the QP xevdm_set_dec_info stores in map_scu (xevdm_util.c:4303-4311): what deblocking reads */
    c->qp_map = (uint8_t)(ctx->pps.cu_qp_delta_enabled_flag ? core->qp : ctx->tile[core->tile_num].qp);
    c->avail = (uint8_t)(core->avail_lr & 3);
    const int ibc = core->pred_mode == MODE_IBC, intra = core->pred_mode == MODE_INTRA;
    if (intra) {
        c->mode = XB200_MODE_INTRA;
        c->refi[0] = (int8_t)core->ipm[0]; c->refi[1] = (int8_t)core->ipm[1];
        if (mcore->ats_intra_cu && ctx->sps->tool_ats && do_l) {
            c->flags |= XB200_CUF_ATS_INTRA;
            c->ats = (uint8_t)(((mcore->ats_intra_mode_h & 1) << 1) | (mcore->ats_intra_mode_v & 1));
        }
        XB200_CU_EXT *e = &g->ext[g->n_ext];
        memset(e, 0, sizeof(*e));
        e->u.intra.up = g->pend.up; e->u.intra.left = g->pend.left; e->u.intra.right = g->pend.right;
        c->avail |= (uint8_t)(g->pend.ul ? 4 : 0);
        const uint32_t ei = (uint32_t)g->n_ext++;
        memcpy(c->mv[1], &ei, 4);
    } else if (ibc) {
        c->mode = XB200_MODE_IBC;
        c->refi[0] = c->refi[1] = -1;
        c->mv[0][0] = core->mv[0][MV_X]; c->mv[0][1] = core->mv[0][MV_Y];
    } else {
        c->refi[0] = core->refi[REFP_0]; c->refi[1] = core->refi[REFP_1];
        if (core->pred_mode == MODE_SKIP) c->flags |= XB200_CUF_SKIP;
        if (mcore->affine_flag) {
            c->mode = XB200_MODE_AFFINE;
            if (mcore->affine_flag == 2) c->flags |= XB200_CUF_AFF6;
            XB200_CU_EXT *e = &g->ext[g->n_ext];
            memset(e, 0, sizeof(*e));
            for (int l = 0; l < REFP_NUM; l++) {
                for (int v = 0; v < 3; v++) { e->u.affine.cp[l][v][0] = mcore->affine_mv[l][v][MV_X]; e->u.affine.cp[l][v][1] = mcore->affine_mv[l][v][MV_Y]; }
                e->u.affine.mv_unref[l][0] = g->pend.have_aff ? g->pend.mv_unref[l][MV_X] : core->mv[l][MV_X];
                e->u.affine.mv_unref[l][1] = g->pend.have_aff ? g->pend.mv_unref[l][MV_Y] : core->mv[l][MV_Y];
            }
            const uint32_t ei = (uint32_t)g->n_ext++;
            memcpy(c->mv[1], &ei, 4);
        } else {
            c->mode = XB200_MODE_INTER;
            for (int l = 0; l < REFP_NUM; l++) { c->mv[l][0] = core->mv[l][MV_X]; c->mv[l][1] = core->mv[l][MV_Y]; }
            if (g->pend.dmvr && ctx->sps->tool_dmvr) c->flags |= XB200_CUF_DMVR;
        }
        if (ats_inter_info && ctx->sps->tool_ats)
            c->ats = (uint8_t)(((ats_inter_info & 7) << 2) | (((ats_inter_info >> 4) & 1) << 5));
    }
    /* coded-block flags: nnz_sub per 64x64 sub-block (src_base/xevd_eco.c:618-625); planes the CU does not carry stay 0 */
    int tlw = log2w, tlh = log2h;
    if (!intra && !ibc && ats_inter_info && ctx->sps->tool_ats) xevdm_get_tu_size(ats_inter_info, log2w, log2h, &tlw, &tlh);
    const int big = log2w > MAX_TR_LOG2 || log2h > MAX_TR_LOG2;
    for (int p = 0; p < N_C; p++) {
        if (p == Y_C ? !do_l : !do_c) continue;
        int bits = 0;
        if (big) { for (int sb = 0; sb < MAX_SUB_TB_NUM; sb++) bits |= (core->is_coef_sub[p][sb] ? 1 : 0) << sb; }
        else bits = core->is_coef[p] ? 1 : 0;
        if (core->pred_mode == MODE_SKIP) bits = 0;
        c->cbf |= (uint16_t)(bits << (4 * p));
    }
    /* coefficient blocks: CU-raster (what coef_rect_to_series left in core->coef; an ats_inter CU holds its TU only), planes without
     * coefficients absent, every block padded to a multiple of 8 entries */
    c->coef_off = (uint32_t)g->n_coef;
    for (int p = 0; p < N_C; p++) {
        if (!((c->cbf >> (4 * p)) & 15)) continue;
        const int n = (1 << (tlw + tlh)) >> (p ? 2 : 0), n8 = (n + 7) & ~7;
        if (g->n_coef + (size_t)n8 > g->cap_coef) {
            size_t nc = g->cap_coef ? g->cap_coef * 2 : (size_t)1 << 20;
            while (nc < g->n_coef + (size_t)n8) nc *= 2;
            int16_t *q = (int16_t *)realloc(g->coef, nc * sizeof(int16_t));
            if (!q) { g->err = XEVD_ERR_OUT_OF_MEMORY; return; }
            g->coef = q; g->cap_coef = nc;
        }
        memcpy(g->coef + g->n_coef, coef[p], (size_t)n * sizeof(int16_t));
        if (n8 > n) memset(g->coef + g->n_coef + n, 0, (size_t)(n8 - n) * sizeof(int16_t));
        g->n_coef += (size_t)n8;
    }
    g->last_cu = g->n_cu++;
    g->chunk[g->n_chunk - 1].cu1 = g->n_cu;
    g->chunk[g->n_chunk - 1].coef1 = g->n_coef;
    memset(&g->pend, 0, sizeof(g->pend));
}

/* xevdm_htdf (src_main/xevdm_recon.c:299): runs in the wavefront kernel; the availability word the reference computed for it
 * (xevdm.c:1385) is order-derived and travels with the CU */
void glue_htdf(s16 *rec, int qp, int w, int h, int s, BOOL intra_block_flag, pel *rec_pic, int s_pic, int avail_cu, int scup, int w_scu, int h_scu,
               u32 *map_scu, int constrained_intra_pred, int bit_depth)
{
    (void)rec; (void)qp; (void)w; (void)h; (void)s; (void)intra_block_flag; (void)rec_pic; (void)s_pic; (void)w_scu; (void)h_scu;
    (void)constrained_intra_pred; (void)bit_depth;
    /* ctx is not among the arguments: map_scu + scup identifies the instance */
    for (int i = 0; i < 8; i++) {
        GLUE *g = g_inst[i];
        if (g && g->ctx->map_scu == map_scu && g->last_cu >= 0 && g->last_cu < g->n_cu) {
            XB200_CU *c = &g->cus[g->last_cu];
            if ((int)(((c->y >> 2) * g->ctx->w_scu) + (c->x >> 2)) == scup) c->avail_cu = (uint16_t)avail_cu;
        }
    }
}

/* ---- picture buffers: every XEVD_PIC gets a device twin ------------------------------------------------------------------ */
/* The plane buffers of a picture's XEVD_IMGB are the destination of one D2H copy per decoded picture, so they have to be page-locked.
 * Registering the buffers xevd_imgb_create malloc'ed (cudaHostRegister) cost 13 ms per 6 MB plane on the test boxes - 20 ms per 1080p
 * picture, most of the set-up time of a short stream.  The planes are therefore REPLACED by page-locked allocations (2-3 ms per picture):
 * baddr / a of the imgb and buf_* / y,u,v of the picture are re-based, a 16-byte malloc block stands in for every original buffer, and a
 * wrapper around imgb->release puts those stand-ins back (imgb_delete frees baddr with free(), src_base/xevd_util.c:102-118) and frees
 * the page-locked planes when the last reference goes - which can be after the picture manager and the decoder are gone. */
#include <pthread.h>
typedef struct PIN_REC { XEVD_IMGB *imgb; void *pinned[3]; void *stub[3]; int (*release)(XEVD_IMGB *); struct PIN_REC *next; } PIN_REC;
static PIN_REC *g_pins;
static pthread_mutex_t g_pins_lock = PTHREAD_MUTEX_INITIALIZER;
static int glue_imgb_release(XEVD_IMGB *imgb)
{
    PIN_REC *r = NULL, **pp;
    pthread_mutex_lock(&g_pins_lock);
    for (pp = &g_pins; *pp; pp = &(*pp)->next) if ((*pp)->imgb == imgb) { r = *pp; break; }
    if (!r) { pthread_mutex_unlock(&g_pins_lock); return XEVD_ERR_UNEXPECTED; }
    int (*release)(XEVD_IMGB *) = r->release;
    if (imgb->getref(imgb) <= 1) {           /* the last reference: hand the imgb back the way xevd_imgb_create built it */
        *pp = r->next;
        pthread_mutex_unlock(&g_pins_lock);
        for (int k = 0; k < 3; k++)
            if (r->pinned[k]) { imgb->baddr[k] = r->stub[k]; imgb->a[k] = r->stub[k]; xb200_host_free(r->pinned[k]); }
        free(r);
    } else pthread_mutex_unlock(&g_pins_lock);
    return release(imgb);
}
/* returns the number of planes now page-locked */
static int pin_planes(XEVD_PIC *pic)
{
    XEVD_IMGB *imgb = pic->imgb;
    if (!imgb || !imgb->release) return 0;
    PIN_REC *r = (PIN_REC *)calloc(1, sizeof(PIN_REC));
    if (!r) return 0;
    int n = 0;
    for (int k = 0; k < 3; k++) {
        if (!imgb->baddr[k] || imgb->bsize[k] <= 0) continue;
        void *stub = malloc(16), *pinned = stub ? xb200_host_alloc((size_t)imgb->bsize[k]) : NULL;
        if (!pinned) { free(stub); continue; }
        const ptrdiff_t off = (unsigned char *)imgb->a[k] - (unsigned char *)imgb->baddr[k];
        free(imgb->baddr[k]);
        imgb->baddr[k] = pinned; imgb->a[k] = (unsigned char *)pinned + off;
        r->pinned[k] = pinned; r->stub[k] = stub;
        n++;
    }
    if (!n) { free(r); return 0; }
    pic->buf_y = imgb->baddr[0]; pic->buf_u = imgb->baddr[1]; pic->buf_v = imgb->baddr[2];
    pic->y = imgb->a[0]; pic->u = imgb->a[1]; pic->v = imgb->a[2];
    r->imgb = imgb; r->release = imgb->release;
    imgb->release = glue_imgb_release;
    pthread_mutex_lock(&g_pins_lock);
    r->next = g_pins; g_pins = r;
    pthread_mutex_unlock(&g_pins_lock);
    return n;
}

XEVD_PIC *glue_picbuf_alloc(PICBUF_ALLOCATOR *pa, int *ret, int bitdepth)
{
    XEVD_PIC *pic = xevdm_picbuf_alloc(pa, ret, bitdepth);
    if (!pic) return NULL;
    /* xevdm_picman_init copies ctx->pa into the picture manager (src_main/xevdm_picman.c:709); pdata[0] names the instance */
    GLUE *g = NULL;
    for (int i = 0; i < 8; i++) if (g_inst[i] && (void *)g_inst[i] == pa->pdata[0]) g = g_inst[i];
    GLUE_PIC *slot = g ? pic_of(g, NULL) : NULL;
    int err = 0;
    const double t0 = now_s();
    if (slot) {
        slot->dev = xb200_pic_alloc(g->dev, pa->w, pa->h, &err);
        if (slot->dev) {
            slot->host = pic;
            /* every plane buffer of the XEVD_IMGB (xevd_imgb_generate allocates them one by one): page-locked, the copy in
             * glue_picbuf_expand is a true asynchronous DMA; a pageable plane would make that call wait for the device */
            for (int k = 0; k < 3; k++) slot->registered[k] = NULL;
            if (pin_planes(pic) < 3)        /* (a plane that could not be replaced: register it in place; failing that the copy is a staged DMA) */
                for (int k = 0; k < 3; k++) {
                    int replaced = 0;
                    pthread_mutex_lock(&g_pins_lock);
                    for (PIN_REC *r = g_pins; r; r = r->next) if (r->imgb == pic->imgb && r->pinned[k]) replaced = 1;
                    pthread_mutex_unlock(&g_pins_lock);
                    if (!replaced && pic->imgb && pic->imgb->baddr[k] && pic->imgb->bsize[k] > 0 &&
                        xb200_host_register(pic->imgb->baddr[k], (size_t)pic->imgb->bsize[k]) == XB200_OK)
                        slot->registered[k] = pic->imgb->baddr[k];
                }
            g->t_alloc += now_s() - t0;
            return pic;
        }
    }
    fprintf(stderr, "[xevd-b200] no device picture for %dx%d (%d)\n", pa->w, pa->h, err);
    xevdm_picbuf_free(pa, pic);
    if (ret) *ret = XEVD_ERR_OUT_OF_MEMORY;
    return NULL;
}
void glue_picbuf_free(PICBUF_ALLOCATOR *pa, XEVD_PIC *pic)
{
    for (int i = 0; i < 8; i++) {
        GLUE *g = g_inst[i];
        GLUE_PIC *s = (g && pic) ? pic_of(g, pic) : NULL;
        if (!s) continue;
        xb200_sync(g->dev);
        for (int k = 0; k < 3; k++) if (s->registered[k]) xb200_host_unregister(s->registered[k]);
        xb200_pic_free(g->dev, s->dev);
        memset(s, 0, sizeof(*s));
    }
    xevdm_picbuf_free(pa, pic);
}

/* ---- slice: entropy decode + hooked walk on the host, one reconstruction call on the device ------------------------------- */
static void fill_params(GLUE *g)
{
    XEVD_CTX *ctx = g->ctx;
    XEVDM_CTX *mctx = (XEVDM_CTX *)ctx;
    const XEVD_SPS *sps = ctx->sps;
    XB200_PARAMS *p = &g->prm;
    memset(p, 0, sizeof(*p));
    p->w = ctx->w; p->h = ctx->h;
    p->bit_depth_luma = sps->bit_depth_luma_minus8 + 8; p->bit_depth_chroma = sps->bit_depth_chroma_minus8 + 8;
    p->chroma_format_idc = sps->chroma_format_idc;
    p->log2_ctu = ctx->log2_max_cuwh;
    p->tool_admvp = sps->tool_admvp; p->tool_iqt = sps->tool_iqt; p->tool_ats = sps->tool_ats; p->tool_addb = sps->tool_addb;
    p->tool_alf = sps->tool_alf; p->tool_htdf = sps->tool_htdf; p->tool_dmvr = sps->tool_dmvr; p->tool_eipd = sps->tool_eipd;
    p->tool_affine = sps->tool_affine; p->tool_ibc = sps->ibc_flag;
    p->slice_qp = ctx->sh.qp; p->qp_u_offset = ctx->sh.qp_u_offset; p->qp_v_offset = ctx->sh.qp_v_offset;
    p->deblock_alpha_offset = ctx->sh.sh_deblock_alpha_offset; p->deblock_beta_offset = ctx->sh.sh_deblock_beta_offset;
    p->poc = ctx->poc.poc_val;
    p->constrained_intra_pred = ctx->pps.constrained_intra_pred_flag;
    p->tool_suco = sps->sps_suco_flag;
    g->n0 = g->n1 = 0;
    if (ctx->sh.slice_type != SLICE_I) {
        for (int i = 0; i < mctx->dpm.num_refp[REFP_0] && i < XEVD_MAX_NUM_REF_PICS; i++) {
            GLUE_PIC *s = pic_of(g, ctx->refp[i][REFP_0].pic);
            if (!s) break;
            g->l0[g->n0++] = s->dev;
            xb200_pic_set_poc(s->dev, ctx->refp[i][REFP_0].poc);
        }
        if (ctx->sh.slice_type == SLICE_B)
            for (int i = 0; i < mctx->dpm.num_refp[REFP_1] && i < XEVD_MAX_NUM_REF_PICS; i++) {
                GLUE_PIC *s = pic_of(g, ctx->refp[i][REFP_1].pic);
                if (!s) break;
                g->l1[g->n1++] = s->dev;
                xb200_pic_set_poc(s->dev, ctx->refp[i][REFP_1].poc);
            }
    }
}

int glue_dec_slice(XEVD_CTX *ctx, XEVD_CORE *core)
{
    GLUE *g = glue_of(ctx);
    if (!g) return XEVD_ERR_UNEXPECTED;
    if (ctx->sps->chroma_format_idc != 1) return XEVD_ERR_UNSUPPORTED_COLORSPACE;
    {   /* the tile grid (set_tile_info, src_main/xevdm.c:2162): the loop filters need it, reconstruction does not */
        uint16_t col_bd[XB200_MAX_TILE_COLS + 1] = {0}, row_bd[XB200_MAX_TILE_ROWS + 1] = {0};
        if (ctx->w_tile > XB200_MAX_TILE_COLS || ctx->h_tile > XB200_MAX_TILE_ROWS) return XEVD_ERR_UNSUPPORTED;
        for (int i = 0; i < ctx->w_tile; i++) col_bd[i + 1] = (uint16_t)(col_bd[i] + ctx->tile[i].w_ctb);
        for (int j = 0; j < ctx->h_tile; j++) row_bd[j + 1] = (uint16_t)(row_bd[j] + ctx->tile[j * ctx->w_tile].h_ctb);
        if (xb200_set_tiles(g->dev, ctx->w_tile, col_bd, ctx->h_tile, row_bd, ctx->pps.loop_filter_across_tiles_enabled_flag) < 0) {
            dev_fail(g, XEVD_ERR, "xb200_set_tiles");
            return g->err;
        }
    }
    if (g->maps_pending) { xb200_sync(g->dev); g->maps_pending = 0; }     /* refined vectors of the previous picture are in place */
    g->n_cu = 0; g->n_ext = 0; g->n_coef = 0; g->n_chunk = 0; g->last_cu = -1;
    memset(&g->pend, 0, sizeof(g->pend));
    g->dbk_done = g->alf_done = 0;
    const double t_slice = now_s();
    int ret = g->ref_dec_slice(ctx, core);                 /* xevdm_dec_slice: entropy decode, then the walk through the hooks above */
    g->t_host += now_s() - t_slice;
    if (XEVD_FAILED(ret)) return ret;
    if (g->err) return g->err;
    GLUE_PIC *cur = pic_of(g, ctx->pic);
    if (!cur) return XEVD_ERR_UNEXPECTED;
    fill_params(g);
    /* per-CTU index in raster order.  One tile: the walk already is raster.  Chunks out of raster order (tiles) are permuted,
     * coefficient blocks with them, because a CTU's blocks must be one contiguous range of the stream (include/xevd_b200.h) */
    const int n_ctu = ctx->f_lcu;
    if (grow((void **)&g->ctu_first, &g->cap_ctu, n_ctu + 1, sizeof(uint32_t))) return XEVD_ERR_OUT_OF_MEMORY;
    int sorted = 1;
    for (int k = 1; k < g->n_chunk; k++) if (g->chunk[k].ctu <= g->chunk[k - 1].ctu) sorted = 0;
    const XB200_CU *cus = g->cus; const int16_t *coef = g->coef;
    if (sorted) {
        int k = 0;
        for (int t = 0; t <= n_ctu; t++) {
            while (k < g->n_chunk && g->chunk[k].ctu < t) k++;
            g->ctu_first[t] = (uint32_t)(k < g->n_chunk ? g->chunk[k].cu0 : g->n_cu);
        }
    } else {
        if (grow((void **)&g->cus2, &g->cap_cu2, g->n_cu + 1, sizeof(XB200_CU))) return XEVD_ERR_OUT_OF_MEMORY;
        if (g->cap_coef2 < g->n_coef + 8) {
            int16_t *q = (int16_t *)realloc(g->coef2, (g->n_coef + 8) * 2 * sizeof(int16_t));
            if (!q) return XEVD_ERR_OUT_OF_MEMORY;
            g->coef2 = q; g->cap_coef2 = (g->n_coef + 8) * 2;
        }
        int *of_ctu = (int *)malloc(sizeof(int) * (size_t)n_ctu);
        if (!of_ctu) return XEVD_ERR_OUT_OF_MEMORY;
        for (int t = 0; t < n_ctu; t++) of_ctu[t] = -1;
        for (int k = 0; k < g->n_chunk; k++) of_ctu[g->chunk[k].ctu] = k;
        int nc = 0; size_t nk = 0;
        for (int t = 0; t < n_ctu; t++) {
            g->ctu_first[t] = (uint32_t)nc;
            if (of_ctu[t] < 0) continue;
            const GLUE_CHUNK *ch = &g->chunk[of_ctu[t]];
            for (int i = ch->cu0; i < ch->cu1; i++) { g->cus2[nc] = g->cus[i]; g->cus2[nc].coef_off = (uint32_t)(g->cus[i].coef_off - ch->coef0 + nk); nc++; }
            memcpy(g->coef2 + nk, g->coef + ch->coef0, (ch->coef1 - ch->coef0) * sizeof(int16_t));
            nk += ch->coef1 - ch->coef0;
        }
        g->ctu_first[n_ctu] = (uint32_t)nc;
        free(of_ctu);
        cus = g->cus2; coef = g->coef2;
    }
    {   /* debugging aid: XEVD_B200_DUMP=<directory> writes the work lists of every slice (tools/glue_dump.py reads them back) */
        const char *dir = getenv("XEVD_B200_DUMP");
        if (dir) {
            char fn[512];
            snprintf(fn, sizeof(fn), "%s/slice_%04d.bin", dir, g_dump_serial++);
            FILE *fp = fopen(fn, "wb");
            if (fp) {
                int32_t hdr[8] = { g->n_cu, n_ctu, g->n_ext, (int32_t)g->n_coef, g->n0, g->n1, ctx->sh.slice_type, ctx->sh.deblocking_filter_on };
                int32_t pocs[2 * XEVD_MAX_NUM_REF_PICS] = {0};
                for (int i = 0; i < g->n0; i++) pocs[i] = ctx->refp[i][REFP_0].poc;
                for (int i = 0; i < g->n1; i++) pocs[XEVD_MAX_NUM_REF_PICS + i] = ctx->refp[i][REFP_1].poc;
                fwrite(hdr, sizeof(hdr), 1, fp); fwrite(&g->prm, sizeof(g->prm), 1, fp); fwrite(pocs, sizeof(pocs), 1, fp);
                fwrite(cus, sizeof(XB200_CU), (size_t)g->n_cu, fp); fwrite(g->ctu_first, 4, (size_t)n_ctu + 1, fp);
                fwrite(g->ext, sizeof(XB200_CU_EXT), (size_t)g->n_ext, fp); fwrite(coef, 2, g->n_coef, fp);
                fclose(fp);
            }
        }
    }
    if (g->n_cu > 0) {
        {   /* the sequence's chroma QP mapping (xevd_set_chroma_qp_tbl_loc / xevd_derived_chroma_qp_mapping_tables, xevdm.c:471-486) */
            int32_t tbl[2][XEVD_MAX_QP_TABLE_SIZE];
            for (int k = 0; k < 2; k++) for (int i = 0; i < XEVD_MAX_QP_TABLE_SIZE; i++) tbl[k][i] = xevd_qp_chroma_dynamic[k][i];
            xb200_set_chroma_qp_table(g->dev, &tbl[0][0]);
        }
        const double t_r = now_s();
        int r = xb200_recon_frame(g->dev, &g->prm, cur->dev, g->l0, g->n0, g->l1, g->n1, cus, g->n_cu, g->ctu_first, n_ctu,
                                  g->ext, g->n_ext, coef, g->n_coef);
        g->t_recon += now_s() - t_r;
        if (r < 0) { fprintf(stderr, "[xevd-b200] xb200_recon_frame: %d, slice type %d, lists %d / %d\n", r, ctx->sh.slice_type, g->n0, g->n1); dev_fail(g, XEVD_ERR, "xb200_recon_frame"); return g->err; }
        g->n_cus_total += g->n_cu;
        dump_stage(g, cur, "recon");
    }
    return ret;
}

/* ctx->fn_deblock: the reference calls it per tile and per pass (src_main/xevdm.c:3152-3202); both passes of the whole picture are one
 * device call, launched by the first of them */
int glue_deblock(void *arg)
{
    XEVD_CORE *core = (XEVD_CORE *)arg;
    GLUE *g = glue_of(core->ctx);
    if (!g) return XEVD_ERR_UNEXPECTED;
    if (g->dbk_done) return XEVD_OK;
    g->dbk_done = 1;
    GLUE_PIC *cur = pic_of(g, core->ctx->pic);
    if (!cur) return XEVD_ERR_UNEXPECTED;
    fill_params(g);
    const double t_d = now_s();
    if (xb200_deblock(g->dev, &g->prm, cur->dev, g->l0, g->n0, g->l1, g->n1, NULL) < 0) { dev_fail(g, XEVD_ERR, "xb200_deblock"); return g->err; }
    g->t_dbk += now_s() - t_d;
    if (getenv("XEVD_B200_DUMP")) {
        char fn[512];
        snprintf(fn, sizeof(fn), "%s/pic_%04d_dbkprm.bin", getenv("XEVD_B200_DUMP"), g_dump_serial - 1);
        FILE *fp = fopen(fn, "wb");
        if (fp) { fwrite(&g->prm, sizeof(g->prm), 1, fp); fclose(fp); }
    }
    dump_stage(g, cur, "dbk");
    return XEVD_OK;
}

/* alf_process_tile (src_main/xevdm_alf.c:901): alf_process has reconstructed the coefficients from the APS on the host; the filter itself
 * is one device call for the picture.  The argument is alf_process's XEVD_ALF_TMP (xevdm_alf.c:796-803). */
typedef struct { ADAPTIVE_LOOP_FILTER *alf; CODING_STRUCTURE *cs; ALF_SLICE_PARAM *alf_slice_param; int tile_idx; int tsk_num; } GLUE_ALF_TMP;
int alf_process_tile(void *arg)
{
    GLUE_ALF_TMP *t = (GLUE_ALF_TMP *)arg;
    XEVD_CTX *ctx = (XEVD_CTX *)t->cs->ctx;
    GLUE *g = glue_of(ctx);
    if (!g) return XEVD_ERR_UNEXPECTED;
    if (g->alf_done) return XEVD_OK;
    g->alf_done = 1;
    GLUE_PIC *cur = pic_of(g, t->cs->pic);
    if (!cur) return XEVD_ERR_UNEXPECTED;
    XB200_ALF a;
    memset(&a, 0, sizeof(a));
    memcpy(a.coef_luma, t->alf->coef_final, sizeof(a.coef_luma));
    memcpy(a.coef_chroma, t->alf_slice_param->chroma_coef, sizeof(a.coef_chroma));
    for (int c = 0; c < 3; c++) a.enable[c] = t->alf_slice_param->enable_flag[c] ? 1 : 0;
    fill_params(g);
    if (xb200_alf(g->dev, &g->prm, cur->dev, &a, t->alf->ctu_enable_flag[0]) < 0) { dev_fail(g, XEVD_ERR, "xb200_alf"); return g->err; }
    return XEVD_OK;
}

/* ctx->fn_picbuf_expand: border replication on the device, then the picture starts its way to the host XEVD_IMGB that xevd_pull
 * will hand out; with DMVR the refined vectors come back too (temporal candidates of later pictures) */
void glue_picbuf_expand(XEVD_CTX *ctx, XEVD_PIC *pic)
{
    GLUE *g = glue_of(ctx);
    GLUE_PIC *s = g ? pic_of(g, pic) : NULL;
    if (!s) { if (g) dev_fail(g, XEVD_ERR_UNEXPECTED, "picture lookup"); return; }
    const double t_o = now_s();
    if (xb200_pad(g->dev, s->dev) < 0) { dev_fail(g, XEVD_ERR, "xb200_pad"); return; }
    if (xb200_pic_download(g->dev, s->dev, pic->y, pic->s_l, pic->u, pic->s_c, pic->v, pic->s_c) < 0) { dev_fail(g, XEVD_ERR, "xb200_pic_download"); return; }
    if (ctx->sps->tool_dmvr && ctx->sh.slice_type == SLICE_B) {
        if (xb200_pic_download_maps(g->dev, s->dev, (int16_t *)pic->map_mv, NULL, NULL) < 0) { dev_fail(g, XEVD_ERR, "xb200_pic_download_maps"); return; }
    } else g->maps_pending = 1;
    g->n_pictures++;
    g->t_out += now_s() - t_o;
}

/* xevd_imgb_generate is what the reference calls right before it reads a decoded picture on the host (MD5 check, src_main/xevdm.c:3269;
 * DRA copy on pull, :3378): the asynchronous copy has to have landed */
XEVD_IMGB *glue_imgb_generate(int w, int h, int padl, int padc, int idc, int bit_depth)
{
    for (int i = 0; i < 8; i++) if (g_inst[i]) { const double t0 = now_s(); xb200_sync(g_inst[i]->dev); g_inst[i]->t_wait += now_s() - t0; }
    return xevd_imgb_generate(w, h, padl, padc, idc, bit_depth);
}

/* ---- the public API (inc/xevd.h:369-374) ----------------------------------------------------------------------------------- */
XEVD xevd_create(XEVD_CDSC *cdsc, int *err)
{
    int slot = -1;
    for (int i = 0; i < 8; i++) if (!g_inst[i]) { slot = i; break; }
    if (!cdsc || slot < 0) { if (err) *err = XEVD_ERR_INVALID_ARGUMENT; return NULL; }
    GLUE *g = (GLUE *)calloc(1, sizeof(GLUE));
    if (!g) { if (err) *err = XEVD_ERR_OUT_OF_MEMORY; return NULL; }
    int derr = 0, dev_index = 0;
    const char *e = getenv("XEVD_B200_DEVICE");
    if (e) dev_index = atoi(e);
    g->dev = xb200_create(dev_index, &derr);
    if (!g->dev) {                                            /* no CUDA device: there is no CPU path behind this library */
        fprintf(stderr, "[xevd-b200] xb200_create(%d) failed: %d\n", dev_index, derr);
        free(g);
        if (err) *err = XEVD_ERR;
        return NULL;
    }
    /* the host side of a slice is entropy decoding plus bookkeeping in decoding order: one task (the worker threads of the reference
     * exist to spread pixel work, which has left the host) */
    XEVD_CDSC one = *cdsc;
    one.threads = 1;
    XEVD id = xevdref_create(&one, err);
    if (!id) { xb200_destroy(g->dev); free(g); return NULL; }
    XEVD_CTX *ctx = (XEVD_CTX *)id;
    XEVDM_CTX *mctx = (XEVDM_CTX *)ctx;
    (void)mctx;
    g->ctx = ctx;
    g->ref_dec_slice = ctx->fn_dec_slice;
    ctx->fn_dec_slice = glue_dec_slice;
    ctx->fn_deblock = glue_deblock;
    ctx->fn_picbuf_expand = glue_picbuf_expand;
    ctx->pa.pdata[0] = g;
    g->last_cu = -1;
    g_inst[slot] = g;
    return id;
}

void xevd_delete(XEVD id)
{
    GLUE *g = glue_of((XEVD_CTX *)id);
    if (g) xb200_sync(g->dev);
    xevdref_delete(id);                                       /* frees the pictures through glue_picbuf_free */
    if (g) {
        for (int i = 0; i < 8; i++) if (g_inst[i] == g) g_inst[i] = NULL;
        if (getenv("XEVD_B200_STATS"))
            fprintf(stderr, "[xevd-b200] %lld pictures, %lld CUs, %lld kernel launches; host seconds: entropy + walk %.4f, recon call %.4f, deblock call %.4f, "
                    "pad + download calls %.4f, waiting for the device %.4f, picture allocation %.4f\n", g->n_pictures, g->n_cus_total, xb200_launch_count(g->dev),
                    g->t_host, g->t_recon, g->t_dbk, g->t_out, g->t_wait, g->t_alloc);
        xb200_destroy(g->dev);
        free(g->cus); free(g->ext); free(g->coef); free(g->chunk); free(g->ctu_first); free(g->cus2); free(g->coef2);
        free(g);
    }
}

int xevd_decode(XEVD id, XEVD_BITB *bitb, XEVD_STAT *stat)
{
    GLUE *g = glue_of((XEVD_CTX *)id);
    if (g) g->err = 0;
    int ret = xevdref_decode(id, bitb, stat);
    if (g && g->err && XEVD_SUCCEEDED(ret)) ret = g->err;
    return ret;
}

int xevd_pull(XEVD id, XEVD_IMGB **imgb)
{
    GLUE *g = glue_of((XEVD_CTX *)id);
    if (g) { const double t0 = now_s(); xb200_sync(g->dev); g->t_wait += now_s() - t0; }       /* the planes handed out are complete */
    return xevdref_pull(id, imgb);
}

int xevd_config(XEVD id, int cfg, void *buf, int *size) { return xevdref_config(id, cfg, buf, size); }
/* xevd_info is exported by the reference's xevd_util.c unchanged */

/* introspection for tests: launches issued by the instance's device context so far (proves the GPU did the work) */
long long xevd_b200_launch_count(XEVD id)
{
    GLUE *g = glue_of((XEVD_CTX *)id);
    return g ? xb200_launch_count(g->dev) : -1;
}
