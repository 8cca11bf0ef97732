/*
 * oracle/ -- CPU restatement of the XEVD reconstruction hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under xevd_b200/ may include, link or call this code; it
 * exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the CUDA
 * path.  Parity status: PINNED -- every function here is compared against the unmodified reference
 * compiled from /root/reference (oracle/_ref/libxevd_ref.so, built by oracle/Makefile) in
 * tests/test_oracle_vs_ref.py, and against the golden vectors generated from that build
 * (tests/golden/, script tests/golden/make_golden.py).  The reference ships no test vectors of its own
 * (SURVEY 4, 8c).
 *
 * Each function cites the reference file:line whose arithmetic it restates.  The code is written
 * from the arithmetic, in direct (matrix / loop) form; it is not a copy of the reference's
 * butterflies or SIMD.
 */
#ifndef ORC_COMMON_H
#define ORC_COMMON_H

#include <stdint.h>
#include <stddef.h>
#include "../include/xevd_b200.h"

typedef int16_t pel;

/* a host picture in the reference's padded layout (xevd_util.c:153-230, 250-363) */
typedef struct ORC_PIC {
    pel *y, *u, *v;          /* sample (0,0) of each plane                                    */
    int  s_l, s_c;           /* strides in pels                                               */
    int  w_l, h_l, w_c, h_c;
    int  pad_l, pad_c;
    int  poc;
    int16_t  *map_mv;        /* [h_scu*w_scu][2][2]                                           */
    int8_t   *map_refi;      /* [h_scu*w_scu][2]                                              */
    uint32_t *map_scu;       /* [h_scu*w_scu]                                                 */
    int  w_scu, h_scu;
    int16_t  *map_unrefined_mv; /* Main: the vectors before DMVR refinement (mctx->map_unrefined_mv); map_mv holds the refined ones */
} ORC_PIC;

static inline int orc_clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int orc_min(int a, int b) { return a < b ? a : b; }
static inline int orc_max(int a, int b) { return a > b ? a : b; }

/* orc_tables.c */
const int8_t  *orc_dct2_matrix(int log2n);          /* [n][n], row k = basis k (xevd_tbl_tm2..64)   */
const int16_t *orc_mc_luma_taps(int main_tables);   /* [16][8]                                       */
const int16_t *orc_mc_chroma_taps(int main_tables); /* [32][4]                                       */
const int16_t *orc_ats_matrix(int dst7, int log2n); /* inverse ATS matrices, [n][n] (xevd_tbl_inv_tr*) */
int orc_dq_scale(int qp, int iqt);

/* orc_mc.c */
void orc_mc_luma(const pel *ref, int s_ref, int gmv_x, int gmv_y, int ori_mv_x, int ori_mv_y,
                 pel *pred, int s_pred, int w, int h, int bit_depth, int main_tables);
void orc_mc_chroma(const pel *ref, int s_ref, int gmv_x, int gmv_y, int ori_mv_x, int ori_mv_y,
                   pel *pred, int s_pred, int w, int h, int bit_depth, int main_tables);
void orc_interp(const pel *ref, int s_ref, int ix, int iy, const int16_t *cx, const int16_t *cy,
                int fx, int fy, int ntap, pel *pred, int s_pred, int w, int h, int bd);
void orc_mv_clip(int x, int y, int pic_w, int pic_h, int w, int h, const int8_t refi[2],
                 const int16_t mv[2][2], int16_t mv_t[2][2]);
/* full inter prediction of one CU: pred[c] is w*h (luma) / (w/2)*(h/2) (chroma), CU raster */
void orc_inter_pred(const XB200_PARAMS *prm, int x, int y, int w, int h, const int8_t refi[2],
                    const int16_t mv[2][2], const ORC_PIC *const *refs_l0, const ORC_PIC *const *refs_l1,
                    pel *pred_y, pel *pred_u, pel *pred_v);

/* orc_dmvr.c: decoder-side motion vector refinement (Main, tool_dmvr).  Returns 1 and fills the two per-list predictions of all
 * three planes (pred[l][c], CU raster) plus the refined quarter-pel vectors per SCU (dmvr_mv[scu][l][xy], CU-raster SCU order) when
 * the CU is refined; returns 0 (nothing written) when xevdm_mc's conditions rule DMVR out. */
int orc_dmvr_pred(const XB200_PARAMS *prm, int x, int y, int w, int h, const int8_t refi[2], const int16_t mv[2][2],
                  const ORC_PIC *const *refs_l0, const ORC_PIC *const *refs_l1, pel *pred[2][3], int16_t *dmvr_mv);

/* orc_affine.c */
void orc_affine_subblock(const int16_t cp[2][3][2], const int8_t refi[2], int w, int h, int six, int *sub_w, int *sub_h, int *mem_ok);
void orc_affine_pred(const XB200_PARAMS *prm, int x, int y, int w, int h, const int8_t refi[2], const int16_t cp[2][3][2], int six,
                     const ORC_PIC *const *refs_l0, const ORC_PIC *const *refs_l1, pel *py, pel *pu, pel *pv);
void orc_affine_map_mv(const int16_t cp[2][3][2], const int8_t refi[2], int log2w, int log2h, int six, int l, int16_t *out);

/* orc_itdq.c */
void orc_dequant(int16_t *coef, int log2w, int log2h, int qp, int bit_depth, int iqt);
void orc_inv_dct2(int16_t *coef, int log2w, int log2h, int bit_depth, int iqt);
void orc_itdq_block(int16_t *coef, int log2w, int log2h, int qp, int bit_depth, int iqt);
void orc_inv_ats(int16_t *coef, int log2w, int log2h, int bit_depth, int ats_mode);
void orc_ats_inter_tu(int ats, int log2w, int log2h, int *tlw, int *tlh, int *xoff, int *yoff);
/* whole-CU residual (xevd_sub_block_itdq / xevdm_sub_block_itdq): coef[c] CU-raster, in place */
void orc_itdq_cu(const XB200_PARAMS *prm, const XB200_CU *cu, int16_t *cy, int16_t *cu_, int16_t *cv);

/* orc_recon.c */
int orc_recon_frame(const XB200_PARAMS *prm, ORC_PIC *cur,
                    const ORC_PIC *const *refs_l0, int n_l0, const ORC_PIC *const *refs_l1, int n_l1,
                    const XB200_CU *cus, int n_cu, const XB200_CU_EXT *ext, const int16_t *coef);
void orc_pad(ORC_PIC *pic);

/* orc_ipred.c */
void orc_intra_neighbours(const pel *rec, int s, int w, int h, int unit, uint64_t up_mask, uint64_t left_mask, int up_left_avail,
                          int bit_depth, pel *up, pel *left);
void orc_ipred_base(const pel *left, const pel *up, pel *dst, int mode, int w, int h);
void orc_intra_neighbours_main(const pel *rec, int s, int w, int h, int unit, uint64_t up_mask, uint64_t left_mask, uint64_t right_mask,
                               int up_left_avail, int bit_depth, pel *up, pel *left, pel *right);
void orc_ipred_main(const pel *left, const pel *up, const pel *right, int avail_lr, pel *dst, int ipm, int w, int h, int bit_depth);
void orc_ipred_uv_main(const pel *left, const pel *up, const pel *right, int avail_lr, pel *dst, int ipm_c, int ipm, int w, int h, int bit_depth);

/* orc_htdf.c */
void orc_htdf(pel *rec, int s, int w, int h, int qp, int intra, int avail, int bit_depth, const uint32_t *map_scu, int w_scu, int constrained);
const uint8_t *orc_htdf_table(int idx);

/* orc_output.c */
void orc_dra_apply(ORC_PIC *pic, const XB200_DRA *d);
void orc_output(const ORC_PIC *pic, int out_bits, int crop_l, int crop_r, int crop_t, int crop_b, void *y, int sy, void *u, int su, void *v, int sv);

/* orc_alf.c */
/* PPS tile grid for the two loop filters that follow (xb200_set_tiles has the same arguments): column / row boundaries in CTUs and
 * pps.loop_filter_across_tiles_enabled_flag; n_cols == n_rows == 1 = one tile (the initial state) */
typedef struct { int n_cols, n_rows, across; uint16_t col_bd[XB200_MAX_TILE_COLS + 1], row_bd[XB200_MAX_TILE_ROWS + 1]; } ORC_TILES;
void orc_set_tiles(int n_cols, const uint16_t *col_bd, int n_rows, const uint16_t *row_bd, int across);
const ORC_TILES *orc_tiles(void);
/* 1 when luma position `pos` (x for vertical edges, y for horizontal ones) lies on a boundary between two tiles */
int orc_on_tile_boundary(const ORC_TILES *t, int log2_ctu, int pos, int vertical);
int orc_alf_frame(const XB200_PARAMS *prm, ORC_PIC *pic, const XB200_ALF *alf, const uint8_t *ctb_flag_luma);

/* orc_df.c */
int orc_deblock_frame(const XB200_PARAMS *prm, ORC_PIC *pic, const XB200_CU *cus, int n_cu, const int *chroma_qp_tbl);
const uint8_t *orc_df_strength_table(void);
int orc_deblock_frame_addb(const XB200_PARAMS *prm, ORC_PIC *pic, const XB200_CU *cus, int n_cu, const int *chroma_qp_tbl,
                           const int *ref_id_l0, const int *ref_id_l1);
const uint8_t *orc_addb_tables(int which);

#endif
