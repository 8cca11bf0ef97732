/*
 * Deblocking filter, Baseline arithmetic (sps->tool_addb == 0).  TEST INFRASTRUCTURE ONLY (orc_common.h).
 *
 * Restates, for a picture that is one tile and one slice:
 *   driver      xevdm_deblock / deblock_tree            src_main/xevdm.c:1935-2103 (Baseline lib: src_base/xevd.c:1057-1243)
 *   CU walkers  deblock_cu_ver / deblock_cu_hor         src_main/xevdm_df.c:106-360  (= xevd_deblock_cu_*, src_base/xevd_df.c:291-545)
 *   strength    xevdm_get_tbl_qp_to_st                  src_main/xevdm_df.c:38-104
 *   filters     deblock_scu_{hor,ver}(_chroma)          src_base/xevd_df.c:96-289 (T6: truncating division)
 *
 * Pass 1 filters vertical edges (left edge of every CU whose left neighbour is already visited, right edge if the
 * right neighbour is), pass 2 the top edges; "visited" is the COD bit of map_scu, cleared before each pass.
 * CUs wider / taller than 64 are visited as two halves (the 64-sample transform boundary is an edge).
 */
#include <string.h>
#include <stdlib.h>
#include "orc_common.h"

/* The ADDB walkers compare the vectors BEFORE DMVR refinement: deblock_tree hands mctx->map_unrefined_mv to xevdm_deblock_cu_hor/ver
 * (src_main/xevdm.c:2009-2041) and deblock_addb_cu_* use that argument - but xevdm_deblock first copies map_mv over
 * map_unrefined_mv for every SCU WITHOUT the DMVR flag (src_main/xevdm.c:2077-2090), so what ADDB sees is: the unrefined vector of a
 * DMVR-refined CU, and map_mv - i.e. the affine SUB-BLOCK vectors xevdm_set_affine_mvf wrote - everywhere else (found on generated
 * Main streams: an edge along an affine CU changes strength from sub-block to sub-block).  The Baseline-filter walkers ignore the
 * argument (src_main/xevdm_df.c:1143-1166) and read ctx->map_mv (:111-124,207-208), the vectors AFTER refinement.  Both reproduced. */
#define DF_MV_AT(p, scu) (((p)->map_unrefined_mv && (((p)->map_scu[scu] >> 25) & 1)) ? (p)->map_unrefined_mv + 4 * (scu) : (p)->map_mv + 4 * (scu))

/* xevd_tbl_df_st (src_base/xevd_tbl.c:306-324) */
static const uint8_t k_df_st[4][52] = {
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 12, 12, 12, 12, 12},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 11, 11, 11, 11, 11},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 4, 5, 6, 7, 8, 9, 10, 10, 10, 10, 10},
    {0},
};
const uint8_t *orc_df_strength_table(void) { return &k_df_st[0][0]; }

/* rows are contiguous in memory, so an index past 51 reads the next row (what the reference does when the chroma
 * QP table maps above 51); past the whole table we define 0 */
static int st_lookup(int idx, int q)
{
    int f = idx * 52 + q;
    return (q < 0 || f >= 4 * 52) ? 0 : (&k_df_st[0][0])[f];
}

static int strength_class(uint32_t m0, uint32_t m1, const int8_t *r0, const int8_t *r1, const int16_t *mv0, const int16_t *mv1)
{
    if (((m0 >> 15) & 1) || ((m1 >> 15) & 1)) return 0;                 /* MCU_GET_IF  */
    if (((m0 >> 24) & 1) || ((m1 >> 24) & 1)) return 1;                 /* MCU_GET_CBFL */
    if (((m0 >> 26) & 1) || ((m1 >> 26) & 1)) return 2;                 /* MCU_GET_IBC  */
    int a[2][2], b[2][2];
    for (int l = 0; l < 2; l++)
        for (int d = 0; d < 2; d++) {
            a[l][d] = r0[l] >= 0 ? mv0[l * 2 + d] : 0;
            b[l][d] = r1[l] >= 0 ? mv1[l * 2 + d] : 0;
        }
#define FAR(p, q) (abs((p)[0] - (q)[0]) >= 4 || abs((p)[1] - (q)[1]) >= 4)
    if (r0[0] == r1[0] && r0[1] == r1[1]) return (FAR(a[0], b[0]) || FAR(a[1], b[1])) ? 2 : 3;
    if (r0[0] == r1[1] && r0[1] == r1[0]) return (FAR(a[0], b[1]) || FAR(a[1], b[0])) ? 2 : 3;
#undef FAR
    return 2;
}

/* one sample line across an edge: p[-2s] p[-s] | p[0] p[s]   (deblock_scu_hor/ver, xevd_df.c:96-134,193-231) */
static void filt_luma(pel *p, int s, int st, int maxv)
{
    int16_t A = p[-2 * s], B = p[-s], C = p[0], D = p[s];
    int16_t d = (int16_t)((A - (B << 2) + (C << 2) - D) / 8);            /* C division: truncates toward zero (T6) */
    int16_t ad = (int16_t)abs(d);
    int16_t t16 = (int16_t)orc_max(0, (ad - st) << 1);
    int16_t clip = (int16_t)orc_max(0, ad - t16);
    int16_t d1 = d < 0 ? -clip : clip;
    clip >>= 1;
    int16_t d2 = (int16_t)orc_clip3(-clip, clip, (A - D) / 4);
    p[-2 * s] = (pel)orc_clip3(0, maxv, (int16_t)(A - d2));
    p[-s] = (pel)orc_clip3(0, maxv, (int16_t)(B + d1));
    p[0] = (pel)orc_clip3(0, maxv, (int16_t)(C - d1));
    p[s] = (pel)orc_clip3(0, maxv, (int16_t)(D + d2));
}
static void filt_chroma(pel *p, int s, int st, int maxv)
{
    int16_t A = p[-2 * s], B = p[-s], C = p[0], D = p[s];
    int16_t d = (int16_t)((A - (B << 2) + (C << 2) - D) / 8);
    int16_t ad = (int16_t)abs(d);
    int16_t t16 = (int16_t)orc_max(0, (ad - st) << 1);
    int16_t clip = (int16_t)orc_max(0, ad - t16);
    int16_t d1 = d < 0 ? -clip : clip;
    p[-s] = (pel)orc_clip3(0, maxv, (int16_t)(B + d1));
    p[0] = (pel)orc_clip3(0, maxv, (int16_t)(C - d1));
}

typedef struct {
    const XB200_PARAMS *prm;
    ORC_PIC *pic;
    const int *cq[2];        /* chroma QP mapping for qp >= 0 (58 entries each); identity below 0 */
    uint8_t *cod;
    int planes;              /* XB200_CUF_LUMA | XB200_CUF_CHROMA of the CU being visited (tree_cons, xevdm_df.c:155-160,245-250) */
} DfCtx;

static int chroma_map(const DfCtx *c, int k, int q) { return q < 0 ? q : c->cq[k][q]; }

/* one 4-sample segment of an edge between SCU `cur` (right / below) and SCU `nb` (left / above) */
static void edge_segment(DfCtx *c, int cur, int nb, int x, int y, int vertical)
{
    ORC_PIC *p = c->pic;
    const int bdl = c->prm->bit_depth_luma, bdc = c->prm->bit_depth_chroma;
    const int cls = strength_class(p->map_scu[cur], p->map_scu[nb], p->map_refi + 2 * cur, p->map_refi + 2 * nb,
                                   p->map_mv + 4 * cur, p->map_mv + 4 * nb);          /* ctx->map_mv: see DF_MV */
    const int qp = (p->map_scu[cur] >> 16) & 0x7f;                       /* QP of the CURRENT side only (T7) */
    const int st = st_lookup(cls, qp) << (bdl - 8);
    if (st && (c->planes & XB200_CUF_LUMA)) {
        pel *q = p->y + y * p->s_l + x;
        for (int i = 0; i < 4; i++) vertical ? filt_luma(q + i * p->s_l, 1, st, (1 << bdl) - 1) : filt_luma(q + i, p->s_l, st, (1 << bdl) - 1);
    }
    if (!(c->planes & XB200_CUF_CHROMA)) return;
    const int qu = orc_clip3(-6 * (bdc - 8), 57, qp + c->prm->qp_u_offset), qv = orc_clip3(-6 * (bdc - 8), 57, qp + c->prm->qp_v_offset);
    const int st_u = st_lookup(cls, chroma_map(c, 0, qu)) << (bdc - 8), st_v = st_lookup(cls, chroma_map(c, 1, qv)) << (bdc - 8);
    for (int k = 0; k < 2; k++) {
        const int s = k ? st_v : st_u;
        if (!s) continue;
        pel *q = (k ? p->v : p->u) + (y >> 1) * p->s_c + (x >> 1);
        for (int i = 0; i < 2; i++) vertical ? filt_chroma(q + i * p->s_c, 1, s, (1 << bdc) - 1) : filt_chroma(q + i, p->s_c, s, (1 << bdc) - 1);
    }
}

/* ---- PPS tile grid shared by the loop-filter restatements (orc_set_tiles) ------------------------------------------------------ */
static ORC_TILES g_tiles = {1, 1, 0, {0, 0xffff}, {0, 0xffff}};
void orc_set_tiles(int n_cols, const uint16_t *col_bd, int n_rows, const uint16_t *row_bd, int across)
{
    g_tiles.n_cols = n_cols; g_tiles.n_rows = n_rows; g_tiles.across = across != 0;
    for (int i = 0; i <= n_cols; i++) g_tiles.col_bd[i] = col_bd[i];
    for (int i = 0; i <= n_rows; i++) g_tiles.row_bd[i] = row_bd[i];
}
const ORC_TILES *orc_tiles(void) { return &g_tiles; }
int orc_on_tile_boundary(const ORC_TILES *t, int log2_ctu, int pos, int vertical)
{
    if (pos & ((1 << log2_ctu) - 1)) return 0;
    const int c = pos >> log2_ctu, n = vertical ? t->n_cols : t->n_rows;
    const uint16_t *bd = vertical ? t->col_bd : t->row_bd;
    for (int i = 1; i < n; i++) if (bd[i] == c) return 1;
    return 0;
}
/* an edge between two tiles is filtered only with loop_filter_across_tiles_enabled_flag (no_boundary, xevdm_df.c:142,233,274,877,1088,1106) */
static int tile_cut(const XB200_PARAMS *prm, int pos, int vertical)
{
    return !g_tiles.across && orc_on_tile_boundary(&g_tiles, prm->log2_ctu, pos, vertical);
}

static void visit(DfCtx *c, int x, int y, int w, int h, int pass)
{
    ORC_PIC *p = c->pic;
    const int ws = p->w_scu, sx = x >> 2, sy = y >> 2, nw = w >> 2, nh = h >> 2;
    const int t = sy * ws + sx;
    if (pass == 0) {
        if (x > 0 && c->cod[t - 1] && !tile_cut(c->prm, x, 1))
            for (int i = 0; i < nh; i++) edge_segment(c, t + i * ws, t + i * ws - 1, x, y + 4 * i, 1);
        if (x + w < p->w_l && c->cod[t + nw] && !tile_cut(c->prm, x + w, 1))
            for (int i = 0; i < nh; i++) edge_segment(c, t + i * ws + nw, t + i * ws + nw - 1, x + w, y + 4 * i, 1);
    } else if (y > 0 && !tile_cut(c->prm, y, 0)) {
        for (int i = 0; i < nw; i++) edge_segment(c, t + i, t + i - ws, x + 4 * i, y, 0);
    }
    for (int j = 0; j < nh; j++) memset(c->cod + t + j * ws, 1, nw);
}

int orc_deblock_frame(const XB200_PARAMS *prm, ORC_PIC *pic, const XB200_CU *cus, int n_cu, const int *chroma_qp_tbl /* [2][58] */)
{
    DfCtx c;
    c.prm = prm; c.pic = pic; c.cq[0] = chroma_qp_tbl; c.cq[1] = chroma_qp_tbl + 58;
    c.cod = (uint8_t *)malloc((size_t)pic->w_scu * pic->h_scu);
    for (int pass = 0; pass < 2; pass++) {
        memset(c.cod, 0, (size_t)pic->w_scu * pic->h_scu);
        for (int n = 0; n < n_cu; n++) {
            const int w = 1 << cus[n].log2w, h = 1 << cus[n].log2h;
            c.planes = cus[n].flags & (XB200_CUF_LUMA | XB200_CUF_CHROMA);
            if (pass == 0 && w > 64) { visit(&c, cus[n].x, cus[n].y, w >> 1, h, pass); visit(&c, cus[n].x + 64, cus[n].y, w >> 1, h, pass); }
            else if (pass == 1 && h > 64) { visit(&c, cus[n].x, cus[n].y, w, h >> 1, pass); visit(&c, cus[n].x, cus[n].y + 64, w, h >> 1, pass); }
            else visit(&c, cus[n].x, cus[n].y, w, h, pass);
        }
    }
    free(c.cod);
    return XB200_OK;
}

/* ------------------------------------------------------------------------------------------------------------------------
 * Main-profile deblocking (sps->tool_addb == 1): an H.264-style filter.
 * Restates get_bs (src_main/xevdm_df.c:361-513), deblock_scu_line_luma / _chroma (:584-781) and the CU walkers
 * deblock_addb_cu_hor / _ver (:835-1135) for one tile / one slice, TREE_LC.
 *   - only edges on the 8x8 luma grid are filtered (:853,1054,1109)
 *   - bS: 4 intra and the two SCUs in different CTUs, 3 intra or IBC, 2 luma cbf, else 1 / 0 from comparing the reference
 *     PICTURES (not indices) and the motion vectors
 *   - QP = (qp_cur + qp_nb + 1) >> 1; alpha/beta/c1 from the tables below (src_main/xevdm_tbl.c:377-388)
 *   - quirks kept: get_index takes (u8 qp, u8 offset) so negative values wrap before the clip; beta and c1 are u8
 * ---------------------------------------------------------------------------------------------------------------------- */
static const uint8_t k_alpha[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 17, 20, 22, 25, 28, 32, 36, 40, 45,
                                    50, 56, 63, 71, 80, 90, 101, 113, 127, 144, 162, 182, 203, 226, 255, 255};
static const uint8_t k_beta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10,
                                   11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18};
static const uint8_t k_clip[52][5] = {
    {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0},
    {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0},
    {0, 0, 0, 0, 0}, {0, 0, 0, 1, 1}, {0, 0, 0, 1, 1}, {0, 0, 0, 1, 1}, {0, 0, 0, 1, 1}, {0, 0, 1, 1, 1}, {0, 0, 1, 1, 1}, {0, 1, 1, 1, 1},
    {0, 1, 1, 1, 1}, {0, 1, 1, 1, 1}, {0, 1, 1, 1, 1}, {0, 1, 1, 2, 2}, {0, 1, 1, 2, 2}, {0, 1, 1, 2, 2}, {0, 1, 1, 2, 2}, {0, 1, 2, 3, 3},
    {0, 1, 2, 3, 3}, {0, 2, 2, 3, 3}, {0, 2, 2, 4, 4}, {0, 2, 3, 4, 4}, {0, 2, 3, 4, 4}, {0, 3, 3, 5, 5}, {0, 3, 4, 6, 6}, {0, 3, 4, 6, 6},
    {0, 4, 5, 7, 7}, {0, 4, 5, 8, 8}, {0, 4, 6, 9, 9}, {0, 5, 7, 10, 10}, {0, 6, 8, 11, 11}, {0, 6, 8, 13, 13}, {0, 7, 10, 14, 14}, {0, 8, 11, 16, 16},
    {0, 9, 12, 18, 18}, {0, 10, 13, 20, 20}, {0, 11, 15, 23, 23}, {0, 13, 17, 25, 25}};
const uint8_t *orc_addb_tables(int which) { return which == 0 ? k_alpha : (which == 1 ? k_beta : &k_clip[0][0]); }

typedef struct {
    const XB200_PARAMS *prm;
    ORC_PIC *pic;
    const int *cq[2];
    const int *ref_id[2];        /* identity of the picture behind (list, refi) */
    uint8_t *cod;
    uint8_t *ats;                /* mctx->map_ats_inter: non-zero on the SCUs of ats_inter CUs (tool_ats) */
    int planes;                  /* XB200_CUF_LUMA | XB200_CUF_CHROMA of the CU being visited (tree_cons, xevdm_df.c:916-920,986-997) */
} AddbCtx;

static int addb_index(int qp, int offset) { return orc_clip3(0, 51, (int)(uint8_t)qp + (int)(uint8_t)offset); }
static int near_mv(const int *a, const int *b) { return abs(a[0] - b[0]) < 4 && abs(a[1] - b[1]) < 4; }

static int addb_bs(const AddbCtx *c, int cur, int nb, int x0, int y0, int x1, int y1)
{
    const ORC_PIC *p = c->pic;
    const uint32_t m0 = p->map_scu[cur], m1 = p->map_scu[nb];
    const int lg = c->prm->log2_ctu;
    const int intra = ((m0 >> 15) & 1) || ((m1 >> 15) & 1);
    if (intra && ((x0 >> lg) != (x1 >> lg) || (y0 >> lg) != (y1 >> lg))) return 4;
    if (intra) return 3;
    if (((m0 >> 26) & 1) || ((m1 >> 26) & 1)) return 3;
    if (((m0 >> 24) & 1) || ((m1 >> 24) & 1) || c->ats[cur] || c->ats[nb]) return 2;       /* ats_present, xevdm_df.c:415,902-906 */
    const int8_t *r0 = p->map_refi + 2 * cur, *r1 = p->map_refi + 2 * nb;
    const int16_t *v0 = DF_MV_AT(p, cur), *v1 = DF_MV_AT(p, nb);
    int pa[2], pb[2], a[2][2], b[2][2];
    for (int l = 0; l < 2; l++) {
        pa[l] = r0[l] >= 0 ? c->ref_id[l][r0[l]] : -1;          /* NULL picture */
        pb[l] = r1[l] >= 0 ? c->ref_id[l][r1[l]] : -1;
        for (int d = 0; d < 2; d++) { a[l][d] = r0[l] >= 0 ? v0[l * 2 + d] : 0; b[l][d] = r1[l] >= 0 ? v1[l * 2 + d] : 0; }
    }
    if ((pa[0] == pb[0] && pa[1] == pb[1]) || (pa[0] == pb[1] && pa[1] == pb[0])) {
        if (pa[0] == pa[1]) return (near_mv(a[0], b[0]) && near_mv(a[1], b[1]) && near_mv(a[0], b[1]) && near_mv(a[1], b[0])) ? 0 : 1;
        if (pa[0] == pb[0] && pa[1] == pb[1]) return (near_mv(a[0], b[0]) && near_mv(a[1], b[1])) ? 0 : 1;
        return (near_mv(a[0], b[1]) && near_mv(a[1], b[0])) ? 0 : 1;
    }
    return 1;
}

/* one line across the edge: q[i] = buf[i*s], p[i] = buf[-(i+1)*s] */
static void addb_line_luma(pel *buf, int s, int bs, int alpha, int beta, int c1, int bd)
{
    int p[4], q[4], po[4], qo[4];
    const int maxv = (1 << bd) - 1;
    for (int i = 0; i < 4; i++) { q[i] = buf[i * s]; p[i] = buf[-(i + 1) * s]; po[i] = p[i]; qo[i] = q[i]; }
    if (!(bs && abs(p[0] - q[0]) < alpha && abs(p[1] - p[0]) < beta && abs(q[1] - q[0]) < beta)) return;
    const int ap = abs(p[0] - p[2]) < beta, aq = abs(q[0] - q[2]) < beta;
    if (bs == 4) {
        const int small = abs(p[0] - q[0]) < ((alpha >> 2) + 2);
        if (ap && small) {
            po[0] = (p[2] + 2 * (p[1] + p[0] + q[0]) + q[1] + 4) >> 3;
            po[1] = (p[2] + p[1] + p[0] + q[0] + 2) >> 2;
            po[2] = (2 * p[3] + 3 * p[2] + p[1] + p[0] + q[0] + 4) >> 3;
        } else po[0] = (2 * p[1] + p[0] + q[1] + 2) >> 2;
        if (aq && small) {
            qo[0] = (q[2] + 2 * (q[1] + q[0] + p[0]) + p[1] + 4) >> 3;
            qo[1] = (q[2] + q[1] + q[0] + p[0] + 2) >> 2;
            qo[2] = (2 * q[3] + 3 * q[2] + q[1] + q[0] + p[0] + 4) >> 3;
        } else qo[0] = (2 * q[1] + q[0] + p[1] + 2) >> 2;
    } else {
        const int c0 = (uint8_t)(c1 + ((ap + aq) << orc_max(0, bd - 9)));
        const int d0 = orc_clip3(-c0, c0, (4 * (q[0] - p[0]) + p[1] - q[1] + 4) >> 3);
        po[0] = orc_clip3(0, maxv, p[0] + d0);
        qo[0] = orc_clip3(0, maxv, q[0] - d0);
        if (ap) po[1] = (int16_t)(p[1] + orc_clip3(-c1, c1, (((p[2] + p[0] + q[0]) * 3) - 8 * p[1] - q[1]) >> 4));
        if (aq) qo[1] = (int16_t)(q[1] + orc_clip3(-c1, c1, (((q[2] + q[0] + p[0]) * 3) - 8 * q[1] - p[1]) >> 4));
    }
    for (int i = 0; i < 4; i++) { buf[i * s] = (pel)orc_clip3(0, maxv, qo[i]); buf[-(i + 1) * s] = (pel)orc_clip3(0, maxv, po[i]); }
}
static void addb_line_chroma(pel *buf, int s, int bs, int alpha, int beta, int c0, int bd)
{
    int p[2], q[2], po[2], qo[2];
    const int maxv = (1 << bd) - 1;
    for (int i = 0; i < 2; i++) { q[i] = buf[i * s]; p[i] = buf[-(i + 1) * s]; po[i] = p[i]; qo[i] = q[i]; }
    if (!(bs && abs(p[0] - q[0]) < alpha && abs(p[1] - p[0]) < beta && abs(q[1] - q[0]) < beta)) return;
    if (bs == 4) {
        po[0] = (2 * p[1] + p[0] + q[1] + 2) >> 2;
        qo[0] = (2 * q[1] + q[0] + p[1] + 2) >> 2;
    } else {
        const int d0 = orc_clip3(-c0, c0, (4 * (q[0] - p[0]) + p[1] - q[1] + 4) >> 3);
        po[0] = orc_clip3(0, maxv, p[0] + d0);
        qo[0] = orc_clip3(0, maxv, q[0] - d0);
    }
    for (int i = 0; i < 2; i++) { buf[i * s] = (pel)orc_clip3(0, maxv, qo[i]); buf[-(i + 1) * s] = (pel)orc_clip3(0, maxv, po[i]); }
}

static void addb_segment(AddbCtx *c, int cur, int nb, int x, int y, int vertical)
{
    ORC_PIC *p = c->pic;
    const XB200_PARAMS *prm = c->prm;
    const int bdl = prm->bit_depth_luma, bdc = prm->bit_depth_chroma, scale = bdl - 8;
    const int bs = addb_bs(c, cur, nb, x, y, vertical ? x - 1 : x, vertical ? y : y - 1);
    const int qp = (((p->map_scu[cur] >> 16) & 0x7f) + ((p->map_scu[nb] >> 16) & 0x7f) + 1) >> 1;
    int ia = addb_index(qp, prm->deblock_alpha_offset), ib = addb_index(qp, prm->deblock_beta_offset);
    int alpha = (uint16_t)(k_alpha[ia] << scale), beta = (uint8_t)(k_beta[ib] << scale);
    int c1 = (uint8_t)(k_clip[ia][bs] << orc_max(0, bdl - 9));
    pel *q = p->y + y * p->s_l + x;
    for (int i = 0; i < 4 && (c->planes & XB200_CUF_LUMA); i++)
        vertical ? addb_line_luma(q + i * p->s_l, 1, bs, alpha, beta, c1, bdl) : addb_line_luma(q + i, p->s_l, bs, alpha, beta, c1, bdl);
    for (int k = 0; k < 2 && (c->planes & XB200_CUF_CHROMA); k++) {
        const int qc = orc_clip3(-6 * (bdc - 8), 57, qp + (k ? prm->qp_v_offset : prm->qp_u_offset));
        const int qm = qc < 0 ? qc : c->cq[k][qc];
        ia = addb_index(qm, prm->deblock_alpha_offset); ib = addb_index(qm, prm->deblock_beta_offset);
        alpha = (uint16_t)(k_alpha[ia] << scale); beta = (uint8_t)(k_beta[ib] << scale);
        const int c0 = (uint8_t)((k_clip[ia][bs] + 1) << orc_max(0, bdc - 9));
        pel *qq = (k ? p->v : p->u) + (y >> 1) * p->s_c + (x >> 1);
        for (int i = 0; i < 2; i++) vertical ? addb_line_chroma(qq + i * p->s_c, 1, bs, alpha, beta, c0, bdc) : addb_line_chroma(qq + i, p->s_c, bs, alpha, beta, c0, bdc);
    }
}

static void addb_visit(AddbCtx *c, int x, int y, int w, int h, int pass)
{
    ORC_PIC *p = c->pic;
    const int ws = p->w_scu, sx = x >> 2, sy = y >> 2, nw = w >> 2, nh = h >> 2;
    const int t = sy * ws + sx;
    if (pass == 0) {
        if ((x & 7) == 0 && x > 0 && c->cod[t - 1] && !tile_cut(c->prm, x, 1))
            for (int i = 0; i < nh; i++) addb_segment(c, t + i * ws, t + i * ws - 1, x, y + 4 * i, 1);
        if (((x + w) & 7) == 0 && x + w < p->w_l && c->cod[t + nw] && !tile_cut(c->prm, x + w, 1))
            for (int i = 0; i < nh; i++) addb_segment(c, t + i * ws + nw, t + i * ws + nw - 1, x + w, y + 4 * i, 1);
    } else if ((y & 7) == 0 && y > 0 && !tile_cut(c->prm, y, 0)) {
        for (int i = 0; i < nw; i++) addb_segment(c, t + i, t + i - ws, x + 4 * i, y, 0);
    }
    for (int j = 0; j < nh; j++) memset(c->cod + t + j * ws, 1, nw);
}

int orc_deblock_frame_addb(const XB200_PARAMS *prm, ORC_PIC *pic, const XB200_CU *cus, int n_cu, const int *chroma_qp_tbl,
                           const int *ref_id_l0, const int *ref_id_l1)
{
    AddbCtx c;
    c.prm = prm; c.pic = pic; c.cq[0] = chroma_qp_tbl; c.cq[1] = chroma_qp_tbl + 58; c.ref_id[0] = ref_id_l0; c.ref_id[1] = ref_id_l1;
    c.cod = (uint8_t *)malloc((size_t)pic->w_scu * pic->h_scu);
    c.ats = (uint8_t *)calloc((size_t)pic->w_scu * pic->h_scu, 1);
    if (prm->tool_ats)
        for (int n = 0; n < n_cu; n++)
            if (cus[n].mode != XB200_MODE_INTRA && cus[n].mode != XB200_MODE_IBC && XB200_ATS_INTER_IDX(cus[n].ats))
                for (int j = 0; j < (1 << (cus[n].log2h - 2)); j++)
                    memset(c.ats + ((cus[n].y >> 2) + j) * pic->w_scu + (cus[n].x >> 2), 1, 1 << (cus[n].log2w - 2));
    for (int pass = 0; pass < 2; pass++) {
        memset(c.cod, 0, (size_t)pic->w_scu * pic->h_scu);
        for (int n = 0; n < n_cu; n++) {
            const int w = 1 << cus[n].log2w, h = 1 << cus[n].log2h;
            c.planes = cus[n].flags & (XB200_CUF_LUMA | XB200_CUF_CHROMA);
            if (pass == 0 && w > 64) { addb_visit(&c, cus[n].x, cus[n].y, w >> 1, h, pass); addb_visit(&c, cus[n].x + 64, cus[n].y, w >> 1, h, pass); }
            else if (pass == 1 && h > 64) { addb_visit(&c, cus[n].x, cus[n].y, w, h >> 1, pass); addb_visit(&c, cus[n].x, cus[n].y + 64, w, h >> 1, pass); }
            else addb_visit(&c, cus[n].x, cus[n].y, w, h, pass);
        }
    }
    free(c.cod); free(c.ats);
    return XB200_OK;
}
