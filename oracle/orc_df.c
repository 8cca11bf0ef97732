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
} DfCtx;

static int chroma_map(const DfCtx *c, int k, int q) { return q < 0 ? q : c->cq[k][q]; }

/* one 4-sample segment of an edge between SCU `cur` (right / below) and SCU `nb` (left / above) */
static void edge_segment(DfCtx *c, int cur, int nb, int x, int y, int vertical)
{
    ORC_PIC *p = c->pic;
    const int bdl = c->prm->bit_depth_luma, bdc = c->prm->bit_depth_chroma;
    const int cls = strength_class(p->map_scu[cur], p->map_scu[nb], p->map_refi + 2 * cur, p->map_refi + 2 * nb,
                                   p->map_mv + 4 * cur, p->map_mv + 4 * nb);
    const int qp = (p->map_scu[cur] >> 16) & 0x7f;                       /* QP of the CURRENT side only (T7) */
    const int st = st_lookup(cls, qp) << (bdl - 8);
    if (st) {
        pel *q = p->y + y * p->s_l + x;
        for (int i = 0; i < 4; i++) vertical ? filt_luma(q + i * p->s_l, 1, st, (1 << bdl) - 1) : filt_luma(q + i, p->s_l, st, (1 << bdl) - 1);
    }
    const int qu = orc_clip3(-6 * (bdc - 8), 57, qp + c->prm->qp_u_offset), qv = orc_clip3(-6 * (bdc - 8), 57, qp + c->prm->qp_v_offset);
    const int st_u = st_lookup(cls, chroma_map(c, 0, qu)) << (bdc - 8), st_v = st_lookup(cls, chroma_map(c, 1, qv)) << (bdc - 8);
    for (int k = 0; k < 2; k++) {
        const int s = k ? st_v : st_u;
        if (!s) continue;
        pel *q = (k ? p->v : p->u) + (y >> 1) * p->s_c + (x >> 1);
        for (int i = 0; i < 2; i++) vertical ? filt_chroma(q + i * p->s_c, 1, s, (1 << bdc) - 1) : filt_chroma(q + i, p->s_c, s, (1 << bdc) - 1);
    }
}

static void visit(DfCtx *c, int x, int y, int w, int h, int pass)
{
    ORC_PIC *p = c->pic;
    const int ws = p->w_scu, sx = x >> 2, sy = y >> 2, nw = w >> 2, nh = h >> 2;
    const int t = sy * ws + sx;
    if (pass == 0) {
        if (x > 0 && c->cod[t - 1])
            for (int i = 0; i < nh; i++) edge_segment(c, t + i * ws, t + i * ws - 1, x, y + 4 * i, 1);
        if (x + w < p->w_l && c->cod[t + nw])
            for (int i = 0; i < nh; i++) edge_segment(c, t + i * ws + nw, t + i * ws + nw - 1, x + w, y + 4 * i, 1);
    } else if (y > 0) {
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
            if (pass == 0 && w > 64) { visit(&c, cus[n].x, cus[n].y, w >> 1, h, pass); visit(&c, cus[n].x + 64, cus[n].y, w >> 1, h, pass); }
            else if (pass == 1 && h > 64) { visit(&c, cus[n].x, cus[n].y, w, h >> 1, pass); visit(&c, cus[n].x, cus[n].y + 64, w, h >> 1, pass); }
            else visit(&c, cus[n].x, cus[n].y, w, h, pass);
        }
    }
    free(c.cod);
    return XB200_OK;
}
