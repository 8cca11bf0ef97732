/*
 * Picture-level reconstruction driver and border padding.  TEST INFRASTRUCTURE ONLY (orc_common.h).
 * Restates the per-CU sequence of xevd_recon_unit (src_base/xevd.c:678-756; Main
 * src_main/xevdm.c:1230-1405) on fully-resolved CU work items: residual (itdq) -> prediction ->
 * rec = clip(pred + resid) (src_base/xevd_recon.c:36-68) -> publish per-SCU maps
 * (xevd_set_dec_info, src_base/xevd_util.c:1574-1650).
 */
#include <string.h>
#include <stdlib.h>
#include "orc_common.h"

/* rec = clip(pred + residual); the residual block is tw x th at (tx, ty) inside the w x h prediction block - the whole block
 * normally, the sub-block TU for ats_inter CUs (xevdm_recon, src_main/xevdm_recon.c:42-126) */
static void put_block_tu(pel *rec, int s_rec, const pel *pred, const int16_t *res, int w, int h, int bd, int tx, int ty, int tw, int th)
{
    const int maxv = (1 << bd) - 1;
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            const int in_tu = res && i >= ty && i < ty + th && j >= tx && j < tx + tw;
            /* t0 is an s16 in the reference (xevd_recon.c:40,60): the sum wraps to 16 bits first */
            int16_t t = (int16_t)((in_tu ? res[(i - ty) * tw + (j - tx)] : 0) + pred[i * w + j]);
            rec[i * s_rec + j] = (pel)orc_clip3(0, maxv, t);
        }
}
static void put_block(pel *rec, int s_rec, const pel *pred, const int16_t *res, int w, int h, int bd)
{
    put_block_tu(rec, s_rec, pred, res, w, h, bd, 0, 0, w, h);
}

/* dmvr_mv: refined vectors per SCU of a DMVR CU (or NULL); aff: extension record of an affine CU (or NULL) */
static void publish_maps(const XB200_PARAMS *prm, ORC_PIC *cur, const XB200_CU *cu, const int16_t *dmvr_mv, const XB200_CU_EXT *aff)
{
    int16_t *amv[2] = { NULL, NULL };
    if (aff)
        for (int l = 0; l < 2; l++)
            if (cu->refi[l] >= 0) {
                amv[l] = (int16_t *)malloc(sizeof(int16_t) * 2 * 32 * 32);
                orc_affine_map_mv(aff->u.affine.cp, cu->refi, cu->log2w, cu->log2h, (cu->flags & XB200_CUF_AFF6) != 0, l, amv[l]);
            }
    const int x0 = cu->x >> 2, y0 = cu->y >> 2, nw = 1 << (cu->log2w - 2), nh = 1 << (cu->log2h - 2);
    const int intra = cu->mode == XB200_MODE_INTRA;
    for (int j = 0; j < nh; j++)
        for (int i = 0; i < nw; i++) {
            const int p = (y0 + j) * cur->w_scu + x0 + i;
            /* MCU_SET_IF_SN_QP | CBFL | SF | COD (xevd_def.h:372-437); slice number 0 */
            uint32_t m = ((uint32_t)(cu->qp_map & 0x7f) << 16) | ((uint32_t)intra << 15) | (1u << 31);
            if (cu->mode == XB200_MODE_IBC) m |= 1u << 26;                /* MCU_SET_IBC (xevdm_def.h:325) */
            int cbfl = cu->cbf & 1;
            if (cbfl && prm->tool_ats && !intra && cu->mode != XB200_MODE_IBC && XB200_ATS_INTER_IDX(cu->ats)) {
                /* xevdm_set_cu_cbf_flags (src_main/xevdm_util.c:3669-3714): luma cbf only on the SCUs of the sub-block TU */
                int tlw, tlh, xo, yo;
                orc_ats_inter_tu(cu->ats, cu->log2w, cu->log2h, &tlw, &tlh, &xo, &yo);
                cbfl = 4 * i >= xo && 4 * i < xo + (1 << tlw) && 4 * j >= yo && 4 * j < yo + (1 << tlh);
            }
            if (cbfl) m |= 1u << 24;
            if (cu->flags & XB200_CUF_SKIP) m |= 1u << 23;
            if (dmvr_mv) m |= 1u << 25;                                   /* MCU_SET_DMVRF (xevdm_def.h:318) */
            if (aff) m |= ((cu->flags & XB200_CUF_AFF6) ? 2u : 1u) << 8;  /* MCU_SET_AFF (xevdm_def.h:336) */
            cur->map_scu[p] = m;
            for (int l = 0; l < 2; l++) {
                cur->map_refi[p * 2 + l] = (intra || cu->mode == XB200_MODE_IBC) ? -1 : cu->refi[l];
                for (int d = 0; d < 2; d++) {
                    /* xevdm_set_dec_info (xevdm_util.c:4313-4340): map_mv gets the refined vectors of a DMVR CU, map_unrefined_mv the
                     * signalled ones (spatial prediction and deblocking read the latter) */
                    /* affine CUs: core->mv as the host passes it (extension record) in map_unrefined_mv and in the lists without a
                     * reference; xevdm_set_affine_mvf's sub-block vectors in map_mv of the lists with one */
                    const int16_t v = intra ? 0 : (aff ? aff->u.affine.mv_unref[l][d] : cu->mv[l][d]);
                    cur->map_mv[(p * 2 + l) * 2 + d] = dmvr_mv ? dmvr_mv[((j * nw + i) * 2 + l) * 2 + d] : (amv[l] ? amv[l][(j * nw + i) * 2 + d] : v);
                    if (cur->map_unrefined_mv) cur->map_unrefined_mv[(p * 2 + l) * 2 + d] = v;
                }
            }
        }
    free(amv[0]); free(amv[1]);
}

int orc_recon_frame(const XB200_PARAMS *prm, ORC_PIC *cur,
                    const ORC_PIC *const *refs_l0, int n_l0, const ORC_PIC *const *refs_l1, int n_l1,
                    const XB200_CU *cus, int n_cu, const XB200_CU_EXT *ext, const int16_t *coef)
{
    pel     *pred = (pel *)malloc(3 * 128 * 128 * sizeof(pel));
    int16_t *res  = (int16_t *)malloc(3 * 128 * 128 * sizeof(int16_t));
    int16_t *dmvr_mv = (int16_t *)malloc(32 * 32 * 4 * sizeof(int16_t));
    (void)ext; (void)n_l0; (void)n_l1;
    for (int n = 0; n < n_cu; n++) {
        const XB200_CU *cu = &cus[n];
        const int w = 1 << cu->log2w, h = 1 << cu->log2h, cw = w >> 1, ch = h >> 1;
        pel *py = pred, *pu = pred + w * h, *pv = pu + cw * ch;
        int16_t *ry = res, *ru = res + w * h, *rv = ru + cw * ch;
        const int16_t *c = coef + cu->coef_off;
        /* local dual tree (src_main/xevdm.c:1828-1846): a TREE_L CU carries luma only, the TREE_C CU that follows its siblings chroma only;
         * every per-plane step of xevd_recon_unit is gated by xevd_check_luma / xevd_check_chroma (:611-640,1344-1391, xevdm_recon.c:135-150).
         * The coefficient stream of such a CU holds the blocks of its own planes only. */
        const int do_l = (cu->flags & XB200_CUF_LUMA) != 0, do_c = (cu->flags & XB200_CUF_CHROMA) != 0;
        const int has_y = do_l && (cu->cbf & 0x00f) != 0, has_u = do_c && (cu->cbf & 0x0f0) != 0, has_v = do_c && (cu->cbf & 0xf00) != 0;
        if ((!do_l && !do_c) || (!(do_l && do_c) && cu->mode != XB200_MODE_INTRA && cu->mode != XB200_MODE_IBC) || (!do_l && cu->mode == XB200_MODE_IBC) ||
            (!do_l && (cu->cbf & 0x00f)) || (!do_c && (cu->cbf & 0xff0))) {      /* cbf bits of planes the CU does not carry must be clear */
            free(pred); free(res); free(dmvr_mv);
            return XB200_ERR_INVALID_ARGUMENT;           /* inter CUs are always TREE_LC, IBC needs luma (xevdm.c:1113-1122) */
        }

        /* plane blocks are padded to multiples of 8 coefficients (include/xevd_b200.h); an ats_inter CU carries only its TU */
        int tlw = cu->log2w, tlh = cu->log2h, txo = 0, tyo = 0;
        if (prm->tool_ats && cu->mode != XB200_MODE_INTRA && cu->mode != XB200_MODE_IBC) orc_ats_inter_tu(cu->ats, cu->log2w, cu->log2h, &tlw, &tlh, &txo, &tyo);
        const int tw = 1 << tlw, th = 1 << tlh;
        if (has_y) { memcpy(ry, c, sizeof(int16_t) * tw * th); c += (tw * th + 7) & ~7; }
        if (has_u) { memcpy(ru, c, sizeof(int16_t) * (tw * th / 4)); c += (tw * th / 4 + 7) & ~7; }
        if (has_v) { memcpy(rv, c, sizeof(int16_t) * (tw * th / 4)); }
        orc_itdq_cu(prm, cu, ry, ru, rv);

        int dmvr = 0;
        if (cu->mode == XB200_MODE_INTER && prm->tool_dmvr && (cu->flags & XB200_CUF_DMVR)) {
            /* xevdm_mc with apply_DMVR (src_main/xevdm_mc.c:1860-2038): both list predictions come from the refinement, then average */
            pel *p1 = (pel *)malloc(sizeof(pel) * (w * h + 2 * cw * ch));
            pel *pp[2][3] = { { py, pu, pv }, { p1, p1 + w * h, p1 + w * h + cw * ch } };
            dmvr = orc_dmvr_pred(prm, cu->x, cu->y, w, h, cu->refi, cu->mv, refs_l0, refs_l1, pp, dmvr_mv);
            if (dmvr) {
                for (int i = 0; i < w * h; i++) py[i] = (pel)((py[i] + pp[1][0][i] + 1) >> 1);
                for (int i = 0; i < cw * ch; i++) { pu[i] = (pel)((pu[i] + pp[1][1][i] + 1) >> 1); pv[i] = (pel)((pv[i] + pp[1][2][i] + 1) >> 1); }
            }
            free(p1);
        }
        const XB200_CU_EXT *aff = NULL;
        if (dmvr) {
        } else if (cu->mode == XB200_MODE_AFFINE) {
            uint32_t ei;
            memcpy(&ei, cu->mv[1], 4);
            aff = &ext[ei];
            orc_affine_pred(prm, cu->x, cu->y, w, h, cu->refi, aff->u.affine.cp, (cu->flags & XB200_CUF_AFF6) != 0, refs_l0, refs_l1, py, pu, pv);
        } else if (cu->mode == XB200_MODE_INTER) {
            orc_inter_pred(prm, cu->x, cu->y, w, h, cu->refi, cu->mv, refs_l0, refs_l1, py, pu, pv);
        } else if (cu->mode == XB200_MODE_IBC) {
            /* xevdm_IBC_mc (src_main/xevdm_mc.c:2040-2106): whole-sample copy from the current picture; chroma vector = luma >> 1 */
            const int bx = cu->mv[0][0], by = cu->mv[0][1];
            for (int i = 0; i < h; i++) memcpy(py + i * w, cur->y + (cu->y + by + i) * cur->s_l + cu->x + bx, sizeof(pel) * w);
            for (int i = 0; i < ch && do_c; i++) {
                memcpy(pu + i * cw, cur->u + ((cu->y >> 1) + (by >> 1) + i) * cur->s_c + (cu->x >> 1) + (bx >> 1), sizeof(pel) * cw);
                memcpy(pv + i * cw, cur->v + ((cu->y >> 1) + (by >> 1) + i) * cur->s_c + (cu->x >> 1) + (bx >> 1), sizeof(pel) * cw);
            }
        } else if (cu->mode == XB200_MODE_INTRA && !prm->tool_eipd) {
            /* xevd_recon_unit intra branch (src_base/xevd.c:732-741): neighbours from the CURRENT picture, so CUs must be
             * reconstructed in decoding order; refi[] carries ipm[0..1], mv[1] the index of the availability masks */
            uint32_t ei;
            pel nb_up[2 * 128 + 2], nb_le[2 * 128 + 2];
            memcpy(&ei, cu->mv[1], 4);
            const XB200_CU_EXT *e = &ext[ei];
            const int ul = (cu->avail >> 2) & 1;
            if (do_l) {
                orc_intra_neighbours(cur->y + cu->y * cur->s_l + cu->x, cur->s_l, w, h, 4, e->u.intra.up, e->u.intra.left, ul,
                                     prm->bit_depth_luma, nb_up + 1, nb_le + 1);
                orc_ipred_base(nb_le + 1, nb_up + 1, py, cu->refi[0], w, h);
            }
            for (int k = 0; k < 2 && do_c; k++) {
                pel *pl = k ? cur->v : cur->u;
                orc_intra_neighbours(pl + (cu->y >> 1) * cur->s_c + (cu->x >> 1), cur->s_c, cw, ch, 2, e->u.intra.up, e->u.intra.left, ul,
                                     prm->bit_depth_luma, nb_up + 1, nb_le + 1);
                orc_ipred_base(nb_le + 1, nb_up + 1, k ? pv : pu, cu->refi[1], cw, ch);
            }
        } else if (cu->mode == XB200_MODE_INTRA) {
            /* Main profile, tool_eipd (src_main/xevdm.c:1344-1361): three reference arrays, 33 luma / 5 chroma modes; chroma is
             * predicted with the CHROMA bit depth, neighbours default to the LUMA depth's mid value (xevdm.c:600-652) */
            uint32_t ei;
            pel nb_up[2 * 128 + 2], nb_le[2 * 128 + 2], nb_ri[2 * 128 + 2];
            memcpy(&ei, cu->mv[1], 4);
            const XB200_CU_EXT *e = &ext[ei];
            const int ul = (cu->avail >> 2) & 1, lr = cu->avail & 3;
            if (do_l) {
                orc_intra_neighbours_main(cur->y + cu->y * cur->s_l + cu->x, cur->s_l, w, h, 4, e->u.intra.up, e->u.intra.left, e->u.intra.right, ul,
                                          prm->bit_depth_luma, nb_up + 1, nb_le + 1, nb_ri + 1);
                orc_ipred_main(nb_le + 1, nb_up + 1, nb_ri + 1, lr, py, cu->refi[0], w, h, prm->bit_depth_luma);
            }
            /* a TREE_C CU takes ipm[0] (the DM mode) from map_ipm at its luma position, IPD_DC when that SCU is not intra
             * (src_main/xevdm.c:1081-1092): host derivation, delivered in refi[0] like any other CU's */
            for (int k = 0; k < 2 && do_c; k++) {
                pel *pl = k ? cur->v : cur->u;
                orc_intra_neighbours_main(pl + (cu->y >> 1) * cur->s_c + (cu->x >> 1), cur->s_c, cw, ch, 2, e->u.intra.up, e->u.intra.left,
                                          e->u.intra.right, ul, prm->bit_depth_luma, nb_up + 1, nb_le + 1, nb_ri + 1);
                orc_ipred_uv_main(nb_le + 1, nb_up + 1, nb_ri + 1, lr, k ? pv : pu, cu->refi[1], cu->refi[0], cw, ch, prm->bit_depth_chroma);
            }
        } else {
            free(pred); free(res); free(dmvr_mv);
            return XB200_ERR_UNSUPPORTED;
        }
        if (do_l) put_block_tu(cur->y + cu->y * cur->s_l + cu->x, cur->s_l, py, has_y ? ry : NULL, w, h, prm->bit_depth_luma, txo, tyo, tw, th);
        /* the reference passes the LUMA bit depth to all three planes (xevd_recon.c:70-91) */
        if (do_c) put_block_tu(cur->u + (cu->y >> 1) * cur->s_c + (cu->x >> 1), cur->s_c, pu, has_u ? ru : NULL, cw, ch, prm->bit_depth_luma, txo >> 1, tyo >> 1, tw >> 1, th >> 1);
        if (do_c) put_block_tu(cur->v + (cu->y >> 1) * cur->s_c + (cu->x >> 1), cur->s_c, pv, has_v ? rv : NULL, cw, ch, prm->bit_depth_luma, txo >> 1, tyo >> 1, tw >> 1, th >> 1);
        /* Main tool_htdf (src_main/xevdm.c:1381-1391): luma post-filter of CUs with a luma residual and of every intra CU, slice QP */
        if (prm->tool_htdf && cu->mode != XB200_MODE_IBC && (has_y || cu->mode == XB200_MODE_INTRA) && (cu->flags & XB200_CUF_LUMA))
            orc_htdf(cur->y + cu->y * cur->s_l + cu->x, cur->s_l, w, h, prm->slice_qp, cu->mode == XB200_MODE_INTRA, cu->avail_cu, prm->bit_depth_luma,
                     cur->map_scu + (cu->y >> 2) * cur->w_scu + (cu->x >> 2), cur->w_scu, cu->mode == XB200_MODE_INTRA && prm->constrained_intra_pred);
        /* xevdm_set_dec_info writes the maps of luma-carrying CUs only (src_main/xevdm_util.c:4241); a TREE_C CU leaves what its luma
         * siblings published (its COD bits are already set) */
        if (do_l) publish_maps(prm, cur, cu, dmvr ? dmvr_mv : NULL, aff);
    }
    free(pred); free(res); free(dmvr_mv);
    return XB200_OK;
}

/* xevd_picbuf_expand -> picbuf_expand (xevd_util.c:365-427): replicate the outermost samples into the
 * pad_l / pad_c wide border, rows first then whole rows up and down */
static void pad_plane(pel *a, int s, int w, int h, int pad)
{
    for (int y = 0; y < h; y++) {
        pel *row = a + y * s;
        for (int x = 1; x <= pad; x++) { row[-x] = row[0]; row[w - 1 + x] = row[w - 1]; }
    }
    for (int y = 1; y <= pad; y++) {
        memcpy(a - y * s - pad, a - pad, sizeof(pel) * (w + 2 * pad));
        memcpy(a + (h - 1 + y) * s - pad, a + (h - 1) * s - pad, sizeof(pel) * (w + 2 * pad));
    }
}

void orc_pad(ORC_PIC *pic)
{
    pad_plane(pic->y, pic->s_l, pic->w_l, pic->h_l, pic->pad_l);
    pad_plane(pic->u, pic->s_c, pic->w_c, pic->h_c, pic->pad_c);
    pad_plane(pic->v, pic->s_c, pic->w_c, pic->h_c, pic->pad_c);
}
