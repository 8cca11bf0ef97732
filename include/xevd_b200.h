/*
 * xevd_b200.h -- C ABI of the B200-native XEVD picture-reconstruction path.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no torch / C++ types.
 * Everything below replaces, at frame granularity, what the reference reaches through its
 * internal function tables after entropy decode has filled XEVD_CU_DATA:
 *
 *   reference seam (file:line)                                  -> entry point here
 *   -----------------------------------------------------------------------------------------
 *   xevd_platform_init / xevdm_platform_init table wiring        -> xb200_create / xb200_destroy
 *     (src_base/xevd.c:2074-2149, src_main/xevdm.c:3388-3479)
 *   PICBUF_ALLOCATOR.fn_alloc / fn_free (xevd_def.h:685-705,      -> xb200_pic_alloc / xb200_pic_free
 *     installed src_base/xevd.c:335-342)                             xb200_pic_upload / xb200_pic_download
 *   xevd_ctu_row_rec_mt -> xevd_recon_tree -> xevd_recon_unit     -> xb200_recon_frame[_dev]
 *     (src_base/xevd.c:1470,1019,678; src_main/xevdm.c:2463,1854,1230):
 *     cu_init + coef_rect_to_series + xevd_sub_block_itdq + xevd_mc /
 *     xevd_ipred + xevd_recon_yuv + xevd_set_dec_info
 *   ctx->fn_deblock = xevd_deblock / xevdm_deblock                -> xb200_deblock
 *     (src_base/xevd.c:1116, src_main/xevdm.c:2048)
 *   mctx->fn_alf = xevd_alf (src_main/xevdm.c:2105)               -> xb200_alf
 *   ctx->fn_picbuf_expand = xevd_picbuf_expand                    -> xb200_pad
 *     (src_base/xevd_util.c:1487)
 *   leaf tables xevd_func_mc_l/c, fn_itxb, xevdm_fn_itx,          -> xb200_mc_blocks / xb200_itdq_blocks
 *     xevd_func_itrans (per-block CPU callbacks; kept only as        (batched micro-benchmark entry points,
 *     batched entry points)                       BASELINE.json config 5)
 *
 * All entry points return XEVD-style codes: >= 0 success, < 0 failure (inc/xevd.h:50-77).
 * There is no CPU fallback: every call fails with XB200_ERR_NO_DEVICE when no CUDA device
 * is usable.
 */
#ifndef XEVD_B200_H
#define XEVD_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XB200_ABI_VERSION 1

/* error codes: same numeric values as inc/xevd.h:50-73 where a counterpart exists */
#define XB200_OK                     0
#define XB200_ERR                   (-1)
#define XB200_ERR_INVALID_ARGUMENT  (-101)
#define XB200_ERR_OUT_OF_MEMORY     (-102)
#define XB200_ERR_UNSUPPORTED       (-104)
#define XB200_ERR_UNEXPECTED        (-105)
#define XB200_ERR_NO_DEVICE         (-401)   /* no CUDA device / driver: there is no CPU path */
#define XB200_ERR_CUDA              (-402)   /* a CUDA runtime call failed; see xb200_last_error */

typedef int16_t xb200_pel;           /* pel == s16 always (src_base/xevd_port.h:51) */

/* prediction modes of a CU work item (values follow xevd_def.h:287-290 for the first three) */
#define XB200_MODE_INTRA   0
#define XB200_MODE_INTER   1         /* MODE_INTER / MODE_SKIP / MODE_DIR once motion is resolved */
#define XB200_MODE_IBC     4         /* intra block copy: mv[0] = block vector in whole samples, refi = {-1,-1}; the vector must
                                        stay inside the current CTU row at or left of the CU's CTU (as conforming streams do) */
#define XB200_MODE_AFFINE  5

/* XB200_CU.flags */
#define XB200_CUF_LUMA     0x01      /* CU carries luma   (tree_cons != TREE_C) */
#define XB200_CUF_CHROMA   0x02      /* CU carries chroma (tree_cons != TREE_L) */
/* Local dual tree (src_main/xevdm.c:1828-1846,1908-1927): the luma-only leaves of a node come first, then ONE chroma-only CU with the
 * node's position and size, all in the decoding order of the list and all intra (leaves may be IBC).  Such a CU's coefficient blocks
 * hold the planes it carries only, the cbf bits of the other planes are 0; a chroma-only CU delivers its DM luma mode (map_ipm at the
 * node's centre, xevdm.c:1081-1092) in refi[0].  Per-SCU maps are published by the luma-carrying CUs (xevdm_util.c:4241).            */
#define XB200_CUF_SKIP     0x04      /* MODE_SKIP: published to map_scu (MCU_SET_SF) */
#define XB200_CUF_DMVR     0x08      /* DMVR enabled for this CU (Main) */
#define XB200_CUF_ATS_INTRA 0x10     /* ats_intra_cu */
#define XB200_CUF_AFF6     0x20      /* 6-parameter affine (vertex_num == 3) */

/*
 * One fully-resolved coding unit: what xevd_recon_unit holds in XEVD_CORE after cu_init and motion
 * derivation (src_base/xevd.c:567-730), flattened.  32 bytes, little endian, no pointers.
 * The producer is the host motion-derivation pass or a synthetic generator.
 */
#define XB200_ATS_INTER_IDX(a)   (((a) >> 2) & 7)
#define XB200_ATS_INTER_POS(a)   (((a) >> 5) & 1)
typedef struct XB200_CU {
    uint16_t x, y;            /* luma position of the CU's top-left sample                       */
    uint8_t  log2w, log2h;    /* 2..7                                                            */
    uint8_t  mode;            /* XB200_MODE_*                                                    */
    uint8_t  flags;           /* XB200_CUF_*                                                     */
    uint8_t  qp_y, qp_u, qp_v;/* XEVD_CU_DATA.qp_y/u/v: already include 6*(bit_depth-8)          */
    uint8_t  qp_map;          /* core->qp as stored in map_scu (deblock QP)                      */
    int8_t   refi[2];         /* inter: reference indices (list0, list1), < 0 = unused
                                 intra: ipm[0] (luma mode), ipm[1] (chroma mode)                 */
    uint16_t cbf;             /* nnz_sub: bits 0-3 luma 64x64 sub-blocks, 4-7 Cb, 8-11 Cr
                                 (bit (j<<1)|i as xevd_eco.c:618-625); for CUs <= 64 only bit 0  */
    int16_t  mv[2][2];        /* inter: final UNCLIPPED motion vectors [list][x,y], quarter-pel
                                 IBC: mv[0] = block vector in whole samples
                                 intra / affine: mv[1] holds a uint32 index into the extension
                                 array (XB200_CU_EXT), mv[0] is unused                           */
    uint8_t  ats;             /* Main, tool_ats.  bits 0-1: ats_intra mode (ats_intra_mode_h << 1 | ats_intra_mode_v,
                                 0 = DST-7, 1 = DCT-8), used when flags has XB200_CUF_ATS_INTRA;
                                 bits 2-4: ats_inter_idx (0 none, 1 vertical half, 2 horizontal half, 3 vertical
                                 quarter, 4 horizontal quarter), bit 5: ats_inter_pos -- i.e. ats_inter_info =
                                 idx | pos << 4 of xevdm_def.h:232-236.  With ats_inter_idx != 0 the CU's three
                                 coefficient blocks hold only the sub-block transform unit (TU-raster, TU size per
                                 xevdm_get_tu_size, xevdm_util.c:3585-3608)                            */
    uint8_t  avail;           /* avail_lr (bits 0-1) | up-left available (bit 2)                 */
    uint16_t avail_cu;        /* Main tool_htdf: the AVAIL_* bits xevd_get_avail_intra yields for this CU when it is
                                 reconstructed (src_base/xevd_util.c:689-745; bit numbers xevd_def.h:237-247: UP 0,
                                 LE 1, RI 3, UP_LE 5, UP_RI 6, LO_LE 7, LO_RI 8); 0 when the tool is off               */
    uint32_t coef_off;        /* offset (in int16 units) of this CU's coefficients inside the
                                 coefficient stream: [Y w*h][Cb w*h/4][Cr w*h/4], CU-raster;
                                 planes whose cbf bits are all 0 are absent (no bytes); every
                                 plane block starts on a multiple of 8 int16 (16 bytes) and is
                                 zero-padded up to one (only 4-wide/4-high CUs ever need pad).
                                 The stream is in decoding order: coef_off of CU k+1 equals
                                 coef_off of CU k plus the (padded) size of CU k's coded planes,
                                 also for CUs without coefficients (size 0)                      */
} XB200_CU;

/* extension record, 32 bytes: meaning depends on XB200_CU.mode */
typedef struct XB200_CU_EXT {
    union {
        struct {              /* XB200_MODE_INTRA: neighbour availability, one bit per SCU */
            uint64_t up;      /* bit i: SCU i of the row above, i in [0, scuw+scuh)  (up then up-right)   */
            uint64_t left;    /* bit i: SCU i of the column left, i in [0, scuh+scuw)                    */
            uint64_t right;   /* Main/SUCO: column right                                                 */
            uint64_t pad;
        } intra;
        struct {              /* XB200_MODE_AFFINE: control point MVs [list][vertex][x,y], 1/4 pel (mcore->affine_mv);   */
            int16_t cp[2][3][2];  /* vertex 2 is used only with XB200_CUF_AFF6                                            */
            int16_t mv_unref[2][2]; /* core->mv[list] when xevdm_set_dec_info runs: published to map_unrefined_mv, and to
                                       map_mv of a list without reference (xevdm_util.c:4313-4340)                      */
        } affine;
    } u;
} XB200_CU_EXT;

/* sequence / picture level switches that change arithmetic */
typedef struct XB200_PARAMS {
    int32_t w, h;                 /* luma picture size in samples (ctx->w, ctx->h)                     */
    int32_t bit_depth_luma;       /* 8..14                                                             */
    int32_t bit_depth_chroma;
    int32_t chroma_format_idc;    /* only 1 (4:2:0) is implemented; others -> XB200_ERR_UNSUPPORTED     */
    int32_t log2_ctu;             /* ctx->log2_max_cuwh: 6 for Baseline, 5..7 Main                      */
    int32_t tool_admvp;           /* 1/16-pel interpolation tables (xevdm_mc.c:121-175)                 */
    int32_t tool_iqt;             /* IQT transform + dq table {..72} (xevdm_itdq.c:423-706)             */
    int32_t tool_ats;
    int32_t tool_addb;
    int32_t tool_alf;
    int32_t tool_htdf;
    int32_t tool_dmvr;
    int32_t tool_eipd;
    int32_t tool_affine;
    int32_t tool_ibc;
    int32_t slice_qp;             /* sh.qp (HTDF, T11)                                                  */
    int32_t qp_u_offset, qp_v_offset;        /* pic_qp_u/v_offset used by deblock chroma QP            */
    int32_t deblock_alpha_offset, deblock_beta_offset;
    int32_t poc;                  /* POC of the current picture                                         */
    int32_t ctu_row0, ctu_rows;   /* band mode (intra-picture multi-GPU sharding): when ctu_rows > 0 only CTU rows
                                     [ctu_row0, ctu_row0 + ctu_rows) are reconstructed; cus / ctu_first / n_ctu describe just
                                     those CTUs (n_ctu = ctu_rows * CTUs per row).  Inter CUs only: intra / IBC / HTDF need the
                                     bands above, which live on other GPUs (bands that are tiles would lift this)            */
    int32_t constrained_intra_pred;   /* pps.constrained_intra_pred_flag: the HTDF ring of an intra CU takes left / right / upper samples
                                         from intra neighbours only (xevdm_recon.c:317,338,359); the intra neighbour masks of
                                         XB200_CU_EXT already carry the same test                                  */
    int32_t tool_suco;                /* sps.sps_suco_flag: CUs of one row may be decoded right-to-left.  Matters to the Baseline deblocking
                                         filter only (tool_addb == 0): chroma edges of 4-wide CUs are 2 samples apart and read what the
                                         neighbouring edge wrote, and the reference filters an edge when the LATER of its two CUs is
                                         visited (src_main/xevdm_df.c:272-300) - with the flag set xb200_recon_frame also publishes the
                                         decoding order per SCU and xb200_deblock walks such runs in that order                       */
    int32_t reserved[6];
} XB200_PARAMS;

typedef struct xb200_ctx xb200_ctx;   /* device context: stream, uploaded tables, scratch              */
typedef struct xb200_pic xb200_pic;   /* device picture: 3 padded planes + per-SCU maps                 */

/* geometry of a device picture; the padded layout follows xevd_imgb_create (xevd_util.c:153-230):
 * pad 144 / 72 on every side, strides in pels                                                          */
typedef struct XB200_PIC_INFO {
    int32_t w_l, h_l, w_c, h_c;
    int32_t s_l, s_c;             /* strides in pels                                                    */
    int32_t pad_l, pad_c;
    void   *dev_y, *dev_u, *dev_v;/* device addresses of sample (0,0) of each plane                     */
    void   *dev_map_mv;           /* int16[w_scu*h_scu][2][2]                                           */
    void   *dev_map_refi;         /* int8 [w_scu*h_scu][2]                                              */
    void   *dev_map_scu;          /* uint32[w_scu*h_scu], bit layout xevd_def.h:372-437                 */
    int32_t w_scu, h_scu;
    int32_t poc;
    void   *dev_map_edge;         /* uint8[w_scu*h_scu], XB200_EDGE_* flags written by xb200_recon_frame      */
    void   *dev_map_unrefined_mv; /* int16[w_scu*h_scu][2][2]: vectors before DMVR refinement (mctx->map_unrefined_mv,
                                     read by spatial MV prediction and by ADDB deblocking); dev_map_mv holds the refined ones
                                     (read by temporal MV prediction, SURVEY T12, and by the Baseline deblocking filter,
                                     src_main/xevdm_df.c:111-124,1143-1166)                                      */
} XB200_PIC_INFO;

/* ---- context ------------------------------------------------------------------------------------- */
int  xb200_abi_version(void);
int  xb200_device_count(void);                            /* < 0 on driver failure                       */
xb200_ctx *xb200_create(int device, int *err);            /* NULL + *err on failure                      */
void xb200_destroy(xb200_ctx *ctx);
const char *xb200_last_error(xb200_ctx *ctx);             /* last CUDA error string (may be "")          */
int  xb200_sync(xb200_ctx *ctx);                          /* wait for the context's stream               */
void *xb200_stream(xb200_ctx *ctx);                       /* the cudaStream_t all launches go to         */
int  xb200_set_stream(xb200_ctx *ctx, void *cuda_stream); /* use a caller-owned stream                   */
long long xb200_launch_count(xb200_ctx *ctx);             /* kernels launched by this context so far     */

/* page-locked host memory: buffers handed to the host-pointer entry points are copied by DMA without an
 * intermediate copy when they come from here (XEVD_CU_DATA.coef in the reference is a plain xevd_malloc,
 * src_base/xevd.c:1685-1738; an integration allocates it with xb200_host_alloc instead) */
void *xb200_host_alloc(size_t bytes);
void  xb200_host_free(void *p);
/* page-lock memory the caller already owns (the XEVD_IMGB planes xevd_picbuf_alloc allocates, src_base/xevd_util.c:153-230), so that
 * xb200_pic_download into it is a true asynchronous DMA; returns XB200_OK or XB200_ERR_CUDA */
int   xb200_host_register(void *p, size_t bytes);
int   xb200_host_unregister(void *p);

/* ---- pictures (PICBUF_ALLOCATOR) ------------------------------------------------------------------ */
xb200_pic *xb200_pic_alloc(xb200_ctx *ctx, int w, int h, int *err);
void xb200_pic_free(xb200_ctx *ctx, xb200_pic *pic);
int  xb200_pic_info(xb200_pic *pic, XB200_PIC_INFO *info);
int  xb200_pic_set_poc(xb200_pic *pic, int poc);
/* host planes are tightly described by (ptr, stride in pels) and hold w x h valid samples;
 * upload copies them into the padded device planes (borders are NOT replicated: call xb200_pad) */
int  xb200_pic_upload(xb200_ctx *ctx, xb200_pic *pic,
                      const xb200_pel *y, int sy, const xb200_pel *u, int su, const xb200_pel *v, int sv);
int  xb200_pic_download(xb200_ctx *ctx, xb200_pic *pic,
                        xb200_pel *y, int sy, xb200_pel *u, int su, xb200_pel *v, int sv);
/* full padded planes (incl. borders), for checking xb200_pad against xevd_picbuf_expand */
int  xb200_pic_download_padded(xb200_ctx *ctx, xb200_pic *pic, xb200_pel *y, xb200_pel *u, xb200_pel *v);
int  xb200_pic_download_maps(xb200_ctx *ctx, xb200_pic *pic, int16_t *map_mv, int8_t *map_refi, uint32_t *map_scu);
int  xb200_pic_download_unrefined_mv(xb200_ctx *ctx, xb200_pic *pic, int16_t *map_unrefined_mv);
int  xb200_pic_download_edge_map(xb200_ctx *ctx, xb200_pic *pic, uint8_t *map_edge);

/* ---- per-picture reconstruction (xevd_ctu_row_rec_mt) ---------------------------------------------- */
/*
 * Reconstruct every CU of one picture.  `cus` are in decoding order; `ctu_first[k]` is the index of the
 * first CU of CTU k (raster CTU order), `ctu_first[n_ctu] == n_cu`.  `refs[l][i]` is the device picture
 * ctx->refp[i][l].pic.  Host variant: all array arguments are HOST pointers; the call stages them
 * through pinned memory and is asynchronous on the context stream.  _dev variant: device pointers.
 */
int  xb200_recon_frame(xb200_ctx *ctx, const XB200_PARAMS *prm, xb200_pic *cur,
                       xb200_pic *const *refs_l0, int n_l0, xb200_pic *const *refs_l1, int n_l1,
                       const XB200_CU *cus, int n_cu, const uint32_t *ctu_first, int n_ctu,
                       const XB200_CU_EXT *ext, int n_ext,
                       const int16_t *coef, size_t n_coef);
/* The same call with the coefficient stream in SPARSE form, for callers that sit behind PCIe: quantised levels are mostly zero, and the
 * dense stream is the larger half of what a picture moves to the device (25.9 MB of a 4K picture's 26.9 MB).  The dense stream of
 * n_coef int16 is cut into chunks of XB200_SPARSE_CHUNK entries; chunk k owns entries[chunk_first[k] .. chunk_first[k + 1]), each
 * (position inside the chunk) | (non-zero level) << 16; chunk_first has ceil(n_coef / XB200_SPARSE_CHUNK) + 1 elements.  The device
 * expands the chunks into the dense stream (one extra kernel) and goes on exactly as xb200_recon_frame; XB200_CU.coef_off and every
 * rule of the dense layout keep their meaning.  This is what a run-length entropy decoder produces before it scatters into a block
 * (xevd_eco_run_length_cc, src_base/xevd_eco.c:354-400: (run, level) pairs along the scan). */
#define XB200_SPARSE_CHUNK 4096
int  xb200_recon_frame_sparse(xb200_ctx *ctx, const XB200_PARAMS *prm, xb200_pic *cur,
                              xb200_pic *const *refs_l0, int n_l0, xb200_pic *const *refs_l1, int n_l1,
                              const XB200_CU *cus, int n_cu, const uint32_t *ctu_first, int n_ctu,
                              const XB200_CU_EXT *ext, int n_ext,
                              const uint32_t *entries, size_t n_entries, const uint32_t *chunk_first, size_t n_coef);
int  xb200_recon_frame_dev(xb200_ctx *ctx, const XB200_PARAMS *prm, xb200_pic *cur,
                       xb200_pic *const *refs_l0, int n_l0, xb200_pic *const *refs_l1, int n_l1,
                       const void *d_cus, int n_cu, const void *d_ctu_first, int n_ctu,
                       const void *d_ext, int n_ext,
                       const void *d_coef, size_t n_coef, int has_intra, int max_cu_per_ctu);
#define XB200_HAS_INTRA      1
#define XB200_HAS_DUAL_TREE  2        /* some CU is luma-only / chroma-only: selects the per-plane owner path (with XB200_HAS_INTRA) */
#define XB200_HAS_DENSE_WAVEFRONT 4   /* most CUs are intra / IBC / HTDF-filtered (an I picture): the CTU wavefront kernel is launched with about as
                                         many persistent CTAs as the wavefront is wide instead of one per CTU, so that the waiting CTAs of
                                         one picture do not fill the device and independent pictures of other contexts can be in flight */
/* has_intra: XB200_HAS_* bits.  XB200_HAS_INTRA when the wavefront pass is needed: the picture has intra or IBC CUs, or tool_htdf is on and some CU has a
 * luma residual.  max_cu_per_ctu: upper bound of ctu_first[k+1]-ctu_first[k] (sizes on-chip work lists); 0 = unknown (worst case) */

/* ---- picture-wide in-loop filters ------------------------------------------------------------------ */
/* edge flags, one byte per SCU */
#define XB200_EDGE_LEFT   0x01    /* a CU/TU boundary runs along the left side of this SCU           */
#define XB200_EDGE_TOP    0x02    /* ... along the top side                                          */
#define XB200_EDGE_LEFT_NOC 0x08  /* the left-side boundary is a luma edge only (inner leaf boundary of a local dual tree node) */
#define XB200_EDGE_TOP_NOC  0x10  /* ... the top-side boundary                                                                   */
#define XB200_EDGE_ATS    0x04    /* the SCU belongs to an ats_inter CU (mctx->map_ats_inter != 0): raises the Main
                                     deblocking strength to "coded" (xevdm_df.c:902-906,977-981)            */
/* Both passes (vertical edges, then horizontal edges) over the whole picture, in place.  The per-SCU maps
 * (map_scu, map_mv, map_refi) and the edge flags are the ones xb200_recon_frame left in `cur`; edge_flags
 * (host, w_scu*h_scu bytes) optionally replaces the device edge map first.  refs_* are only needed by the
 * Main-profile filter (tool_addb), which compares reference PICTURES rather than indices.                  */
int  xb200_deblock(xb200_ctx *ctx, const XB200_PARAMS *prm, xb200_pic *cur,
                   xb200_pic *const *refs_l0, int n_l0, xb200_pic *const *refs_l1, int n_l1,
                   const uint8_t *edge_flags);
/* chroma QP mapping used by deblocking: what xevd_qp_chroma_dynamic[0..1][0..57] holds for the sequence
 * (src_base/xevd_tbl.c:359-425); the context starts with the Baseline default table                         */
int  xb200_set_chroma_qp_table(xb200_ctx *ctx, const int32_t *tbl /* [2][58] */);
/* Tile grid of the pictures that follow (PPS; set_tile_info, src_main/xevdm.c:2162-2327): n_cols + 1 column and n_rows + 1 row
 * boundaries in CTUs (first 0, last = the picture's size in CTUs) and pps.loop_filter_across_tiles_enabled_flag.  Reconstruction
 * needs nothing of this (availability arrives per CU, CTUs in raster order); the loop filters do:
 *   deblocking  - an edge between two tiles is filtered only with the flag set (xevdm_df.c:142,233,877,1088)
 *   ALF         - the 7x7 / 5x5 windows never read another tile: without the flag they mirror at the tile border like at the
 *                 picture border, with it they replicate the tile's border samples and mirror only at the picture's left and top
 *                 (alf_process_tile, xevdm_alf.c:989-1046: tile_boundary_check against the tile or against the picture, windows
 *                 taken from the per-tile extended copy)
 * n_cols == n_rows == 1 (the state of a new context) = one tile; at most 20 columns and 22 rows (xevd_def.h MAX_NUM_TILES_*). */
#define XB200_MAX_TILE_COLS 20
#define XB200_MAX_TILE_ROWS 22
int  xb200_set_tiles(xb200_ctx *ctx, int n_cols, const uint16_t *col_bd, int n_rows, const uint16_t *row_bd, int loop_filter_across_tiles);
/* test support: overwrite the per-SCU maps of a device picture from host arrays (any pointer may be NULL)   */
int  xb200_pic_upload_maps(xb200_ctx *ctx, xb200_pic *pic, const int16_t *map_mv, const int8_t *map_refi,
                           const uint32_t *map_scu, const uint8_t *map_edge);
/* Adaptive loop filter (Main profile, mctx->fn_alf = xevd_alf, src_main/xevdm.c:2105 -> alf_process,
 * src_main/xevdm_alf.c:1167): what reaches the filter after the APS has been parsed and alf_recon_coef
 * (src_main/xevdm_alf.c:700-794) has run on the host.                                                     */
typedef struct XB200_ALF {
    int16_t coef_luma[25][13];    /* alf->coef_final: 25 classes x 13 coefficients of the 7x7 diamond       */
    int16_t coef_chroma[7];       /* alf_slice_param.chroma_coef: 5x5 diamond                               */
    uint8_t enable[3];            /* alf_slice_param.enable_flag[Y, Cb, Cr]                                 */
    uint8_t reserved;
} XB200_ALF;
/* ctb_flag_luma: host array, one byte per CTU (alf_ctb_flag[Y_C]); NULL = every CTU on; luma is filtered
 * where enable[0] and the CTU flag are both set, chroma wherever enable[c] is set.  In place.              */
int  xb200_alf(xb200_ctx *ctx, const XB200_PARAMS *prm, xb200_pic *pic, const XB200_ALF *alf, const uint8_t *ctb_flag_luma);
int  xb200_pad(xb200_ctx *ctx, xb200_pic *pic);           /* xevd_picbuf_expand                        */

/* ---- band exchange (intra-picture sharding across GPUs, BASELINE config 4) ---------------------------------------------------
 * A band = luma rows [y0, y0 + rows) of a picture (y0, rows multiples of the CTU size; the last band may be shorter) with the
 * matching chroma rows and per-SCU map rows.  pack copies it into one contiguous device buffer - the unit of the all-gather
 * that follows reconstruction - and unpack writes a band received from another GPU into the picture.
 * Layout: Y rows (w samples each) | U | V | map_mv | map_unrefined_mv | map_scu | map_refi | map_edge.                        */
size_t xb200_band_bytes(xb200_pic *pic, int rows);
/* Fused alternative (NVLink peer stores, no separate exchange): every rank exports its copy of the picture as a CUDA IPC handle
 * (64 bytes, shipped to the other ranks by the host plumbing) and opens the others' handles; a band-mode xb200_recon_frame on a
 * picture with open peers then writes every reconstructed sample and map entry into all copies directly from the kernel, so the
 * transfer overlaps the arithmetic.  The ranks only need a barrier (any tiny collective on the stream) before using the picture.
 * Pictures must have identical geometry on all ranks.  Available for the Baseline-transform 64x64-CTU kernel; other
 * configurations use xb200_band_pack / all-gather / xb200_band_unpack.                                                      */
int  xb200_pic_export(xb200_ctx *ctx, xb200_pic *pic, void *handle64);
int  xb200_pic_open_peer(xb200_ctx *ctx, xb200_pic *pic, const void *handle64);
int  xb200_band_pack(xb200_ctx *ctx, xb200_pic *pic, int y0, int rows, void *d_dst);
int  xb200_band_unpack(xb200_ctx *ctx, xb200_pic *pic, int y0, int rows, const void *d_src);

/* ---- output path (xevd_pull: SURVEY 8f rows N2 and N4) ------------------------------------------------------------------------
 * Pictures stay in the device-resident DPB; only what xevd_pull hands out crosses PCIe.  xb200_pic_pull applies, in one kernel,
 * the Main-profile dynamic range adjustment (tool_dra: xevd_apply_filter -> xevd_apply_dra_chroma_plane / _luma_plane,
 * src_main/xevdm.c:3311-3348, xevdm_dra.c:272-354; the LUTs are what xevd_init_dra built on the host from the APS), the SPS
 * cropping window (src_main/xevdm.c:3366-3373) and optionally the application's 16 -> 8-bit conversion
 * (app/xevd_app_util.h:359-381), then copies the result to the caller's host planes.  The device picture is not modified
 * (the reference filters a copy too, xevdm.c:3376-3383).                                                                      */
typedef struct XB200_DRA {
    int32_t luma_inv_scale_lut[1024];          /* DRA_CONTROL.luma_inv_scale_lut                                              */
    int32_t chroma_inv_scale_lut[2][1024];     /* DRA_CONTROL.int_chroma_inv_scale_lut                                        */
} XB200_DRA;
/* dra: NULL = no adjustment.  out_bits: 16 (int16 planes) or 8 (uint8 planes, (v + 2) >> 2 clipped).  Strides in samples.
 * Asynchronous on the context stream like xb200_pic_download.                                                                */
int  xb200_pic_pull(xb200_ctx *ctx, xb200_pic *pic, const XB200_DRA *dra, int out_bits, int crop_l, int crop_r, int crop_t, int crop_b,
                    void *y, int sy, void *u, int su, void *v, int sv);

/* ---- batched leaf kernels (micro-benchmarks, BASELINE.json config 5) ------------------------------ */
/* n blocks of (1<<log2w) x (1<<log2h) coefficients, contiguous, in place semantics of
 * xevd_itdq (xevd_itdq.c:494-542): dequant with `scale` then 2-D inverse DCT-2.  iqt selects the Main
 * IQT variant.  d_in / d_out are device pointers.                                                      */
int  xb200_itdq_blocks_dev(xb200_ctx *ctx, const void *d_in, void *d_out, int n, int log2w, int log2h,
                           int qp, int bit_depth, int iqt);
/* n luma (is_chroma=0) or chroma blocks of w x h: d_mv = int32[n][4] {gmv_x, gmv_y (1/16 or 1/32 pel,
 * absolute, already clipped), ori_mv_x, ori_mv_y}; output blocks contiguous w*h each.                  */
int  xb200_mc_blocks_dev(xb200_ctx *ctx, xb200_pic *ref, int plane, const void *d_mv, void *d_out, int n,
                         int w, int h, int bit_depth, int main_tables);

#ifdef __cplusplus
}
#endif
#endif /* XEVD_B200_H */
