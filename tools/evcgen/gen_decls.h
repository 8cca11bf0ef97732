/* force-included when tools/evcgen/Makefile compiles the reference's xevd_eco.c / xevdm_eco.c with the DEFINITIONS of the CABAC
 * decoding primitives renamed away: their uses then bind to the generator's versions (gen_engine.c) */
#include "xevd_def.h"
u32 xevd_sbac_decode_bin(XEVD_BSR *bs, XEVD_SBAC *sbac, SBAC_CTX_MODEL *model);
u32 sbac_decode_bin_ep(XEVD_BSR *bs, XEVD_SBAC *sbac);
u32 xevd_sbac_decode_bin_trm(XEVD_BSR *bs, XEVD_SBAC *sbac);
u32 gen_run(XEVD_BSR *bs, XEVD_SBAC *sbac, SBAC_CTX_MODEL *model, u32 num_ctx, int max_run);
int gen_force_next(int bin);
u32 gen_merge_idx(XEVD_BSR *bs, XEVD_SBAC *sbac, SBAC_CTX_MODEL *model, u32 num_ctx, u32 max_num, u32 limit);
u32 gen_unary_ep(XEVD_BSR *bs, XEVD_SBAC *sbac, u32 max_val, u32 max_symbol);
int gen_last_bound(int pos, int width, int height);
int gen_ibc_pick(XEVD_CTX *ctx, XEVD_CORE *core);
int gen_ibc_mvd(XEVD_BSR *bs, XEVD_SBAC *sbac, s16 mvd[MV_D]);
int gen_tile_start(void);
