/* trace_engine.c -- debugging twin of gen_engine.c: the REAL decoding primitives of the reference, wrapped so that the sequence of
 * (context model, bin) can be compared with what the generator chose (libxevd_trace.so; tools/evcgen/Makefile `make trace`). */
#include <stdint.h>
#include <stdlib.h>
#include "xevd_def.h"
u32 xtrace_real_decode_bin(XEVD_BSR *bs, XEVD_SBAC *sbac, SBAC_CTX_MODEL *model);
u32 xtrace_real_bin_ep(XEVD_BSR *bs, XEVD_SBAC *sbac);
u32 xtrace_real_bin_trm(XEVD_BSR *bs, XEVD_SBAC *sbac);
static int32_t *T; static size_t nT, capT;
static void put(int32_t off, int32_t bin)
{
    if (nT + 2 > capT) { capT = capT ? capT * 2 : 1 << 16; T = (int32_t *)realloc(T, capT * sizeof(int32_t)); }
    T[nT++] = off; T[nT++] = bin;
}
void trace_reset(void) { nT = 0; }
size_t trace_log(int32_t *out, size_t cap) { size_t n = nT < cap ? nT : cap; for (size_t i = 0; i < n; i++) out[i] = T[i]; return nT / 2; }
u32 xevd_sbac_decode_bin(XEVD_BSR *bs, XEVD_SBAC *sbac, SBAC_CTX_MODEL *model)
{
    const int32_t off = (int32_t)((char *)model - (char *)&sbac->ctx);
    const u32 b = xtrace_real_decode_bin(bs, sbac, model);
    put(off, (int32_t)b);
    return b;
}
u32 sbac_decode_bin_ep(XEVD_BSR *bs, XEVD_SBAC *sbac) { const u32 b = xtrace_real_bin_ep(bs, sbac); put(-1, (int32_t)b); return b; }
u32 xevd_sbac_decode_bin_trm(XEVD_BSR *bs, XEVD_SBAC *sbac) { const u32 b = xtrace_real_bin_trm(bs, sbac); put(-1, (int32_t)b); return b; }
