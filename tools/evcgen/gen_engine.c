/*
 * gen_engine.c -- turns the reference's own slice-data PARSER into a bitstream GENERATOR (test tooling, not product code).
 *
 * tools/evcgen/Makefile links the unmodified reference decoder with one substitution: the three CABAC decoding primitives of
 * src_base/xevd_eco.c (xevd_sbac_decode_bin :35, sbac_decode_bin_ep :96, xevd_sbac_decode_bin_trm :123; and the private copy of
 * sbac_decode_bin_ep in src_main/xevdm_eco.c:43) are replaced by the functions below.  Instead of decoding a bin from the stream
 * they CHOOSE one (pseudo-randomly, following the probability the context model holds) and ENCODE it with the arithmetic encoder
 * that mirrors the decoder's interval arithmetic.  The reference's syntax code (xevd_eco_cu / xevdm_eco_cu, the split tree, all the
 * tool syntax) therefore walks a random but syntactically valid slice, and the bytes the encoder produced ARE a stream that the
 * unmodified decoder parses into exactly that slice.  The reference ships no streams and there is no encoder on the
 * box; this is how the repository gets real EVC access units to decode end to end.
 *
 * Interval arithmetic restated from the decoder (xevd_eco.c:35-165): 14-bit range (8192..16383 after renormalisation), MPS takes the
 * lower part, LPS width = max(437, state * range >> 9), state update +/- ((x + 16) >> 5); bypass halves the range with floor
 * (an odd range loses one unit); the terminating bin takes one unit off the top.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "xevd_def.h"

static struct {
    uint8_t *out; size_t n, cap;        /* settled bytes (a carry can still ripple into them) */
    uint64_t low; int nb;               /* the live window: nb bits that follow the settled bytes */
    uint32_t range;
    uint64_t rng;                       /* xorshift64* state */
    int lps_scale_q8;                   /* 256 = choose the LPS with the probability the model holds */
    int ep_one_q8;                      /* probability of a 1 for bypass bins, /256 */
    int finished;
    long long n_ctx, n_ep, n_trm;
    uint32_t *log; int32_t *log_off; size_t n_log, cap_log;     /* every bin: kind << 30 | model-before << 1 | bin, for the self-check */
} E;

/* tiles: every tile of a slice is its own arithmetic code word (xevd_tile_eco resets the decoder, src_main/xevdm.c:2396); the slice data is
 * their concatenation and the slice header carries the byte sizes as entry points (xevdm_eco.c:2789-2795) */
#define GEN_MAX_TILES 512
static size_t g_tile_off[GEN_MAX_TILES + 1], g_tile_log[GEN_MAX_TILES + 1];
static int g_n_tiles;

static const uint8_t *g_replay; static size_t g_replay_n, g_replay_pos; static int g_replay_err;
/* replay: the next slice takes its bins from a recorded sequence instead of drawing them (used to bring a fresh decoder instance to
 * the state after the pictures accepted so far, before another attempt at the next picture) */
void gen_replay(const uint8_t *bins, size_t n) { g_replay = bins; g_replay_n = n; g_replay_pos = 0; g_replay_err = 0; }
int gen_replay_done(void) { int ok = g_replay && !g_replay_err && g_replay_pos == g_replay_n; g_replay = 0; return ok; }
static int replay_bin(void) { if (g_replay_pos >= g_replay_n) { g_replay_err = 1; return 0; } return g_replay[g_replay_pos++]; }
static int g_force_ep = -1;
static int g_force = -1;             /* -1 free choice, else the bin the next context-coded call must take (gen_run) */
static int32_t g_off;                 /* which context model (offset inside XEVD_SBAC_CTX): lets a trace of the real decoder be compared */
static void log_bin(uint32_t kind, uint32_t model, uint32_t bin)
{
    if (E.n_log == E.cap_log) {
        E.cap_log = E.cap_log ? E.cap_log * 2 : 1 << 16;
        E.log = (uint32_t *)realloc(E.log, E.cap_log * sizeof(uint32_t));
        E.log_off = (int32_t *)realloc(E.log_off, E.cap_log * sizeof(int32_t));
    }
    E.log_off[E.n_log] = kind == 0 ? g_off : -1;
    E.log[E.n_log++] = (kind << 30) | (model << 1) | bin;
}
/* the bin sequence of the slice generated last: out[2k] = context offset (-1 bypass / terminate), out[2k+1] = bin; returns the count */
size_t gen_log(int32_t *out, size_t cap)
{
    size_t n = E.n_log < cap / 2 ? E.n_log : cap / 2;
    for (size_t k = 0; k < n; k++) { out[2 * k] = E.log_off[k]; out[2 * k + 1] = (int32_t)(E.log[k] & 1); }
    return E.n_log;
}

static uint32_t rnd32(void)
{
    E.rng ^= E.rng >> 12; E.rng ^= E.rng << 25; E.rng ^= E.rng >> 27;
    return (uint32_t)((E.rng * 0x2545F4914F6CDD1DULL) >> 32);
}
static void put_byte(uint8_t b)
{
    if (E.n == E.cap) { E.cap = E.cap ? E.cap * 2 : 1 << 16; E.out = (uint8_t *)realloc(E.out, E.cap); }
    E.out[E.n++] = b;
}
static void add_low(uint64_t x)
{
    E.low += x;
    if (E.low >> E.nb) {                /* carry out of the window: ripple into the settled bytes */
        E.low &= (1ULL << E.nb) - 1;
        size_t i = E.n;
        while (i > 0 && E.out[i - 1] == 0xFF) E.out[--i] = 0;
        if (i > 0) E.out[i - 1]++;
    }
}
static void shift1(void)
{
    E.low <<= 1; E.nb++;
    if (E.nb >= 48) { put_byte((uint8_t)(E.low >> (E.nb - 8))); E.nb -= 8; E.low &= (1ULL << E.nb) - 1; }
}

void gen_reset(uint64_t seed, int lps_scale_q8, int ep_one_q8)
{
    E.n = 0; E.low = 0; E.nb = 14; E.range = 16384; E.finished = 0;     /* xevd_eco_sbac_reset: range 16384, value = first 14 bits */
    E.rng = seed * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
    if (!E.rng) E.rng = 1;
    E.lps_scale_q8 = lps_scale_q8 > 0 ? lps_scale_q8 : 256;
    E.ep_one_q8 = ep_one_q8 > 0 ? ep_one_q8 : 128;
    E.n_ctx = E.n_ep = E.n_trm = 0;
    E.n_log = 0;
    g_n_tiles = 0;
}
/* called where xevd_tile_eco resets the arithmetic decoder (tools/evcgen/Makefile): the first tile continues the state gen_reset left,
 * every further tile starts a new code word behind the bytes of the previous one */
int gen_tile_start(void)
{
    if (g_n_tiles > 0) { E.low = 0; E.nb = 14; E.range = 16384; E.finished = 0; }
    if (g_n_tiles < GEN_MAX_TILES) { g_tile_off[g_n_tiles] = E.n; g_tile_log[g_n_tiles] = E.n_log; g_n_tiles++; }
    return 0;
}
/* byte sizes of the tiles of the slice generated last; returns their number */
int gen_tile_sizes(int32_t *out, int cap)
{
    for (int i = 0; i < g_n_tiles && i < cap; i++) out[i] = (int32_t)((i + 1 < g_n_tiles ? g_tile_off[i + 1] : E.n) - g_tile_off[i]);
    return g_n_tiles;
}
/* the slice data produced since gen_reset: whole bytes, the terminating bin included (0 when no slice was finished) */
size_t gen_take(uint8_t *dst, size_t cap)
{
    if (!E.finished || E.n > cap) return 0;
    memcpy(dst, E.out, E.n);
    return E.n;
}
void gen_stats(long long *s) { s[0] = E.n_ctx; s[1] = E.n_ep; s[2] = E.n_trm; }

u32 xevd_sbac_decode_bin(XEVD_BSR *bs, XEVD_SBAC *sbac, SBAC_CTX_MODEL *model)
{
    (void)bs;
    g_off = (int32_t)((char *)model - (char *)&sbac->ctx);
    uint32_t state = (*model) >> 1, mps = (*model) & 1;
    uint32_t lps = (state * E.range) >> 9;
    lps = lps < 437 ? 437 : lps;
    /* choice: LPS with probability state / 512 (scaled) */
    uint32_t p = (state * (uint32_t)E.lps_scale_q8) >> 8;
    if (p > 480) p = 480;
    int is_lps = (rnd32() & 511) < p;
    if (g_force >= 0) { is_lps = ((uint32_t)g_force != mps); g_force = -1; }      /* a forced value applies to one bin */
    if (g_replay) is_lps = ((uint32_t)replay_bin() != mps);
    log_bin(0, *model, is_lps ? 1 - mps : mps);
    E.range -= lps;
    uint32_t bin;
    if (is_lps) {
        bin = 1 - mps;
        add_low(E.range);
        E.range = lps;
        state = state + ((512 - state + 16) >> 5);
        if (state > 256) { mps = 1 - mps; state = 512 - state; }
    } else {
        bin = mps;
        state = state - ((state + 16) >> 5);
    }
    *model = (SBAC_CTX_MODEL)((state << 1) + mps);
    while (E.range < 8192) { E.range <<= 1; shift1(); }
    E.n_ctx++;
    return bin;
}

/* sbac_read_unary_sym (xevd_eco.c:207) for the zero run of the run-length coefficient code, with the one constraint the syntax
 * leaves to the encoder: the run must end inside the block.  g_force: -1 free choice, else the bin the next context-coded call must take. */
/* the next context-coded bin must take this value (syntax the decoder parses but a conforming encoder has no choice about,
 * e.g. the split flag of a Baseline coding block that crosses the picture boundary, src_main/xevdm.c:1713) */
int gen_force_next(int bin) { g_force = bin; return 0; }

/* ---- intra block copy -------------------------------------------------------------------------------------------------------
 * The block vector is coded like a motion vector difference (xevdm_eco.c:1796) and nothing in the syntax keeps it inside the part of the
 * picture that is already decoded; the decoder copies from wherever it points (xevdm_IBC_mc).  An encoder only emits vectors whose source
 * block is decoded, inside the picture, and not right of / below the current CTU.  The generator therefore CHOOSES the vector: before the
 * ibc_flag bin it looks for such a vector (gen_ibc_pick; none found -> the flag is forced to 0), and the vector is then written through
 * the decoder's own mvd syntax with every bin forced (gen_ibc_mvd). */
u32 sbac_decode_bin_ep(XEVD_BSR *bs, XEVD_SBAC *sbac);
static int g_ibc_bv[2];
int gen_ibc_pick(XEVD_CTX *ctx, XEVD_CORE *core)
{
    const int x = core->x_scu << MIN_CU_LOG2, y = core->y_scu << MIN_CU_LOG2, w = 1 << core->log2_cuw, h = 1 << core->log2_cuh;
    const int ctu = 1 << ctx->log2_max_cuwh, cx = x & ~(ctu - 1), cy = y & ~(ctu - 1);
    for (int t = 0; t < 24; t++) {
        const uint32_t r = rnd32();
        int dx, dy;
        switch (r & 3) {
        case 0: dx = -(w + (int)((r >> 2) % 48)); dy = (int)((r >> 10) % 17) - 8; break;
        case 1: dy = -(h + (int)((r >> 2) % 40)); dx = (int)((r >> 10) % 33) - 16; break;
        case 2: dx = -(w + (int)((r >> 2) % 24)); dy = -(int)((r >> 10) % 24); break;
        default: dx = -(int)((r >> 2) % 130); dy = -(h + (int)((r >> 12) % 8)); break;
        }
        const int rx = x + dx, ry = y + dy;
        if (rx < 0 || ry < 0 || rx + w > ctx->w || ry + h > ctx->h || rx + w > cx + ctu || ry + h > cy + ctu) continue;
        int ok = 1;
        for (int sy = ry >> MIN_CU_LOG2; sy <= (ry + h - 1) >> MIN_CU_LOG2 && ok; sy++)
            for (int sx = rx >> MIN_CU_LOG2; sx <= (rx + w - 1) >> MIN_CU_LOG2; sx++)
                if (!ctx->cod_eco[sy * ctx->w_scu + sx]) { ok = 0; break; }
        if (!ok) continue;
        g_ibc_bv[0] = dx; g_ibc_bv[1] = dy;
        return 1;
    }
    g_force = 0;                          /* nowhere to copy from: this CU is not an IBC CU */
    return 0;
}
/* xevd_eco_abs_mvd (src_base/xevd_eco.c:520-555) with every bin forced so that it reads `a` */
static u32 ibc_abs(XEVD_BSR *bs, XEVD_SBAC *sbac, SBAC_CTX_MODEL *model, int a)
{
    u32 val = 0, code, len;
    int tl = 0;
    while (((1 << (tl + 1)) - 1) <= a) tl++;
    g_force = a == 0 ? 1 : 0;
    code = xevd_sbac_decode_bin(bs, sbac, model);
    if (code == 0) {
        len = 0;
        while (!(code & 1)) {
            const int want = ((int)len + 1 == tl) ? 1 : 0;
            if (len == 0) { g_force = want; code = xevd_sbac_decode_bin(bs, sbac, model); }
            else { g_force_ep = want; code = sbac_decode_bin_ep(bs, sbac); }
            len++;
        }
        val = (1u << len) - 1;
        const int suffix = a - ((1 << tl) - 1);
        while (len != 0) {
            g_force_ep = (suffix >> (len - 1)) & 1;
            code = sbac_decode_bin_ep(bs, sbac);
            val += code << (--len);
        }
    }
    return val;
}
int gen_ibc_mvd(XEVD_BSR *bs, XEVD_SBAC *sbac, s16 mvd[MV_D])
{
    for (int c = 0; c < 2; c++) {
        const int t = g_ibc_bv[c];
        const s16 v = (s16)ibc_abs(bs, sbac, sbac->ctx.mvd, t < 0 ? -t : t);
        if (v == 0) mvd[c] = 0;
        else {
            g_force_ep = t < 0 ? 1 : 0;
            mvd[c] = sbac_decode_bin_ep(bs, sbac) ? -v : v;
        }
    }
    return XEVD_OK;
}

/* ADCC last-position prefixes in a block with a 64 dimension: stop at 7 (see tools/evcgen/Makefile, LAST_BOUND) */
int gen_last_bound(int pos, int width, int height) { if ((width == 64 || height == 64) && pos >= 7) g_force = 0; return 0; }

u32 gen_run(XEVD_BSR *bs, XEVD_SBAC *sbac, SBAC_CTX_MODEL *model, u32 num_ctx, int max_run)
{
    u32 ctx_idx = 0, symbol = 0, t;
    if (max_run <= 0) g_force = 0;
    symbol = xevd_sbac_decode_bin(bs, sbac, model);
    g_force = -1;
    if (symbol == 0) return 0;
    symbol = 0;
    do {
        if (ctx_idx < num_ctx - 1) ctx_idx++;
        symbol++;
        if ((int)symbol >= max_run) g_force = 0;
        t = xevd_sbac_decode_bin(bs, sbac, &model[ctx_idx]);
        g_force = -1;
    } while (t);
    return symbol;
}

/* sbac_read_truncate_unary_sym (xevdm_eco.c:113) with the value kept below `limit` (<= max_num): the merge index of a small block */
u32 gen_merge_idx(XEVD_BSR *bs, XEVD_SBAC *sbac, SBAC_CTX_MODEL *model, u32 num_ctx, u32 max_num, u32 limit)
{
    u32 ctx_idx = 0;
    if (max_num > 1)
        for (; ctx_idx < max_num - 1; ++ctx_idx) {
            if (ctx_idx + 1 >= limit) g_force = 0;
            const u32 symbol = xevd_sbac_decode_bin(bs, sbac, model + (ctx_idx > num_ctx - 1 ? num_ctx - 1 : ctx_idx));
            g_force = -1;
            if (symbol == 0) break;
        }
    return ctx_idx;
}

u32 sbac_decode_bin_ep(XEVD_BSR *bs, XEVD_SBAC *sbac);
/* sbac_read_unary_sym_ep (xevd_eco.c:166) with the value kept at or below max_symbol */
u32 gen_unary_ep(XEVD_BSR *bs, XEVD_SBAC *sbac, u32 max_val, u32 max_symbol)
{
    u32 t, symbol;
    int counter = 0;
    if (max_symbol == 0) g_force_ep = 0;
    symbol = sbac_decode_bin_ep(bs, sbac); counter++;
    if (symbol == 0) return 0;
    symbol = 0;
    do {
        if (counter == (int)max_val) t = 0;
        else { if (symbol + 1 >= max_symbol) g_force_ep = 0; t = sbac_decode_bin_ep(bs, sbac); }
        counter++; symbol++;
    } while (t);
    return symbol;
}

u32 sbac_decode_bin_ep(XEVD_BSR *bs, XEVD_SBAC *sbac)
{
    (void)bs; (void)sbac;
    uint32_t bin = (rnd32() & 255) < (uint32_t)E.ep_one_q8;
    if (g_force_ep >= 0) { bin = (uint32_t)g_force_ep; g_force_ep = -1; }
    if (g_replay) bin = (uint32_t)replay_bin();
    const uint32_t r2 = E.range >> 1;
    log_bin(1, 0, bin);
    shift1();
    if (bin) add_low(2ULL * r2);
    E.range = r2 << 1;
    E.n_ep++;
    return bin;
}

/* only decoded at the end of a tile (xevd_eco_tile_end_flag): always 1, and the arithmetic code word ends here */
u32 xevd_sbac_decode_bin_trm(XEVD_BSR *bs, XEVD_SBAC *sbac)
{
    (void)sbac;
    {   /* leave the bit reader where the real primitive leaves it (byte aligned, xevd_eco.c:137-145): xevd_eco_cabac_zero_word, which
           runs next, then walks whole 16-bit words of the zero padding the generator appended and stops exactly at the buffer end */
        u32 t0;
        while (!XEVD_BSR_IS_BYTE_ALIGN(bs)) xevd_bsr_read1(bs, &t0);
    }
    if (g_replay) (void)replay_bin();
    log_bin(2, 0, 1);
    E.range--;
    add_low(E.range);
    /* the decoder has consumed exactly the window; what follows up to the byte boundary must be zero bits */
    while (E.nb >= 8) { put_byte((uint8_t)(E.low >> (E.nb - 8))); E.nb -= 8; E.low &= (1ULL << E.nb) - 1; }
    if (E.nb > 0) { put_byte((uint8_t)(E.low << (8 - E.nb))); E.nb = 0; E.low = 0; }
    E.finished = 1;
    E.n_trm++;
    return 1;
}

/* Self-check of the arithmetic encoder: decode the produced bytes with the decoder's arithmetic (xevd_eco.c:35-165) along the logged
 * bin sequence.  Returns -1 when every bin comes back as chosen, else the index of the first bin that does not. */
long long gen_selfcheck(void)
{
    size_t pos = 0;                 /* bit position */
    uint32_t range = 16384, value = 0;
#define NEXTBIT() ((pos >> 3) < E.n ? (uint32_t)((E.out[pos >> 3] >> (7 - (pos & 7))) & 1) : 0u); pos++
    for (int i = 0; i < 14; i++) { uint32_t b = NEXTBIT(); value = ((value << 1) | b) & 0xFFFF; }
    int tile = 1;
    for (size_t k = 0; k < E.n_log; k++) {
        if (tile < g_n_tiles && k == g_tile_log[tile]) {        /* the next tile's code word starts at its byte offset */
            pos = g_tile_off[tile] * 8; range = 16384; value = 0; tile++;
            for (int i = 0; i < 14; i++) { uint32_t b = NEXTBIT(); value = ((value << 1) | b) & 0xFFFF; }
        }
        const uint32_t kind = E.log[k] >> 30, want = E.log[k] & 1;
        uint32_t bin;
        if (kind == 0) {
            const uint32_t model = (E.log[k] >> 1) & 0xFFFF, state = model >> 1, mps = model & 1;
            uint32_t lps = (state * range) >> 9;
            lps = lps < 437 ? 437 : lps;
            range -= lps;
            if (value >= range) { bin = 1 - mps; value -= range; range = lps; } else bin = mps;
            while (range < 8192) { range <<= 1; uint32_t b = NEXTBIT(); value = ((value << 1) | b) & 0xFFFF; }
        } else if (kind == 1) {
            range >>= 1;
            if (value >= range) { bin = 1; value -= range; } else bin = 0;
            range <<= 1;
            uint32_t b = NEXTBIT(); value = ((value << 1) | b) & 0xFFFF;
        } else {
            range--;
            bin = value >= range;
        }
        if (bin != want) return (long long)k;
    }
    return -1;
}

/* ---- conformance watch ------------------------------------------------------------------------------------------------------
 * Random syntax can describe a coding unit no encoder would emit although every element parses: an inter CU whose derived motion
 * has no valid reference picture in either list (a merge / MMVD index pointing at a candidate that does not exist; the reference then
 * "predicts" from whatever its scratch buffer holds).  tools/evcgen/Makefile routes the generator's xevdm_mc / xevdm_affine_mc calls
 * through these wrappers; a slice that trips one is reported by gen_invalid() and the stream is discarded. */
#include "xevdm_def.h"
#include "xevdm_ipred.h"
static int g_invalid;
static char g_why[160];
int gen_invalid(char *why, int cap) { if (why && cap > 0) { strncpy(why, g_why, (size_t)cap - 1); why[cap - 1] = 0; } int v = g_invalid; g_invalid = 0; return v; }
typedef void (*MC_FN)(int, int, int, int, int, int, s8 *, s16 (*)[MV_D], XEVD_REFP (*)[REFP_NUM], void *, int, pel *, void *, void *, BOOL, void *, u8 *, void *,
                      int, int, int, int);
typedef void (*AFF_FN)(int, int, int, int, int, int, s8 *, void *, XEVD_REFP (*)[REFP_NUM], void *, int, pel *, int, int, int);
static void watch(const char *what, int x, int y, int w, int h, s8 refi[REFP_NUM], XEVD_REFP (*refp)[REFP_NUM], void *pred)
{
    XEVD_CORE *core = (XEVD_CORE *)((char *)pred - offsetof(XEVD_CORE, pred));
    XEVDM_CORE *m = (XEVDM_CORE *)core;
    int bad = !REFI_IS_VALID(refi[REFP_0]) && !REFI_IS_VALID(refi[REFP_1]);
    for (int l = 0; l < REFP_NUM && !bad; l++)
        if (REFI_IS_VALID(refi[l]) && (refi[l] >= XEVD_MAX_NUM_REF_PICS || !refp[refi[l]][l].pic)) bad = 1;
    if (bad && !g_invalid) {
        g_invalid = 1;
        snprintf(g_why, sizeof(g_why), "%s CU %dx%d at (%d,%d): refi %d/%d, pred_mode %d, inter_dir %d, mmvd %d, affine %d, mvp_idx %d/%d", what, w, h, x, y,
                 refi[0], refi[1], core->pred_mode, core->inter_dir, m->mmvd_flag, m->affine_flag, core->mvp_idx[0], core->mvp_idx[1]);
    }
}
void gen_watch_mc(int x, int y, int pic_w, int pic_h, int w, int h, s8 refi[REFP_NUM], s16 (*mv)[MV_D], XEVD_REFP (*refp)[REFP_NUM],
                  void *pred, int poc_c, pel *dmvr_current_template, void *a, void *b, BOOL apply_DMVR, void *c, u8 *cu_dmvr_flag,
                  void *d, int sps_admvp_flag, int bit_depth_luma, int bit_depth_chroma, int chroma_format_idc)
{
    s8 r0[REFP_NUM] = {0, -1};
    watch("inter", x, y, w, h, refi, refp, pred);
    if (g_invalid) { if (!refp[0][0].pic) return; refi = r0; }       /* keep the generator itself on defined ground */
    ((MC_FN)xevdm_mc)(x, y, pic_w, pic_h, w, h, refi, mv, refp, pred, poc_c, dmvr_current_template, a, b, apply_DMVR, c, cu_dmvr_flag, d, sps_admvp_flag,
                      bit_depth_luma, bit_depth_chroma, chroma_format_idc);
}
void gen_watch_affine_mc(int x, int y, int pic_w, int pic_h, int w, int h, s8 refi[REFP_NUM], void *mv, XEVD_REFP (*refp)[REFP_NUM],
                         void *pred, int vertex_num, pel *tmp_buffer, int bit_depth_luma, int bit_depth_chroma, int chroma_format_idc)
{
    watch("affine", x, y, w, h, refi, refp, pred);
    if (g_invalid) return;
    ((AFF_FN)xevdm_affine_mc)(x, y, pic_w, pic_h, w, h, refi, mv, refp, pred, vertex_num, tmp_buffer, bit_depth_luma, bit_depth_chroma, chroma_format_idc);
}

/* intra modes outside the defined sets (chroma: IPD_CHROMA_CNT, luma: IPD_CNT) */
typedef void (*IPUV_FN)(pel *, pel *, pel *, u16, pel *, int, int, int, int, int);
typedef void (*IP_FN)(pel *, pel *, pel *, u16, pel *, int, int, int, int);
void gen_watch_ipred_uv(pel *src_le, pel *src_up, pel *src_ri, u16 avail_lr, pel *dst, int ipm_c, int ipm, int w, int h, int bit_depth)
{
    if ((ipm_c < 0 || ipm_c >= IPD_CHROMA_CNT || ipm < 0 || ipm >= IPD_CNT) && !g_invalid) {
        g_invalid = 1;
        snprintf(g_why, sizeof(g_why), "intra chroma mode %d (luma %d) in a %dx%d block", ipm_c, ipm, w, h);
    }
    ((IPUV_FN)xevdm_ipred_uv)(src_le, src_up, src_ri, avail_lr, dst, ipm_c, ipm, w, h, bit_depth);
}
void gen_watch_ipred(pel *src_le, pel *src_up, pel *src_ri, u16 avail_lr, pel *dst, int ipm, int w, int h, int bit_depth)
{
    if ((ipm < 0 || ipm >= IPD_CNT) && !g_invalid) { g_invalid = 1; snprintf(g_why, sizeof(g_why), "intra luma mode %d in a %dx%d block", ipm, w, h); }
    ((IP_FN)xevdm_ipred)(src_le, src_up, src_ri, avail_lr, dst, ipm, w, h, bit_depth);
}
