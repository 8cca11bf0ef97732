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
#include <stdint.h>
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
    if (g_force >= 0) is_lps = ((uint32_t)g_force != mps);
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

u32 sbac_decode_bin_ep(XEVD_BSR *bs, XEVD_SBAC *sbac)
{
    (void)bs; (void)sbac;
    const uint32_t bin = (rnd32() & 255) < (uint32_t)E.ep_one_q8;
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
    for (size_t k = 0; k < E.n_log; k++) {
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
