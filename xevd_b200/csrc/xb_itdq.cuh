// xb_itdq.cuh -- inverse quantisation and inverse DCT-2 device code.
//
// Replaces xevd_dquant + xevd_itx_pb{2..64}b (src_base/xevd_itdq.c:48-517, dispatched AVX2 variant
// src_base/avx/xevd_itdq_avx.c:2486) and the Main IQT xevdm_itx_pb{2..64} (src_main/xevdm_itdq.c:423-724).
//
// Each thread owns one line (a column in pass 1, a row in pass 2) of a transform block and evaluates
// the N-point inverse DCT-2 in registers with the even/odd factorisation
//      x[n], x[N-1-n] = E[n] +- O[n],  E = IDCT_{N/2}(even inputs),  O[n] = sum_{k odd} T[k][n] X[k]
// with the kernel entries as compile-time literals.  All sums are plain 32-bit integers: this
// reproduces the dispatched x86 path of the reference bit for bit (mod-2^32 accumulation in pass 2,
// SURVEY T4) and the plain-C path whenever nothing overflows.
#pragma once
#include "xb_common.cuh"

namespace xb {

__device__ constexpr int8_t kTM64[64][64] = {
#include "gen/dct2_tm64.inc"
};

// entry [k][n] of the N-point kernel (xevd_tbl_tmN), n < N
template <int N> __device__ __forceinline__ constexpr int tm(int k, int n) { return kTM64[k * (64 / N)][n]; }

template <int N> struct InvDct2 {
    // out[n] = sum_k tm<N>(k, n) * in[k]
    static __device__ __forceinline__ void run(const int (&in)[N], int (&out)[N])
    {
        int ev[N / 2], E[N / 2];
#pragma unroll
        for (int k = 0; k < N / 2; k++) ev[k] = in[2 * k];
        InvDct2<N / 2>::run(ev, E);
#pragma unroll
        for (int n = 0; n < N / 2; n++) {
            int o = 0;
#pragma unroll
            for (int k = 1; k < N; k += 2) o += tm<N>(k, n) * in[k];
            out[n] = E[n] + o;
            out[N - 1 - n] = E[n] - o;
        }
    }
};
template <> struct InvDct2<2> {
    static __device__ __forceinline__ void run(const int (&in)[2], int (&out)[2])
    {
        out[0] = 64 * (in[0] + in[1]);
        out[1] = 64 * (in[0] - in[1]);
    }
};

// xevd_itdq prologue (xevd_itdq.c:494-517): the dequantiser of one transform block
struct Dequant {
    long long mul;      // scale * (181 when log2w + log2h is odd)
    long long off;
    int shift;
    __device__ __forceinline__ void init(int log2w, int log2h, int qp, int bit_depth, int iqt)
    {
        const int odd = (log2w + log2h) & 1;
        shift = 20 - 14 - (15 - bit_depth - ((log2w + log2h) >> 1)) + (odd ? 8 : 0);
        off = shift == 0 ? 0 : (1LL << (shift - 1));
        mul = (long long)(c_dq_scale[iqt][qp % 6] << (qp / 6)) * (odd ? 181 : 1);
    }
    __device__ __forceinline__ int apply(int c) const
    {
        if (c == 0) return 0;
        long long v = (c * mul + off) >> shift;
        return (int)max(-32768LL, min(32767LL, v));
    }
};

// Pass 1 of one line: N dequantised inputs read with stride `sstride` from `src`, results to dst[n*dstride].
//   Baseline: no shift, s32 results.  IQT: (sum + 64) >> 7 clipped to s16 (xevdm_itdq.c:35-39,714-716).
template <int N, bool IQT, typename LoadFn, typename StoreFn>
__device__ __forceinline__ void itx_line(LoadFn load, StoreFn store, int shift)
{
    int in[N], out[N];
#pragma unroll
    for (int k = 0; k < N; k++) in[k] = load(k);
    InvDct2<N>::run(in, out);
#pragma unroll
    for (int n = 0; n < N; n++) {
        int v = out[n];
        if (shift > 0) v = (v + (1 << (shift - 1))) >> shift;
        if (IQT || shift > 0) v = xb_clip16(v);
        store(n, v);
    }
}

// ---- ATS (Main, tool_ats): inverse DST-7 / DCT-8, full matrix product per line --------------------------------------------
// xevdm_itrans_ats_intra_{DST7,DCT8}_B{4,8,16,32} (src_main/xevdm_itdq.c:163-402): out[n] = clip16((sum_k inv[n][k] in[k] + rnd) >> shift).
// The matrices are generated on the host with the reference's own double-precision expression (xevdm_itdq.c:81-159)
// at context creation and uploaded here: [type 0 = DCT-8, 1 = DST-7][log2 n = 2..5 at offsets 0, 16, 80, 336], inv[n][k] at n * N + k.
__device__ int16_t g_ats_inv[2][1360];
__device__ __forceinline__ int ats_offset(int log2n) { return log2n == 2 ? 0 : (log2n == 3 ? 16 : (log2n == 4 ? 80 : 336)); }

template <int N, typename LoadFn, typename StoreFn>
__device__ __forceinline__ void ats_line(const int16_t *__restrict__ m, LoadFn load, StoreFn store, int shift)
{
    int in[N];
#pragma unroll
    for (int k = 0; k < N; k++) in[k] = load(k);
    const int rnd = 1 << (shift - 1);
#pragma unroll 4
    for (int n = 0; n < N; n++) {
        int acc = rnd;
#pragma unroll
        for (int k = 0; k < N; k++) acc += (int)__ldg(m + n * N + k) * in[k];
        store(n, xb_clip16(acc >> shift));
    }
}
// type: 0 = DST-7, 1 = DCT-8 (the ats_mode bit, xevd_tbl_tr_subset_intra, xevdm_tbl.c:51)
template <typename LoadFn, typename StoreFn>
__device__ __forceinline__ void ats_line_dyn(int log2n, int type, LoadFn load, StoreFn store, int shift)
{
    const int16_t *m = g_ats_inv[type ? 0 : 1] + ats_offset(log2n);
    switch (log2n) {
    case 2: ats_line<4>(m, load, store, shift); break;
    case 3: ats_line<8>(m, load, store, shift); break;
    case 4: ats_line<16>(m, load, store, shift); break;
    default: ats_line<32>(m, load, store, shift); break;
    }
}

// dispatch on the line length (2..64)
template <bool IQT, typename LoadFn, typename StoreFn>
__device__ __forceinline__ void itx_line_dyn(int log2n, LoadFn load, StoreFn store, int shift)
{
    switch (log2n) {
    case 1: itx_line<2, IQT>(load, store, shift); break;
    case 2: itx_line<4, IQT>(load, store, shift); break;
    case 3: itx_line<8, IQT>(load, store, shift); break;
    case 4: itx_line<16, IQT>(load, store, shift); break;
    case 5: itx_line<32, IQT>(load, store, shift); break;
    default: itx_line<64, IQT>(load, store, shift); break;
    }
}

}  // namespace xb

// ---------------------------------------------------------------------------------------------------------------
// Packed first pass: inputs are dequantised coefficients (s16 range), kernel entries are s8, so two
// multiply-adds fit one IDP.2A (dp2a: s16x2 . s8x2 + s32).  Same even/odd factorisation; the pairs
// (X[k], X[k+2s]) are packed once per level and reused by every output of that level.
// Measured on B200 (profiles/r1/int_pipes_b200.txt): IDP.2A and IMAD both issue at 64 lanes/clk/SM, so this
// halves the multiplier-pipe work of the pass.
namespace xb {

__device__ __forceinline__ constexpr int pk8(int lo, int hi) { return (lo & 0xff) | ((hi & 0xff) << 8); }
__device__ __forceinline__ int pack16(int lo, int hi) { return __byte_perm(lo, hi, 0x5410); }
__device__ __forceinline__ int pack_sat16(int lo, int hi)
{
    int r;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(r) : "r"(hi), "r"(lo));
    return r;
}

// M-point inverse DCT-2 over v[S * k], k = 0..M-1; every input is used by exactly one pair, and the saturating
// pack (I2IP.S16.S32.SAT) applies the s16 clip of xevd_dquant on the way
// two kernel-entry pairs share one 32-bit constant (dp2a.lo takes bytes 0-1, dp2a.hi bytes 2-3): the constants travel through uniform
// registers (one UMOV each), so this halves them
template <int M, int S, int NV> struct InvDct2P {
    static __device__ __forceinline__ void run(const int (&v)[NV], int (&out)[M])
    {
        int E[M / 2];
        InvDct2P<M / 2, 2 * S, NV>::run(v, E);
        int pr[M / 4];
#pragma unroll
        for (int j = 0; j < M / 4; j++) pr[j] = pack_sat16(v[S * (4 * j + 1)], v[S * (4 * j + 3)]);
#pragma unroll
        for (int n = 0; n < M / 2; n++) {
            int o = 0;
            if constexpr (M / 4 == 1) o = __dp2a_lo(pr[0], pk8(tm<M>(1, n), tm<M>(3, n)), o);
            else {
#pragma unroll
                for (int j = 0; j < M / 4; j += 2) {
                    const int k = pk8(tm<M>(4 * j + 1, n), tm<M>(4 * j + 3, n)) | (pk8(tm<M>(4 * j + 5, n), tm<M>(4 * j + 7, n)) << 16);
                    o = __dp2a_lo(pr[j], k, o);
                    o = __dp2a_hi(pr[j + 1 < M / 4 ? j + 1 : j], k, o);
                }
            }
            out[n] = E[n] + o;
            out[M - 1 - n] = E[n] - o;
        }
    }
    // two independent lines with the same constants
    static __device__ __forceinline__ void run2(const int (&v)[NV], const int (&w)[NV], int (&out)[M], int (&outw)[M])
    {
        int E[M / 2], F[M / 2];
        InvDct2P<M / 2, 2 * S, NV>::run2(v, w, E, F);
        int pr[M / 4], qr[M / 4];
#pragma unroll
        for (int j = 0; j < M / 4; j++) { pr[j] = pack_sat16(v[S * (4 * j + 1)], v[S * (4 * j + 3)]); qr[j] = pack_sat16(w[S * (4 * j + 1)], w[S * (4 * j + 3)]); }
#pragma unroll
        for (int n = 0; n < M / 2; n++) {
            int o = 0, p = 0;
            if constexpr (M / 4 == 1) {
                const int k = pk8(tm<M>(1, n), tm<M>(3, n));
                o = __dp2a_lo(pr[0], k, o); p = __dp2a_lo(qr[0], k, p);
            } else {
#pragma unroll
                for (int j = 0; j < M / 4; j += 2) {
                    const int k = pk8(tm<M>(4 * j + 1, n), tm<M>(4 * j + 3, n)) | (pk8(tm<M>(4 * j + 5, n), tm<M>(4 * j + 7, n)) << 16);
                    o = __dp2a_lo(pr[j], k, o); p = __dp2a_lo(qr[j], k, p);
                    o = __dp2a_hi(pr[j + 1 < M / 4 ? j + 1 : j], k, o); p = __dp2a_hi(qr[j + 1 < M / 4 ? j + 1 : j], k, p);
                }
            }
            out[n] = E[n] + o; out[M - 1 - n] = E[n] - o;
            outw[n] = F[n] + p; outw[M - 1 - n] = F[n] - p;
        }
    }
};
template <int S, int NV> struct InvDct2P<2, S, NV> {
    static __device__ __forceinline__ void run(const int (&v)[NV], int (&out)[2])
    {
        const int p = pack_sat16(v[0], v[S]);
        const int k = pk8(64, 64) | (pk8(64, -64) << 16);
        out[0] = __dp2a_lo(p, k, 0);
        out[1] = __dp2a_hi(p, k, 0);
    }
    static __device__ __forceinline__ void run2(const int (&v)[NV], const int (&w)[NV], int (&out)[2], int (&outw)[2])
    {
        const int p = pack_sat16(v[0], v[S]), q = pack_sat16(w[0], w[S]);
        const int k = pk8(64, 64) | (pk8(64, -64) << 16);
        out[0] = __dp2a_lo(p, k, 0); out[1] = __dp2a_hi(p, k, 0);
        outw[0] = __dp2a_lo(q, k, 0); outw[1] = __dp2a_hi(q, k, 0);
    }
};

// second pass on s32 inputs with the rounding constant folded into the 2-point base case
template <int N> struct InvDct2R {
    static __device__ __forceinline__ void run(const int (&in)[N], int (&out)[N], int rnd)
    {
        int ev[N / 2], E[N / 2];
#pragma unroll
        for (int k = 0; k < N / 2; k++) ev[k] = in[2 * k];
        InvDct2R<N / 2>::run(ev, E, rnd);
#pragma unroll
        for (int n = 0; n < N / 2; n++) {
            int o = 0;
#pragma unroll
            for (int k = 1; k < N; k += 2) o += tm<N>(k, n) * in[k];
            out[n] = E[n] + o;
            out[N - 1 - n] = E[n] - o;
        }
    }
};
template <> struct InvDct2R<2> {
    static __device__ __forceinline__ void run(const int (&in)[2], int (&out)[2], int rnd)
    {
        out[0] = 64 * (in[0] + in[1]) + rnd;
        out[1] = 64 * (in[0] - in[1]) + rnd;
    }
};

}  // namespace xb
