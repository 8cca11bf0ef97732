// xb_recon2.cuh -- throughput-oriented inter reconstruction kernel (64x64 CTUs; Baseline or Main IQT transform, Baseline or Main taps).
//
// Same contract as k_recon_inter (xb_recon.cuh): one CTA reconstructs one CTU from the flat CU work items.
// What is different is how the work is laid onto the SM:
//
//   * reference windows arrive by TMA (cp.async.bulk.tensor.2d): one 40x23 (luma) / 24x11 (chroma) box per
//     16x16 tile, issued up front (every warp issues the boxes of its own two tile slots) so the HBM latency hides behind the
//     residual phase; completion by mbarrier.  The two lists of a B picture go through the same windows one after the other.
//     The hardware wants the box to start on a 16-byte boundary (measured: profiles/microbench/tma_probe*.cu), so
//     the box is the 8-sample-aligned superset of the window and the horizontal stage absorbs the 0..7 sample
//     offset (word offset by addressing, odd offsets by swapping the tap sets of even and odd outputs).
//   * the CTU's slice of the coefficient stream arrives by one cp.async.bulk into the bytes of the residual planes.
//   * residual: thread-per-line butterflies in registers.  The Baseline 2-D inverse DCT is an exact integer
//     matrix product modulo 2^32, so the row pass is done first on the s16 inputs with IDP.2A
//     (packed s16x2 . s8x2) and the column pass second with IMAD on the s32 intermediate.  IQT rounds the first pass to s16,
//     so the IQT variant runs the reference's order: columns first (strided reads of the staged coefficients), rows second.
//   * Main tools it has no code for (ATS, DMVR, affine): per-CU dispatch - those CUs are marked and treated as absent; the
//     generic kernel, launched next, reconstructs exactly them (cu_needs_generic, xb_recon.cuh).
//   * interpolation: IDP.2A on packed pixel pairs; the horizontal stage emits vertical pairs so the vertical
//     stage needs no repacking; variants 00/n0/0n/nn are one code path (phase-0 taps are an exact copy).
//   * reconstruction: VIADDMNMX.S16x2.RELU = clip3(0, max, (s16)(pred + resid)) for two samples per instruction.
//
// Reference arithmetic restated: xevd_mc (src_base/xevd_mc.c:169-557), xevd_itdq / xevd_dquant
// (src_base/xevd_itdq.c:472-542), xevd_recon (src_base/xevd_recon.c:36-68), xevd_set_dec_info
// (src_base/xevd_util.c:1574-1650).
#pragma once
#include <cuda.h>
#include "xb_common.cuh"
#include "xb_itdq.cuh"
#include "xb_recon.cuh"

namespace xb {

constexpr int kR2Threads = 256;
// Two ways through the prediction stages (template parameter WS of the kernel):
//   WS (warp slots)  8 tile slots, one per warp: a warp takes its tiles (warp, warp + 8, ...) through both stages and the reconstruction
//                    alone, nothing block-wide after the residual; 72 KB and 80 registers -> three CTAs per SM.  Pictures of large CUs.
//   !WS (rounds)     16 slots filled per round, two per warp in the horizontal stage, 16 threads per slot in the vertical stage:
//                    half the per-tile passes when the tiles are small (8x8 CUs), two CTAs per SM.  Pictures of many small CUs.
constexpr int kTileCapWS = 8, kTileCapRounds = 16;
constexpr int kBoxLW = 40, kBoxLH = 23;      // luma TMA box: (offset <= 7) + 16 + 7 = 30 -> 40 samples keeps the row stride
constexpr int kBoxCW = 24, kBoxCH = 11;      //   at 20 words (8 rows = 8 distinct bank quads); chroma: 7 + 8 + 3 = 18 -> 24
constexpr int kWinLBytes = 1920;             // 40 x 23 x 2 = 1840, rounded to a multiple of 128
constexpr int kWinCPlane = kBoxCW * kBoxCH * 2;   // 528 bytes: one chroma plane of a window; Cb and Cr arrive as ONE 3-D box, back to back
constexpr int kWinCBytes = 1152;             // 2 x 528 = 1056, rounded to a multiple of 128
constexpr int kWinLStrideW = kBoxLW / 2;     // window row stride in 32-bit words
constexpr int kWinCStrideW = kBoxCW / 2;
constexpr int kM2LStrideW = 20;              // vertical-pair buffer row stride (16 columns + 4 pad), words
constexpr int kM2LWords = 12 * kM2LStrideW;  // 12 pair-rows
// pass-1 results and the residual: rows 0..63 hold luma, rows 64..95 hold Cb (columns 0..31) and Cr (columns 32..63) side by side, so one
// compile-time row stride serves every plane (the strided accesses of the passes become immediate offsets)
constexpr int kTmpStride = 68;               // pass-1 result row stride, words (64 + 4: int4 stores of 8 consecutive rows hit 8 distinct bank quads)
constexpr int kResStride = 72;               // residual row stride, int16 (64 + 8)
constexpr int kPlaneRows = 96;
__host__ __device__ __forceinline__ constexpr int plane_origin(int pl, int stride) { return pl == 0 ? 0 : 64 * stride + (pl == 2 ? 32 : 0); }

constexpr int kCoefStageBytes = 2 * 2 * 64 * 64;
constexpr int kBatchCap = 34;                // a CTU has at most 1024 luma and 1024 chroma lines per pass
constexpr uint8_t kCuOtherKernel = 0x80;     // private bit in the staged copy of XB200_CU.flags: this CU belongs to the generic kernel

struct TuDesc {                              // one transform block (16 bytes)
    uint32_t coef_off;                       // first coefficient, int16 units
    uint16_t tmp_off;                        // word offset of (0,0) inside the plane's pass-1 buffer
    uint16_t res_off;                        // int16 offset of (0,0) inside the plane's residual buffer
    uint8_t lw_lh;                           // log2w | log2h << 4
    uint8_t plane_wide;                      // plane | wide << 2
    uint8_t cstride_log2;                    // log2 of the coefficient row stride
    uint8_t shift;
    int32_t mul;
};

struct TileDesc {                            // one <=16x16 luma tile of a CU (8 bytes) + per-list data
    uint16_t cu;
    uint8_t px, py;                          // origin inside the CTU (luma samples)
    uint8_t tw, th;                          // luma size
    uint8_t nl;                              // number of prediction lists actually used (1 or 2)
    uint8_t pad;
};
struct TilePred {                            // 16 bytes, per (tile, used list)
    int16_t wx, wy;                          // luma window origin, padded-plane coordinates
    int16_t cwx, cwy;                        // chroma window origin
    uint8_t phx, phy, cphx, cphy;            // tap table rows (clipped vector)
    uint8_t two_d, ctwo_d;                   // variant nn selected by the UNCLIPPED vector (T3)
    uint8_t ref;                             // list * XB_MAX_REFS + refi
    uint8_t offs;                            // sample offset of the window inside its 8-aligned box: luma | chroma << 4
};

// Dynamic shared memory map.  Transform blocks are kept in two tables (luma first, then chroma) so that a warp's
// 32 consecutive line tasks run the same butterfly size; line -> block lookup is a binary search over the
// per-block line prefix sums (no per-line lists).
struct R2Layout {
    int win_l, win_c, scratch, res_y, coef, cus, tus, pre1, pre2, batch, tiles, preds, offs, taps, out, total;
    __host__ __device__ static R2Layout make(int nl, int max_cu, bool peer = false, bool ws = false)
    {
        const int kTileCap = ws ? kTileCapWS : kTileCapRounds;
        R2Layout L;
        int o = 128;                                                       // [0,128): mbarrier + counters
        // windows and vertical-pair buffers hold ONE prediction list at a time: the lists of a bi-predicted picture go through them one
        // after the other (the running prediction lives in registers), so B pictures keep two CTAs per SM like P pictures
        (void)nl;
        L.win_l = o; o += kTileCap * kWinLBytes;
        L.win_c = o; o += kTileCap * kWinCBytes;
        const int tmp_bytes = 4 * kPlaneRows * kTmpStride;
        const int m2_bytes = 4 * kTileCap * kM2LWords;
        L.scratch = o; o += tmp_bytes > m2_bytes ? tmp_bytes : m2_bytes;
        // residual planes; before the row pass the same bytes stage the CTU's slice of the coefficient stream, fetched by one bulk copy
        // while the descriptors are built (at most 2 int16 per luma sample: 4x4 CUs with their chroma blocks padded to 8)
        const int res_bytes = 2 * kPlaneRows * kResStride;
        L.res_y = o; L.coef = o; o += res_bytes > kCoefStageBytes ? res_bytes : kCoefStageBytes;
        L.cus = o; o += 32 * max_cu;
        L.tus = o; o += 16 * 3 * max_cu;
        L.pre1 = o; o += 2 * (3 * max_cu + 2);          // uint16 prefix of pass-1 lines per block (+ end markers)
        L.pre2 = o; o += 2 * (3 * max_cu + 2);
        L.batch = o; o += 2 * 4 * kBatchCap;            // block holding the first line of every 32-line batch: [pass][luma / chroma][batch]
        o = (o + 15) & ~15;
        const int max_tiles = max_cu + 16;              // a CU larger than 16x16 is several tiles: at most 15 extra per CTU
        L.tiles = o; o += 8 * max_tiles;
        L.preds = o; o += 16 * 2 * max_tiles;
        L.offs = o; o += 4 * 8 * max_cu;                // per-CU exclusive offsets (scan output)
        L.taps = o; o += 4 * (16 * 9 + 32 * 6);
        o = (o + 15) & ~15;
        L.out = o; if (peer) o += 2 * (64 * 64 + 2 * 32 * 32) + 256 * (8 + 4 + 2 + 1) + 16;   // PEER: reconstructed CTU + its map entries staged for wide row stores to every GPU
        L.total = (o + 127) & ~127;
        return L;
    }
};

// ---- small PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra W;\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const void *tmap, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// Packed tap words for IDP.2A.  A filter applied to packed sample pairs p[0..4] (5 words = 10 samples) is a sequence
// of five (lo, hi) tap pairs T0..T4, stored as three registers r0 = T0 | T1 << 16, r1 = T2 | T3 << 16, r2 = T4.
//   output aligned to p[0].lo          : A0 = (c0,c1)(c2,c3)(c4,c5)(c6,c7)(0,0)
//   output starting at p[0].hi         : B  = (0,c0)(c1,c2)(c3,c4)(c5,c6)(c7,0)
//   output aligned to p[1].lo          : A1 = (0,0)(c0,c1)(c2,c3)(c4,c5)(c6,c7)
// With an even sample offset even outputs use A0 and odd outputs B; with an odd offset even outputs use B and odd
// outputs A1 -- the same five words either way, so the stage is branch-free.
struct Taps5 { int r0, r1, r2; };
__device__ __forceinline__ int fir5(const Taps5 &t, int p0, int p1, int p2, int p3, int p4, int acc)
{
    acc = __dp2a_lo(p0, t.r0, acc); acc = __dp2a_hi(p1, t.r0, acc);
    acc = __dp2a_lo(p2, t.r1, acc); acc = __dp2a_hi(p3, t.r1, acc);
    acc = __dp2a_lo(p4, t.r2, acc);
    return acc;
}
// 4-tap: three pairs T0..T2 in two registers (r0 = T0 | T1 << 16, r1 = T2)
struct Taps3 { int r0, r1; };
__device__ __forceinline__ int fir3(const Taps3 &t, int p0, int p1, int p2, int acc)
{
    acc = __dp2a_lo(p0, t.r0, acc); acc = __dp2a_hi(p1, t.r0, acc);
    acc = __dp2a_lo(p2, t.r1, acc);
    return acc;
}
__host__ __device__ __forceinline__ int pk2(int lo, int hi) { return (lo & 0xff) | ((hi & 0xff) << 8); }
// tap-set tables: luma [phase][set 0..2 = A0, B, A1][3 regs], chroma [phase][set][2 regs]; built on the host at
// context creation (xb_build_tap_tables) and copied into shared memory by every CTA
// (global memory, not __constant__: every CTA copies them into shared memory with one element per thread, and a constant-bank load
// with 32 different addresses per warp is served one address at a time)
__device__ int c_taps5[2][16 * 9];
__device__ int c_taps3[2][32 * 6];
__host__ __device__ __forceinline__ void build_taps8(const int16_t *c, int *dst)
{
    const int a0 = pk2(c[0], c[1]), a1 = pk2(c[2], c[3]), a2 = pk2(c[4], c[5]), a3 = pk2(c[6], c[7]);
    const int b0 = pk2(0, c[0]), b1 = pk2(c[1], c[2]), b2 = pk2(c[3], c[4]), b3 = pk2(c[5], c[6]), b4 = pk2(c[7], 0);
    dst[0] = a0 | (a1 << 16); dst[1] = a2 | (a3 << 16); dst[2] = 0;
    dst[3] = b0 | (b1 << 16); dst[4] = b2 | (b3 << 16); dst[5] = b4;
    dst[6] = (a0 << 16);      dst[7] = a1 | (a2 << 16); dst[8] = a3;
}
__host__ __device__ __forceinline__ void build_taps4(const int16_t *c, int *dst)
{
    const int a0 = pk2(c[0], c[1]), a1 = pk2(c[2], c[3]);
    const int b0 = pk2(0, c[0]), b1 = pk2(c[1], c[2]), b2 = pk2(c[3], 0);
    dst[0] = a0 | (a1 << 16); dst[1] = 0;
    dst[2] = b0 | (b1 << 16); dst[3] = b2;
    dst[4] = (a0 << 16);      dst[5] = a1;
}

// ---- residual passes ---------------------------------------------------------------------------------------------------------------
// A task is one line of a transform block or - Baseline transform, lines of at most 16 points - two neighbouring lines: two rows share
// the descriptor lookup and the kernel constants, two columns are loaded as 8-byte words and stored as packed sample pairs.
__host__ __device__ __forceinline__ int row_tasks(int lw, int lh, bool iqt) { return (!iqt && lw <= 4) ? (1 << lh) >> 1 : 1 << lh; }
__host__ __device__ __forceinline__ int col_tasks(int lw, int lh, bool iqt) { return (!iqt && lh <= 4) ? (1 << lw) >> 1 : 1 << lw; }

// N coefficients of one line -> dequantised values.  xevd_dquant: clip16((c * scale + offset) >> shift); the clip happens in the saturating
// pack (I2IP.S16.S32.SAT) that forms the butterfly's operand pairs.  Baseline reads a ROW (16-byte loads); IQT a COLUMN (its first pass
// rounds to s16, so it has to be the reference's first pass)
template <int N, bool IQT>
__device__ __forceinline__ void load_dequant(const int16_t *__restrict__ src, int sstride, int (&v)[N], int mul, int off, int shift, bool wide)
{
    if (IQT) {
#pragma unroll
        for (int k = 0; k < N; k++) v[k] = src[k * sstride];
    } else if (N >= 8) {
#pragma unroll
        for (int q = 0; q < N / 8; q++) {
            const int4 w = ((const int4 *)src)[q];
            const int ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int j = 0; j < 4; j++) { v[8 * q + 2 * j] = (int)(int16_t)(ww[j] & 0xffff); v[8 * q + 2 * j + 1] = ww[j] >> 16; }
        }
    } else if (N == 4) {
        const int2 w = *(const int2 *)src;
        v[0] = (int)(int16_t)(w.x & 0xffff); v[1] = w.x >> 16; v[2] = (int)(int16_t)(w.y & 0xffff); v[3] = w.y >> 16;
    } else {
        const int w = *(const int *)src;
        v[0] = (int)(int16_t)(w & 0xffff); v[1] = w >> 16;
    }
    if (!wide) {
#pragma unroll
        for (int k = 0; k < N; k++) v[k] = (v[k] * mul + off) >> shift;
    } else {
#pragma unroll
        for (int k = 0; k < N; k++) v[k] = (int)(((long long)v[k] * mul + (long long)off) >> shift);
    }
}
template <int N, bool IQT>
__device__ __forceinline__ void store_pass1(const int (&out)[N], int *__restrict__ dst, const int ts)
{
    if (IQT) {          // Main IQT: same kernel, first pass rounded to s16 (xevdm_itdq.c:35-39,714-716); the line is a column of the block
#pragma unroll
        for (int k = 0; k < N; k++) dst[k * ts] = xb_clip16((out[k] + 64) >> 7);
        return;
    }
    if (N >= 4 && ((smem_u32(dst) & 15) == 0)) {
#pragma unroll
        for (int q = 0; q < N / 4; q++) ((int4 *)dst)[q] = make_int4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
    } else {            // chroma blocks of ternary-split CUs can start on an 8-byte boundary only
#pragma unroll
        for (int q = 0; q < N / 2; q++) ((int2 *)dst)[q] = make_int2(out[2 * q], out[2 * q + 1]);
    }
}
template <int N, bool IQT>
__device__ __forceinline__ void row_pass(const int16_t *__restrict__ src, int sstride, int *__restrict__ dst, const int ts, int mul, int off, int shift, bool wide)
{
    int v[N], out[N];
    load_dequant<N, IQT>(src, sstride, v, mul, off, shift, wide);
    InvDct2P<N, 1, N>::run(v, out);
    store_pass1<N, IQT>(out, dst, ts);
}
// two neighbouring rows of a Baseline block (row stride of the coefficients: sstride; of the results: ts)
template <int N>
__device__ __forceinline__ void row_pass2(const int16_t *__restrict__ src, int sstride, int *__restrict__ dst, const int ts, int mul, int off, int shift, bool wide)
{
    int v0[N], v1[N], o0[N], o1[N];
    load_dequant<N, false>(src, sstride, v0, mul, off, shift, wide);
    load_dequant<N, false>(src + sstride, sstride, v1, mul, off, shift, wide);
    InvDct2P<N, 1, N>::run2(v0, v1, o0, o1);
    store_pass1<N, false>(o0, dst, ts);
    store_pass1<N, false>(o1, dst + ts, ts);
}

// second pass, one line.  Baseline: a column (stride ts in, rs out); IQT: a row, contiguous
template <int N, bool IQT>
__device__ __forceinline__ void col_pass(const int *__restrict__ src, const int ts, int16_t *__restrict__ dst, const int rs, int sh2)
{
    int in[N], out[N];
    if (IQT) {
#pragma unroll
        for (int k = 0; k < N / 2; k++) { const int2 v = ((const int2 *)src)[k]; in[2 * k] = v.x; in[2 * k + 1] = v.y; }
    } else {
#pragma unroll
        for (int k = 0; k < N; k++) in[k] = src[k * ts];
    }
    InvDct2R<N>::run(in, out, 1 << (sh2 - 1));
    if (IQT) {
#pragma unroll
        for (int k = 0; k < N / 2; k++) ((int *)dst)[k] = pack_sat16(out[2 * k] >> sh2, out[2 * k + 1] >> sh2);
    } else {
#pragma unroll
        for (int k = 0; k < N; k++) dst[k * rs] = (int16_t)pack_sat16(out[k] >> sh2, 0);
    }
}
// two neighbouring columns of a Baseline block: 8-byte loads, packed 4-byte stores
template <int N>
__device__ __forceinline__ void col_pass2(const int *__restrict__ src, const int ts, int16_t *__restrict__ dst, const int rs, int sh2)
{
    int a[N], b[N], oa[N], ob[N];
#pragma unroll
    for (int k = 0; k < N; k++) { const int2 v = *(const int2 *)(src + k * ts); a[k] = v.x; b[k] = v.y; }
    InvDct2R<N>::run(a, oa, 1 << (sh2 - 1));
    InvDct2R<N>::run(b, ob, 1 << (sh2 - 1));
#pragma unroll
    for (int k = 0; k < N; k++) *(int *)(dst + k * rs) = pack_sat16(oa[k] >> sh2, ob[k] >> sh2);
}

// one past the last coefficient of a CU inside the stream (int16 units): its coded plane blocks, each padded to 8 entries; an ats_inter CU
// carries only its sub-block transform unit (a full-CU figure here once made the CTU's bulk copy read past the end of the stream)
__device__ __forceinline__ int cu_coef_end(const XbFrameArgs &a, const XB200_CU &cu)
{
    int tlw = cu.log2w, tlh = cu.log2h;
    const int aidx = a.ats ? ats_inter_idx(cu) : 0;
    if (aidx) { int xo, yo; ats_inter_tu(cu, aidx, tlw, tlh, xo, yo); }
    const int n = 1 << (tlw + tlh);
    int end = cu.coef_off;
    if (cu.cbf & 0x00f) end += (n + 7) & ~7;
    if (cu.cbf & 0x0f0) end += ((n >> 2) + 7) & ~7;
    if (cu.cbf & 0xf00) end += ((n >> 2) + 7) & ~7;
    return end;
}

// PEER (band mode over NVLink): the reconstructed CTU is collected in shared memory and written out as whole 128-byte rows to the
// local picture AND to its twins on the peer GPUs, so the exchange rides on the kernel's own stores at full NVLink request size.
template <bool BI, bool PEER = false, bool IQT = false, bool DISP = false, bool WS = false>
__global__ void __launch_bounds__(kR2Threads, WS ? 3 : 2)
k_recon_inter_v2(const __grid_constant__ XbFrameArgs a, const int max_cu)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NL = BI ? 2 : 1;
    constexpr int kTileCap = WS ? kTileCapWS : kTileCapRounds;
    const R2Layout L = R2Layout::make(NL, max_cu, PEER, WS);
    int16_t *s_out = (int16_t *)(smem + L.out);                 // [64][64] luma, then [2][32][32] chroma
    uint64_t *mbar = (uint64_t *)smem, *mbar_coef = (uint64_t *)(smem + 64);     // [0,64): WS: one mbarrier per tile slot (= per warp); !WS: the first one
    int *cnt = (int *)(smem + 72);           // totals: [0] luma blocks [1] chroma blocks [2,3] pass-1 lines y,c [4,5] pass-2 y,c [6] tiles
    XB200_CU *s_cu = (XB200_CU *)(smem + L.cus);
    TuDesc *s_tu = (TuDesc *)(smem + L.tus);
    uint16_t *s_pre1 = (uint16_t *)(smem + L.pre1), *s_pre2 = (uint16_t *)(smem + L.pre2);
    uint16_t *s_bat1 = (uint16_t *)(smem + L.batch), *s_bat2 = s_bat1 + 2 * kBatchCap;
    TileDesc *s_tile = (TileDesc *)(smem + L.tiles);
    TilePred *s_pred = (TilePred *)(smem + L.preds);
    int *s_offs = (int *)(smem + L.offs);
    int *s_tmp = (int *)(smem + L.scratch);
    int16_t *s_res = (int16_t *)(smem + L.res_y);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctu = blockIdx.y * a.w_ctu + blockIdx.x;
    const int ctu_x = blockIdx.x << 6, ctu_y = (blockIdx.y + a.ctu_row0) << 6;
    // the host entry points route pictures with more CUs per CTU than the lists hold to the generic kernel; the clamp keeps a wrong
    // max_cu_per_ctu handed to the _dev entry from overrunning shared memory
    xb_grid_wait();            // launched early (xb_launch_early): the work lists, the reference pictures and this picture belong to the kernels before
    xb_grid_release();
    const int cu0 = a.ctu_first[ctu], ncu = min((int)(a.ctu_first[ctu + 1] - cu0), max_cu);

    // ---- coefficient slice of this CTU: CUs are in decoding order, so their blocks are one contiguous range of the stream.  One bulk
    //      copy brings it on chip while the descriptors are staged and built; the first pass, three barriers later, reads shared memory
    //      instead of waiting on DRAM.  Thread 0 reads the first and the last CU straight from global memory so that the copy is in flight
    //      before anything else happens (the wait for it was 10 % of all stall samples when it was issued after the staging barrier).
    int coef_base = 0, coef_bytes = 0;
    if (tid == 0) {
        for (int k = 0; k < 8; k++) mbar_init((uint64_t *)smem + k, 1);
        mbar_init(mbar_coef, 1);
        if (ncu > 0) {
            const int4 f0 = __ldg((const int4 *)(a.cus + cu0) + 1), l0 = __ldg((const int4 *)(a.cus + cu0 + ncu - 1)), l1 = __ldg((const int4 *)(a.cus + cu0 + ncu - 1) + 1);
            XB200_CU c1;
            ((int4 *)&c1)[0] = l0; ((int4 *)&c1)[1] = l1;
            coef_base = f0.w;                                   // XB200_CU.coef_off is the last word of the record
            coef_bytes = min(2 * (cu_coef_end(a, c1) - coef_base), kCoefStageBytes);     // (the clamp only matters for a malformed device-resident list)
            if (coef_bytes > 0) {
                mbar_expect_tx(mbar_coef, (uint32_t)coef_bytes);
                bulk_load(smem + L.coef, a.coef + coef_base, (uint32_t)coef_bytes, mbar_coef);
            }
        }
    }
    // ---- stage CU descriptors, build tap tables ------------------------------------------------------------------------------
    {
        const int4 *g = (const int4 *)(a.cus + cu0);
        int4 *s = (int4 *)s_cu;
        for (int i = tid; i < ncu * 2; i += kR2Threads) s[i] = __ldg(g + i);
        int *t8 = (int *)(smem + L.taps), *t4 = t8 + 16 * 9;
        if (tid < 16 * 9) t8[tid] = __ldg(&c_taps5[a.main_tables][tid]);
        if (tid < 32 * 6) t4[tid] = __ldg(&c_taps3[a.main_tables][tid]);
    }
    const int *s_t8 = (const int *)(smem + L.taps), *s_t4 = s_t8 + 16 * 9;
    auto ld_taps5 = [&](int ph, int set) { Taps5 t; const int *q = s_t8 + (ph & 15) * 9 + set * 3; t.r0 = q[0]; t.r1 = q[1]; t.r2 = q[2]; return t; };
    auto ld_taps3 = [&](int ph, int set) { Taps3 t; const int *q = s_t4 + (ph & 31) * 6 + set * 2; t.r0 = q[0]; t.r1 = q[1]; return t; };
    __syncthreads();
    // Per-CU dispatch (Main tools): CUs with ATS / DMVR / affine are reconstructed by the generic kernel, which is launched next and takes
    // exactly those; here they are marked and then treated as absent (no transform blocks, no tiles, no map entries)
    if (DISP) {
        for (int i = tid; i < ncu; i += kR2Threads)
            if (cu_needs_generic(a, s_cu[i])) s_cu[i].flags |= kCuOtherKernel;
        __syncthreads();
    }

    // every thread needs the slice's origin and size (descriptors are relative to it; the first pass waits only if there is one)
    if (ncu > 0) {
        const XB200_CU c0 = s_cu[0], c1 = s_cu[ncu - 1];
        coef_base = c0.coef_off;
        coef_bytes = min(2 * (cu_coef_end(a, c1) - coef_base), kCoefStageBytes);
    }

    // ---- per-CU counts -> exclusive prefix sums: warp q scans quantity q -------------------------------------------------------
    // q: 0 luma blocks, 1 chroma blocks, 2 row-direction luma tasks, 3 row-direction chroma tasks, 4 column-direction luma tasks, 5 column-direction
    //    chroma tasks, 6 tiles (a task = one line, or two neighbouring lines: row_tasks / col_tasks)
    if (warp == 7) {            // number of intra / IBC CUs (their residual is parked in the picture below)
        int n = 0;
        for (int i = lane; i < ncu; i += 32) n += (xb_wavefront_mode(s_cu[i].mode) && !(DISP && (s_cu[i].flags & kCuOtherKernel))) ? 1 : 0;
        n = __reduce_add_sync(0xffffffffu, n);
        if (lane == 0) cnt[7] = n;
    }
    if (warp < 7) {
        int run = 0;
        for (int base = 0; base < ncu; base += 32) {
            const int i = base + lane;
            int c = 0;
            if (i < ncu) {
                const XB200_CU cu = s_cu[i];
                const int w = 1 << cu.log2w, h = 1 << cu.log2h;
                // intra / IBC CUs are predicted by the wavefront kernel, but their residual does not depend on neighbours: it is
                // transformed here with everything else and parked in the picture for that kernel to pick up
                const bool mine = !(DISP && (cu.flags & kCuOtherKernel));
                const bool inter = mine && !xb_wavefront_mode(cu.mode);
                const int ny = (mine && (cu.cbf & 15)) ? 1 : 0, nc = mine ? ((cu.cbf & 0x0f0) ? 1 : 0) + ((cu.cbf & 0xf00) ? 1 : 0) : 0;
                switch (warp) {
                case 0: c = ny; break;
                case 1: c = nc; break;
                case 2: c = ny * row_tasks(cu.log2w, cu.log2h, IQT); break;
                case 3: c = nc * row_tasks(cu.log2w - 1, cu.log2h - 1, IQT); break;
                case 4: c = ny * col_tasks(cu.log2w, cu.log2h, IQT); break;
                case 5: c = nc * col_tasks(cu.log2w - 1, cu.log2h - 1, IQT); break;
                default: c = inter ? max(1, w >> 4) * max(1, h >> 4) : 0; break;
                }
            }
            int v = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
            if (i < ncu) s_offs[i * 8 + warp] = run + v - c;
            run += __shfl_sync(0xffffffffu, v, 31);
        }
        if (lane == 0) cnt[warp] = run;
    }
    __syncthreads();
    const int n_tuy = cnt[0], n_tuc = cnt[1], n_l1y = cnt[2], n_l1c = cnt[3], n_l2y = cnt[4], n_l2c = cnt[5], n_tiles = cnt[6];
    const int n_tu = n_tuy + n_tuc;

    // ---- descriptors: warp role = (luma block, Cb block, Cr block, tiles) x CU chunk, one CU per lane.  (One thread per CU doing all
    //      four in a row left seven warps waiting at the barrier below for ~6 % of the kernel's time.)
    const int role = warp & 3;
    // the 64 lanes of a role take one CU each, except the tile role when the CTU has few (hence large) CUs: then a group of lanes shares a
    // CU and every lane builds one of its 16x16 tiles (the 16 tiles of a 64x64 CU were built by one thread, one after the other)
    const int r_id = (warp >> 2) * 32 + lane;
    const int lgt = role < 3 || ncu > 8 ? 6 : (ncu > 4 ? 3 : (ncu > 2 ? 2 : (ncu > 1 ? 1 : 0)));      // (with more CUs, spreading them over both warps of the role only costs issue slots)
    const int t_lanes = 64 >> lgt, t_q = r_id & (t_lanes - 1);
    for (int ci = r_id >> (6 - lgt); ci < ncu; ci += 1 << lgt) {
        const XB200_CU cu = s_cu[ci];
        if (DISP && (cu.flags & kCuOtherKernel)) continue;
        const bool inter_cu = !xb_wavefront_mode(cu.mode);
        const int *of = s_offs + ci * 8;
        const int w = 1 << cu.log2w, h = 1 << cu.log2h;
        const int lx = cu.x - ctu_x, ly = cu.y - ctu_y;
        if (role < 3) {
            const int pl = role;
            if (!((cu.cbf >> (4 * pl)) & 15)) continue;
            const int sh = pl ? 1 : 0;
            const int lw = cu.log2w - sh, lh = cu.log2h - sh;
            const int n_y = ((1 << (cu.log2w + cu.log2h)) + 7) & ~7, n_c = ((1 << (cu.log2w + cu.log2h - 2)) + 7) & ~7;
            const bool has_y = (cu.cbf & 0x00f) != 0, has_u = (cu.cbf & 0x0f0) != 0;
            const int coef = cu.coef_off + (pl >= 1 && has_y ? n_y : 0) + (pl == 2 && has_u ? n_c : 0);
            TuDesc d;
            d.coef_off = coef - coef_base;      // relative to the staged slice
            d.tmp_off = (uint16_t)(plane_origin(pl, kTmpStride) + (ly >> sh) * kTmpStride + (lx >> sh));
            d.res_off = (uint16_t)(plane_origin(pl, kResStride) + (ly >> sh) * kResStride + (lx >> sh));
            d.lw_lh = (uint8_t)(lw | (lh << 4));
            d.cstride_log2 = (uint8_t)lw;
            const int qp = pl == 0 ? cu.qp_y : (pl == 1 ? cu.qp_u : cu.qp_v);
            const int odd = (lw + lh) & 1;
            const int shift = 6 - (15 - a.bd_l - ((lw + lh) >> 1)) + (odd ? 8 : 0);
            const long long mul = (long long)(c_dq_scale[IQT ? 1 : 0][qp % 6] << (qp / 6)) * (odd ? 181 : 1);
            d.shift = (uint8_t)shift;
            d.mul = (int)mul;
            d.plane_wide = (uint8_t)(pl | ((mul >= 65536) ? 4 : 0));
            const int nr = row_tasks(lw, lh, IQT), ncl = col_tasks(lw, lh, IQT);
            int b, p1, p2;
            if (pl == 0) {
                b = of[0]; p1 = of[2]; p2 = of[4];
                s_tu[b] = d; s_pre1[b] = (uint16_t)p1; s_pre2[b] = (uint16_t)p2;
            } else {
                const int second = pl == 2 && has_u ? 1 : 0;          // Cr follows Cb in the chroma tables
                b = n_tuy + of[1] + second; p1 = of[3] + second * nr; p2 = of[5] + second * ncl;
                s_tu[b] = d;                                            // chroma tables sit after a luma end marker
                s_pre1[b + 1] = (uint16_t)p1; s_pre2[b + 1] = (uint16_t)p2;
            }
            // the block that holds the first task of a 32-task batch is where the passes start their lookup
            for (int k = (p1 + 31) >> 5; (k << 5) < p1 + nr; k++) s_bat1[(pl ? kBatchCap : 0) + k] = (uint16_t)b;
            for (int k = (p2 + 31) >> 5; (k << 5) < p2 + ncl; k++) s_bat2[(pl ? kBatchCap : 0) + k] = (uint16_t)b;
            continue;
        }
        // prediction tiles (inter CUs only)
        if (inter_cu) {
        int mvc[2][2];
        mv_clip(cu.x, cu.y, a.w, a.h, w, h, cu.mv[0][0], cu.mv[0][1], mvc[0][0], mvc[0][1]);
        mv_clip(cu.x, cu.y, a.w, a.h, w, h, cu.mv[1][0], cu.mv[1][1], mvc[1][0], mvc[1][1]);
        bool use0 = cu.refi[0] >= 0, use1 = cu.refi[1] >= 0;
        if (use0 && use1 && a.ref_poc[0][cu.refi[0]] == a.ref_poc[1][cu.refi[1]] && mvc[0][0] == mvc[1][0] && mvc[0][1] == mvc[1][1])
            use1 = false;                 // identical motion -> list 0 only (xevd_mc.c:513-519)
        const int tw = min(w, 16), th = min(h, 16);
        const int ltx = max(0, cu.log2w - 4), n_t = 1 << (ltx + max(0, cu.log2h - 4));
        for (int kt = t_q; kt < n_t; kt += t_lanes) {
            {
                const int tx = (kt & ((1 << ltx) - 1)) << 4, ty = (kt >> ltx) << 4, t = of[6] + kt;
                TileDesc td;
                td.cu = (uint16_t)ci; td.px = (uint8_t)(lx + tx); td.py = (uint8_t)(ly + ty);
                td.tw = (uint8_t)tw; td.th = (uint8_t)th; td.nl = (uint8_t)((use0 ? 1 : 0) + (use1 ? 1 : 0)); td.pad = 0;
                s_tile[t] = td;
                int k = 0;
#pragma unroll
                for (int l = 0; l < 2; l++) {
                    if (!(l ? use1 : use0)) continue;
                    if (k < NL) {
                        const int mvx = mvc[l][0], mvy = mvc[l][1];
                        TilePred p;
                        p.wx = (int16_t)(144 + cu.x + tx + (mvx >> 2) - 3);
                        p.wy = (int16_t)(144 + cu.y + ty + (mvy >> 2) - 3);
                        p.cwx = (int16_t)(72 + ((cu.x + tx) >> 1) + (mvx >> 3) - 1);
                        p.cwy = (int16_t)(72 + ((cu.y + ty) >> 1) + (mvy >> 3) - 1);
                        p.phx = (uint8_t)((mvx & 3) << 2); p.phy = (uint8_t)((mvy & 3) << 2);
                        p.cphx = (uint8_t)((mvx & 7) << 2); p.cphy = (uint8_t)((mvy & 7) << 2);
                        p.two_d = (uint8_t)(((cu.mv[l][0] & 3) != 0) && ((cu.mv[l][1] & 3) != 0));
                        p.ctwo_d = (uint8_t)(((cu.mv[l][0] & 7) != 0) && ((cu.mv[l][1] & 7) != 0));
                        p.ref = (uint8_t)(l * XB_MAX_REFS + cu.refi[l]);
                        p.offs = (uint8_t)((p.wx & 7) | ((p.cwx & 7) << 4));
                        s_pred[t * NL + k] = p;
                    }
                    k++;
                }
            }
        }
        }
    }
    if (tid == 0) {         // end markers of the two prefix tables: [0..n_tuy) luma, [n_tuy] end, [n_tuy+1 .. n_tu+1) chroma, [n_tu+1] end
        s_pre1[n_tuy] = (uint16_t)n_l1y; s_pre2[n_tuy] = (uint16_t)n_l2y;
        s_pre1[n_tu + 1] = (uint16_t)n_l1c; s_pre2[n_tu + 1] = (uint16_t)n_l2c;
    }
    __syncthreads();

    // ---- prediction windows: warp w owns tile slot w (window buffers, mbarrier, vertical-pair buffers) and takes the tiles w, w + 8, ... through
    //      both interpolation stages and the reconstruction on its own - no block-wide barrier after the residual is complete.  Lane 0
    //      issues the two boxes of a (tile, list) and announces their bytes on the slot's mbarrier; the first one flies during the residual
    //      passes, every later one during the vertical stage of its predecessor.
    uint64_t *mbar_w = (uint64_t *)smem + warp;
    auto issue_w = [&](int tile, int l) {
        if (lane == 0) {
            const TilePred p = s_pred[tile * NL + l];
            const CUtensorMap *tm = a.ref_tmap[p.ref];
            mbar_expect_tx(mbar_w, 2u * (kBoxLW * kBoxLH + 2 * kBoxCW * kBoxCH));
            tma_load_2d(smem + L.win_l + warp * kWinLBytes, tm + 0, p.wx & ~7, p.wy, mbar_w);
            tma_load_3d(smem + L.win_c + warp * kWinCBytes, tm + 1, p.cwx & ~7, p.cwy, 0, mbar_w);        // Cb and Cr: one box over the plane dimension
        }
    };
    if (WS && warp < n_tiles) issue_w(warp, 0);

    // ---- MC rounds (normally one: a CTU of >=16x16 CUs has at most 16 tiles) -----------------------------------------------
    const int n_rounds = (n_tiles + kTileCap - 1) / kTileCap;
    auto issue_r = [&](int round, int l) {
        // the windows of list l of this round's tiles.  Every warp issues the boxes of its own two tile slots (lanes 0, 1): a single issuing
        // warp spent ~2000 cycles on 48 serialised TMA instructions and the other seven waited for it at the next barrier (13 % of
        // all stall samples).  Warp 0 announces the byte count of the whole batch.
        const int t0 = round * kTileCap, nt = min(kTileCap, n_tiles - t0);
        if (warp == 0) {
            const bool any = lane < nt && l < s_tile[t0 + lane].nl;
            const int n = __popc(__ballot_sync(0xffffffffu, any));
            if (lane == 0) mbar_expect_tx(mbar, (uint32_t)n * 2 * (kBoxLW * kBoxLH + 2 * kBoxCW * kBoxCH));      // n == 0 completes the phase at once
        }
        const int slot = 2 * warp + lane;
        if (lane < 2 && slot < nt && l < s_tile[t0 + slot].nl) {
            const TilePred p = s_pred[(t0 + slot) * NL + l];
            const CUtensorMap *tm = a.ref_tmap[p.ref];
            tma_load_2d(smem + L.win_l + slot * kWinLBytes, tm + 0, p.wx & ~7, p.wy, mbar);
            tma_load_3d(smem + L.win_c + slot * kWinCBytes, tm + 1, p.cwx & ~7, p.cwy, 0, mbar);        // Cb and Cr: one box over the plane dimension
        }
    };
    if (!WS) issue_r(0, 0);

    // Block lookup for line li of a pass: start at the block that holds the first line of the warp's 32-line batch (table written with the
    // descriptors) and step over the block starts up to li - one step per block boundary inside the batch, none for blocks of 32+ lines.
    // (The previous warp-wide search over the prefix table cost 6 % of the kernel.)
    auto find_tu = [&](const uint16_t *pre, const uint16_t *bat, int li) {
        int b = bat[li >> 5];
        while (li >= (int)pre[b + 1]) b++;
        return b;
    };

    // ---- residual pass 1 (IDP.2A): luma lines first, then chroma lines (each padded to whole warps).  Baseline: rows; IQT: columns
    const uint16_t *preA = IQT ? s_pre2 : s_pre1, *preB = IQT ? s_pre1 : s_pre2;
    const uint16_t *batA = IQT ? s_bat2 : s_bat1, *batB = IQT ? s_bat1 : s_bat2;
    const int nAy = IQT ? n_l2y : n_l1y, nAc = IQT ? n_l2c : n_l1c, nBy = IQT ? n_l1y : n_l2y, nBc = IQT ? n_l1c : n_l2c;
    const int nAy_w = (nAy + 31) & ~31, nAc_w = (nAc + 31) & ~31;
    const int16_t *s_coef = (const int16_t *)(smem + L.coef);
    if (coef_bytes > 0) mbar_wait(mbar_coef, 0);
    for (int i0 = warp * 32; i0 < nAy_w + nAc_w; i0 += kR2Threads) {
        const bool chroma = i0 >= nAy_w;
        const int li0 = chroma ? i0 - nAy_w : i0, li = li0 + lane;
        if (li >= (chroma ? nAc : nAy)) continue;
        const int b = chroma ? find_tu(preA + 1, batA + kBatchCap, li) : find_tu(preA, batA, li);
        const TuDesc d = s_tu[b];
        const int q = li - (int)(chroma ? preA[b + 1] : preA[b]);           // task of the block: row / row pair (Baseline), column (IQT)
        const int ln = IQT ? d.lw_lh >> 4 : d.lw_lh & 15;
        const int off = d.shift ? (1 << (d.shift - 1)) : 0, cs = 1 << d.cstride_log2;
        const bool wide = (d.plane_wide & 4) != 0;
        if (IQT) {
            const int16_t *src = s_coef + d.coef_off + q;
            int *dst = s_tmp + d.tmp_off + q;
            switch (ln) {
            case 1: row_pass<2, true>(src, cs, dst, kTmpStride, d.mul, off, d.shift, wide); break;
            case 2: row_pass<4, true>(src, cs, dst, kTmpStride, d.mul, off, d.shift, wide); break;
            case 3: row_pass<8, true>(src, cs, dst, kTmpStride, d.mul, off, d.shift, wide); break;
            case 4: row_pass<16, true>(src, cs, dst, kTmpStride, d.mul, off, d.shift, wide); break;
            case 5: row_pass<32, true>(src, cs, dst, kTmpStride, d.mul, off, d.shift, wide); break;
            default: row_pass<64, true>(src, cs, dst, kTmpStride, d.mul, off, d.shift, wide); break;
            }
        } else {
            const int r0 = ln <= 4 ? 2 * q : q;
            const int16_t *src = s_coef + d.coef_off + (r0 << d.cstride_log2);
            int *dst = s_tmp + d.tmp_off + r0 * kTmpStride;
            switch (ln) {
            case 1: row_pass2<2>(src, cs, dst, kTmpStride, d.mul, off, d.shift, wide); break;
            case 2: row_pass2<4>(src, cs, dst, kTmpStride, d.mul, off, d.shift, wide); break;
            case 3: row_pass2<8>(src, cs, dst, kTmpStride, d.mul, off, d.shift, wide); break;
            case 4: row_pass2<16>(src, cs, dst, kTmpStride, d.mul, off, d.shift, wide); break;
            case 5: row_pass<32, false>(src, cs, dst, kTmpStride, d.mul, off, d.shift, wide); break;
            default: row_pass<64, false>(src, cs, dst, kTmpStride, d.mul, off, d.shift, wide); break;
            }
        }
    }
    __syncthreads();
    if (n_tu < 3 * ncu) {         // the coefficient slice is consumed: uncoded blocks must read as zero residual
        int4 *z = (int4 *)s_res;
        const int nz = 2 * kPlaneRows * kResStride / 16;
        for (int i = tid; i < nz; i += kR2Threads) z[i] = make_int4(0, 0, 0, 0);
        __syncthreads();
    }
    // ---- residual pass 2 (IMAD): the other direction ---------------------------------------------------------------------------
    {
        const int sh2 = (IQT ? 12 : 19) - (a.bd_l - 8);
        const int nBy_w = (nBy + 31) & ~31, nBc_w = (nBc + 31) & ~31;
        for (int i0 = warp * 32; i0 < nBy_w + nBc_w; i0 += kR2Threads) {
            const bool chroma = i0 >= nBy_w;
            const int li0 = chroma ? i0 - nBy_w : i0, li = li0 + lane;
            if (li >= (chroma ? nBc : nBy)) continue;
            const int b = chroma ? find_tu(preB + 1, batB + kBatchCap, li) : find_tu(preB, batB, li);
            const TuDesc d = s_tu[b];
            const int q = li - (int)(chroma ? preB[b + 1] : preB[b]);       // task of the block: column / column pair (Baseline), row (IQT)
            const int ln = IQT ? d.lw_lh & 15 : d.lw_lh >> 4;
            if (IQT) {
                const int *src = s_tmp + d.tmp_off + q * kTmpStride;
                int16_t *dst = s_res + d.res_off + q * kResStride;
                switch (ln) {
                case 1: col_pass<2, true>(src, kTmpStride, dst, kResStride, sh2); break;
                case 2: col_pass<4, true>(src, kTmpStride, dst, kResStride, sh2); break;
                case 3: col_pass<8, true>(src, kTmpStride, dst, kResStride, sh2); break;
                case 4: col_pass<16, true>(src, kTmpStride, dst, kResStride, sh2); break;
                case 5: col_pass<32, true>(src, kTmpStride, dst, kResStride, sh2); break;
                default: col_pass<64, true>(src, kTmpStride, dst, kResStride, sh2); break;
                }
            } else {
                const int c0 = ln <= 4 ? 2 * q : q;
                const int *src = s_tmp + d.tmp_off + c0;
                int16_t *dst = s_res + d.res_off + c0;
                switch (ln) {
                case 1: col_pass2<2>(src, kTmpStride, dst, kResStride, sh2); break;
                case 2: col_pass2<4>(src, kTmpStride, dst, kResStride, sh2); break;
                case 3: col_pass2<8>(src, kTmpStride, dst, kResStride, sh2); break;
                case 4: col_pass2<16>(src, kTmpStride, dst, kResStride, sh2); break;
                case 5: col_pass<32, false>(src, kTmpStride, dst, kResStride, sh2); break;
                default: col_pass<64, false>(src, kTmpStride, dst, kResStride, sh2); break;
                }
            }
        }
    }
    __syncthreads();              // residual complete; pass-1 buffer is free and becomes the vertical-pair buffers

    // ---- park the residual of intra / IBC CUs in the picture (s16 fits a pel; the wavefront kernel replaces it by the reconstruction)
    for (int i = 0; i < ncu && cnt[7]; i++) {
        const XB200_CU cu = s_cu[i];
        if (!xb_wavefront_mode(cu.mode) || (DISP && (cu.flags & kCuOtherKernel))) continue;                  // uniform
        const int w = 1 << cu.log2w, h = 1 << cu.log2h, lx = cu.x - ctu_x, ly = cu.y - ctu_y;
        for (int k = tid; k < (w * h) >> 1; k += kR2Threads) {      // luma, two samples per thread
            const int y = k >> (cu.log2w - 1), x = (k & ((w >> 1) - 1)) << 1;
            *(int *)(a.cur.y + (size_t)(cu.y + y) * a.s_l + cu.x + x) = *(const int *)(s_res + (ly + y) * kResStride + lx + x);
        }
        for (int k = tid; k < (w * h) >> 2; k += kR2Threads) {      // chroma: both planes, two samples per thread
            const int pl = k >= ((w * h) >> 3), kk = k - pl * ((w * h) >> 3);
            const int y = kk >> (cu.log2w - 2), x = (kk & ((w >> 2) - 1)) << 1;
            *(int *)((pl ? a.cur.v : a.cur.u) + (size_t)((cu.y >> 1) + y) * a.s_c + (cu.x >> 1) + x) =
                *(const int *)(s_res + plane_origin(1 + pl, kResStride) + ((ly >> 1) + y) * kResStride + (lx >> 1) + x);
        }
    }

    int *s_m2l = s_tmp;                                        // [slots][kM2LWords] (chroma needs no pair buffer: its two stages are one pass)
    const int maxv2 = ((1 << a.bd_l) - 1) * 0x00010001;        // the reference clips all planes with the luma depth
    const int maxc2 = ((1 << a.bd_c) - 1) * 0x00010001;
    const int s1l = min(4, a.bd_l - 8), s2l = max(8, 20 - a.bd_l);
    const int s1c = min(4, a.bd_c - 8), s2c = max(8, 20 - a.bd_c);

    int phase = 0;
    if constexpr (WS) {
#pragma unroll 1
    for (int tile = warp; tile < n_tiles; tile += kTileCap) {
        const TileDesc td = s_tile[tile];
        const int slot = warp;
        // the running prediction of this lane's samples: luma 2 columns x 4 rows, chroma 2 columns x 2 rows (packed pairs)
        int outp[4], outc[2];
#pragma unroll 1
        for (int l = 0; l < NL; l++) {
            if (l >= td.nl) break;                                  // (uniform: one tile per warp)
            mbar_wait(mbar_w, phase & 1);
            phase++;
            const TilePred p = s_pred[tile * NL + l];
            // ---- luma, horizontal stage, output = vertical pairs: 24 tasks (2 column halves x 12 row-pairs)
            if (lane < 24) {
                const int half = lane >= 12 ? 1 : 0, rp = lane - 12 * half;
                if (half * 8 < td.tw && 2 * rp < td.th + 7) {
                    const int offx = p.offs & 7, par = offx & 1;
                    const int *win = (const int *)(smem + L.win_l + slot * kWinLBytes) + (offx >> 1) + half * 4;
                    const Taps5 te = ld_taps5(p.phx, par), to = ld_taps5(p.phx, par + 1);   // even outputs: A0 | B, odd outputs: B | A1
                    const int sh = p.two_d ? s1l : 6;
                    int hv[2][8];
#pragma unroll
                    for (int rr = 0; rr < 2; rr++) {
                        const int *rowp = win + (2 * rp + rr) * kWinLStrideW;
                        int q[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) q[j] = rowp[j];
#pragma unroll
                        for (int o = 0; o < 4; o++) {
                            hv[rr][2 * o] = fir5(te, q[o], q[o + 1], q[o + 2], q[o + 3], q[o + 4 < 8 ? o + 4 : 7], 0) >> sh;
                            hv[rr][2 * o + 1] = fir5(to, q[o], q[o + 1], q[o + 2], q[o + 3], q[o + 4 < 8 ? o + 4 : 7], 0) >> sh;
                        }
                    }
                    int4 *dst = (int4 *)(s_m2l + slot * kM2LWords + half * 8 + rp * kM2LStrideW);
                    dst[0] = make_int4(pack16(hv[0][0], hv[1][0]), pack16(hv[0][1], hv[1][1]), pack16(hv[0][2], hv[1][2]), pack16(hv[0][3], hv[1][3]));
                    dst[1] = make_int4(pack16(hv[0][4], hv[1][4]), pack16(hv[0][5], hv[1][5]), pack16(hv[0][6], hv[1][6]), pack16(hv[0][7], hv[1][7]));
                }
            }
            // chroma, both stages in one pass of all 32 lanes (2 planes x 4 column pairs x 4 row groups, 2 columns x 2 rows each): a lane
            // filters its 5 rows horizontally straight from the window and vertically from registers.  (As a separate horizontal pass the
            // 12 row-pair tasks of the two planes kept 12 of 32 lanes busy and went through shared memory once more.)
            {
                const int pl = lane >> 4, cp = lane & 3, rg = (lane >> 2) & 3;
                const int cw = td.tw >> 1, ch = td.th >> 1;
                if (2 * cp < cw && 2 * rg < ch) {
                    const int offx = p.offs >> 4, par = offx & 1;
                    const int *win = (const int *)(smem + L.win_c + slot * kWinCBytes + pl * kWinCPlane) + (offx >> 1) + cp + (2 * rg) * kWinCStrideW;
                    const Taps3 he = ld_taps3(p.cphx, par), ho = ld_taps3(p.cphx, par + 1);
                    const int sh1 = p.ctwo_d ? s1c : 6;
                    int h0[6], h1[6];
#pragma unroll
                    for (int r = 0; r < 5; r++) {
                        const int q0 = win[r * kWinCStrideW], q1 = win[r * kWinCStrideW + 1], q2 = win[r * kWinCStrideW + 2];
                        h0[r] = fir3(he, q0, q1, q2, 0) >> sh1;
                        h1[r] = fir3(ho, q0, q1, q2, 0) >> sh1;
                    }
                    h0[5] = h1[5] = 0;              // weight 0 in the odd-row tap set
                    int P[3][2];
#pragma unroll
                    for (int j = 0; j < 3; j++) { P[j][0] = pack16(h0[2 * j], h0[2 * j + 1]); P[j][1] = pack16(h1[2 * j], h1[2 * j + 1]); }
                    const Taps3 te = ld_taps3(p.cphy, 0), to = ld_taps3(p.cphy, 1);
                    const int sh = p.ctwo_d ? s2c : 6, rnd = p.ctwo_d ? (1 << (s2c - 1)) : 0;
                    int e0 = fir3(te, P[0][0], P[1][0], 0, rnd) >> sh;
                    int e1 = fir3(te, P[0][1], P[1][1], 0, rnd) >> sh;
                    int o0 = fir3(to, P[0][0], P[1][0], P[2][0], rnd) >> sh;
                    int o1 = fir3(to, P[0][1], P[1][1], P[2][1], rnd) >> sh;
                    int pe = __vimin_s16x2_relu(pack16(e0, e1), maxc2);
                    int po = __vimin_s16x2_relu(pack16(o0, o1), maxc2);
                    if (l == 0) { outc[0] = pe; outc[1] = po; }
                    else {
                        outc[0] = ((outc[0] + pe + 0x00010001) >> 1) & 0x7fff7fff;
                        outc[1] = ((outc[1] + po + 0x00010001) >> 1) & 0x7fff7fff;
                    }
                }
            }
            __syncwarp();
            // the warp is done with its windows: the next boxes (other list of this tile, or the warp's next tile) load while the vertical
            // stage of this one runs
            if (l + 1 < td.nl) issue_w(tile, l + 1);
            else if (tile + kTileCap < n_tiles) issue_w(tile + kTileCap, 0);

            // ---- vertical stage: this list's prediction into the running registers (xevd_average_16b_no_clip for the second list) -----
            // luma: one lane = 2 columns x 4 rows (8 column pairs x 4 row groups)
            {
                const int cp = lane & 7, rg = lane >> 3;
                if (2 * cp < td.tw && 4 * rg < td.th) {
                    const Taps5 te = ld_taps5(p.phy, 0), to = ld_taps5(p.phy, 1);
                    const int sh = p.two_d ? s2l : 6, rnd = p.two_d ? (1 << (s2l - 1)) : 0;
                    const int *m2 = s_m2l + slot * kM2LWords + (2 * rg) * kM2LStrideW + 2 * cp;
                    int P[6][2];
#pragma unroll
                    for (int j = 0; j < 6; j++) { const int2 v = *(const int2 *)(m2 + j * kM2LStrideW); P[j][0] = v.x; P[j][1] = v.y; }
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        int e0 = fir5(te, P[q][0], P[q + 1][0], P[q + 2][0], P[q + 3][0], 0, rnd) >> sh;
                        int e1 = fir5(te, P[q][1], P[q + 1][1], P[q + 2][1], P[q + 3][1], 0, rnd) >> sh;
                        int o0 = fir5(to, P[q][0], P[q + 1][0], P[q + 2][0], P[q + 3][0], P[q + 4][0], rnd) >> sh;
                        int o1 = fir5(to, P[q][1], P[q + 1][1], P[q + 2][1], P[q + 3][1], P[q + 4][1], rnd) >> sh;
                        int pe = __vimin_s16x2_relu(pack16(e0, e1), maxv2);
                        int po = __vimin_s16x2_relu(pack16(o0, o1), maxv2);
                        if (l == 0) { outp[2 * q] = pe; outp[2 * q + 1] = po; }
                        else {      // two clipped, non-negative predictions
                            outp[2 * q] = ((outp[2 * q] + pe + 0x00010001) >> 1) & 0x7fff7fff;
                            outp[2 * q + 1] = ((outp[2 * q + 1] + po + 0x00010001) >> 1) & 0x7fff7fff;
                        }
                    }
                }
            }
            __syncwarp();           // the pair buffers of this warp's slot are rewritten by the next list / tile
        }

        // ---- reconstruction: prediction + residual, clip, store -----------------------------------------------------------------------
        {
            const int cp = lane & 7, rg = lane >> 3;
            if (2 * cp < td.tw && 4 * rg < td.th) {
                const int x = td.px + 2 * cp, y = td.py + 4 * rg;
                const int *res = (const int *)(s_res + y * kResStride + x);
                pel *dst = a.cur.y + (size_t)(ctu_y + y) * a.s_l + ctu_x + x;
#pragma unroll
                for (int r = 0; r < 4; r++)
                    if (4 * rg + r < td.th) {
                        const int v = (int)__viaddmin_s16x2_relu(outp[r], res[r * (kResStride / 2)], maxv2);
                        if (PEER) *(int *)(s_out + (y + r) * 64 + x) = v;
                        else *(int *)(dst + (size_t)r * a.s_l) = v;
                    }
            }
            const int pl = lane >> 4, ccp = lane & 3, crg = (lane >> 2) & 3;
            const int cw = td.tw >> 1, ch = td.th >> 1;
            if (2 * ccp < cw && 2 * crg < ch) {
                const int x = (td.px >> 1) + 2 * ccp, y = (td.py >> 1) + 2 * crg;
                const int *res = (const int *)(s_res + plane_origin(1 + pl, kResStride) + y * kResStride + x);
                pel *dst = (pl ? a.cur.v : a.cur.u) + (size_t)((ctu_y >> 1) + y) * a.s_c + (ctu_x >> 1) + x;
#pragma unroll
                for (int r = 0; r < 2; r++)
                    if (2 * crg + r < ch) {
                        const int v = (int)__viaddmin_s16x2_relu(outc[r], res[r * (kResStride / 2)], maxv2);
                        if (PEER) *(int *)(s_out + 64 * 64 + pl * 32 * 32 + (y + r) * 32 + x) = v;
                        else *(int *)(dst + (size_t)r * a.s_c) = v;
                    }
            }
        }
    }

    } else {
    for (int round = 0; round < n_rounds; round++) {
        const int t0 = round * kTileCap, nt = min(kTileCap, n_tiles - t0);
        // the running prediction of this thread's samples: luma 2 columns x 8 rows, chroma 2 columns x 4 rows (packed pairs)
        int outp[8], outc[4];
#pragma unroll 1
        for (int l = 0; l < NL; l++) {
            mbar_wait(mbar, phase & 1);
            phase++;
            // ---- horizontal stage (warp-local: warp w owns tile slots 2w, 2w+1 through both stages), output = vertical pairs ------
            // luma: 24 tasks per slot (2 column halves x 12 row-pairs), chroma: 12 per slot (2 planes x 6 row-pairs)
#pragma unroll 1
            for (int it = 0; it < 2; it++) {
                const int task = it * 32 + lane;
                if (task >= 48) continue;
                const int sidx = task >= 24 ? 1 : 0, k = task - 24 * sidx, half = k >= 12 ? 1 : 0, rp = k - 12 * half;
                const int slot = 2 * warp + sidx;
                if (slot >= nt) continue;
                const TileDesc td = s_tile[t0 + slot];
                if (l >= td.nl || half * 8 >= td.tw || 2 * rp >= td.th + 7) continue;
                const TilePred p = s_pred[(t0 + slot) * NL + l];
                const int offx = p.offs & 7, par = offx & 1;
                const int *win = (const int *)(smem + L.win_l + slot * kWinLBytes) + (offx >> 1) + half * 4;
                const Taps5 te = ld_taps5(p.phx, par), to = ld_taps5(p.phx, par + 1);   // even outputs: A0 | B, odd outputs: B | A1
                const int sh = p.two_d ? s1l : 6;
                int hv[2][8];
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const int *rowp = win + (2 * rp + rr) * kWinLStrideW;
                    int q[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) q[j] = rowp[j];
#pragma unroll
                    for (int o = 0; o < 4; o++) {
                        hv[rr][2 * o] = fir5(te, q[o], q[o + 1], q[o + 2], q[o + 3], q[o + 4 < 8 ? o + 4 : 7], 0) >> sh;
                        hv[rr][2 * o + 1] = fir5(to, q[o], q[o + 1], q[o + 2], q[o + 3], q[o + 4 < 8 ? o + 4 : 7], 0) >> sh;
                    }
                }
                int4 *dst = (int4 *)(s_m2l + slot * kM2LWords + half * 8 + rp * kM2LStrideW);
                dst[0] = make_int4(pack16(hv[0][0], hv[1][0]), pack16(hv[0][1], hv[1][1]), pack16(hv[0][2], hv[1][2]), pack16(hv[0][3], hv[1][3]));
                dst[1] = make_int4(pack16(hv[0][4], hv[1][4]), pack16(hv[0][5], hv[1][5]), pack16(hv[0][6], hv[1][6]), pack16(hv[0][7], hv[1][7]));
            }
            // chroma, both stages in one pass (warp w's 32 threads are the vertical-stage threads of its own two slots): one thread =
            // 2 columns x 4 rows of one plane; it filters its 7 rows horizontally straight from the window and vertically from registers
            {
                const int slot = tid >> 4, kk = tid & 15;
                const int pl = kk >> 3, cp = kk & 3, rg = (kk >> 2) & 1;
                if (slot < nt) {
                    const TileDesc td = s_tile[t0 + slot];
                    const int cw = td.tw >> 1, ch = td.th >> 1;
                    if (2 * cp < cw && 4 * rg < ch && l < td.nl) {
                        const TilePred p = s_pred[(t0 + slot) * NL + l];
                        const int offx = p.offs >> 4, par = offx & 1;
                        const int *win = (const int *)(smem + L.win_c + slot * kWinCBytes + pl * kWinCPlane) + (offx >> 1) + cp + (4 * rg) * kWinCStrideW;
                        const Taps3 he = ld_taps3(p.cphx, par), ho = ld_taps3(p.cphx, par + 1);
                        const int sh1 = p.ctwo_d ? s1c : 6;
                        int h0[8], h1[8];
#pragma unroll
                        for (int r = 0; r < 7; r++) {
                            const int q0 = win[r * kWinCStrideW], q1 = win[r * kWinCStrideW + 1], q2 = win[r * kWinCStrideW + 2];
                            h0[r] = fir3(he, q0, q1, q2, 0) >> sh1;
                            h1[r] = fir3(ho, q0, q1, q2, 0) >> sh1;
                        }
                        h0[7] = h1[7] = 0;              // weight 0 in the odd-row tap set
                        int P[4][2];
#pragma unroll
                        for (int j = 0; j < 4; j++) { P[j][0] = pack16(h0[2 * j], h0[2 * j + 1]); P[j][1] = pack16(h1[2 * j], h1[2 * j + 1]); }
                        const Taps3 te = ld_taps3(p.cphy, 0), to = ld_taps3(p.cphy, 1);
                        const int sh = p.ctwo_d ? s2c : 6, rnd = p.ctwo_d ? (1 << (s2c - 1)) : 0;
#pragma unroll
                        for (int q = 0; q < 2; q++) {
                            int e0 = fir3(te, P[q][0], P[q + 1][0], 0, rnd) >> sh;
                            int e1 = fir3(te, P[q][1], P[q + 1][1], 0, rnd) >> sh;
                            int o0 = fir3(to, P[q][0], P[q + 1][0], P[q + 2][0], rnd) >> sh;
                            int o1 = fir3(to, P[q][1], P[q + 1][1], P[q + 2][1], rnd) >> sh;
                            int pe = __vimin_s16x2_relu(pack16(e0, e1), maxc2);
                            int po = __vimin_s16x2_relu(pack16(o0, o1), maxc2);
                            if (l == 0) { outc[2 * q] = pe; outc[2 * q + 1] = po; }
                            else {
                                outc[2 * q] = ((outc[2 * q] + pe + 0x00010001) >> 1) & 0x7fff7fff;
                                outc[2 * q + 1] = ((outc[2 * q + 1] + po + 0x00010001) >> 1) & 0x7fff7fff;
                            }
                        }
                    }
                }
            }
            // every warp is done with the windows: the next batch (other list of this round, or the next round) loads while the vertical
            // stage of this one runs
            const bool more = l + 1 < NL || round + 1 < n_rounds;
            if (more) { __syncthreads(); if (l + 1 < NL) issue_r(round, l + 1); else issue_r(round + 1, 0); }
            else __syncwarp();

            // ---- vertical stage: this list's prediction into the running registers (xevd_average_16b_no_clip for the second list) -----
            // luma: one thread = 2 columns x 8 rows, 16 threads per slot (8 column pairs x 2 row groups)
            {
                const int slot = tid >> 4, k = tid & 15;
                const int cp = k & 7, rg = k >> 3;
                if (slot < nt) {
                    const TileDesc td = s_tile[t0 + slot];
                    if (2 * cp < td.tw && 8 * rg < td.th && l < td.nl) {
                        const TilePred p = s_pred[(t0 + slot) * NL + l];
                        const Taps5 te = ld_taps5(p.phy, 0), to = ld_taps5(p.phy, 1);
                        const int sh = p.two_d ? s2l : 6, rnd = p.two_d ? (1 << (s2l - 1)) : 0;
                        const int *m2 = s_m2l + slot * kM2LWords + (4 * rg) * kM2LStrideW + 2 * cp;
                        int P[8][2];
#pragma unroll
                        for (int j = 0; j < 8; j++) { const int2 v = *(const int2 *)(m2 + j * kM2LStrideW); P[j][0] = v.x; P[j][1] = v.y; }
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            int e0 = fir5(te, P[q][0], P[q + 1][0], P[q + 2][0], P[q + 3][0], 0, rnd) >> sh;
                            int e1 = fir5(te, P[q][1], P[q + 1][1], P[q + 2][1], P[q + 3][1], 0, rnd) >> sh;
                            int o0 = fir5(to, P[q][0], P[q + 1][0], P[q + 2][0], P[q + 3][0], P[q + 4][0], rnd) >> sh;
                            int o1 = fir5(to, P[q][1], P[q + 1][1], P[q + 2][1], P[q + 3][1], P[q + 4][1], rnd) >> sh;
                            int pe = __vimin_s16x2_relu(pack16(e0, e1), maxv2);
                            int po = __vimin_s16x2_relu(pack16(o0, o1), maxv2);
                            if (l == 0) { outp[2 * q] = pe; outp[2 * q + 1] = po; }
                            else {      // two clipped, non-negative predictions
                                outp[2 * q] = ((outp[2 * q] + pe + 0x00010001) >> 1) & 0x7fff7fff;
                                outp[2 * q + 1] = ((outp[2 * q + 1] + po + 0x00010001) >> 1) & 0x7fff7fff;
                            }
                        }
                    }
                }
            }
            __syncwarp();           // the pair buffers of this warp's tiles are rewritten by the next list / round
        }

        // ---- reconstruction: prediction + residual, clip, store -----------------------------------------------------------------------
        {
            const int slot = tid >> 4, k = tid & 15;
            const int cp = k & 7, rg = k >> 3;
            if (slot < nt) {
                const TileDesc td = s_tile[t0 + slot];
                if (2 * cp < td.tw && 8 * rg < td.th) {
                    const int x = td.px + 2 * cp, y = td.py + 8 * rg;
                    const int *res = (const int *)(s_res + y * kResStride + x);
                    pel *dst = a.cur.y + (size_t)(ctu_y + y) * a.s_l + ctu_x + x;
#pragma unroll
                    for (int r = 0; r < 8; r++)
                        if (8 * rg + r < td.th) {
                            const int v = (int)__viaddmin_s16x2_relu(outp[r], res[r * (kResStride / 2)], maxv2);
                            if (PEER) *(int *)(s_out + (y + r) * 64 + x) = v;
                            else *(int *)(dst + (size_t)r * a.s_l) = v;
                        }
                }
                const int pl = k >> 3, ccp = k & 3, crg = (k >> 2) & 1;
                const int cw = td.tw >> 1, ch = td.th >> 1;
                if (2 * ccp < cw && 4 * crg < ch) {
                    const int x = (td.px >> 1) + 2 * ccp, y = (td.py >> 1) + 4 * crg;
                    const int *res = (const int *)(s_res + plane_origin(1 + pl, kResStride) + y * kResStride + x);
                    pel *dst = (pl ? a.cur.v : a.cur.u) + (size_t)((ctu_y >> 1) + y) * a.s_c + (ctu_x >> 1) + x;
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        if (4 * crg + r < ch) {
                            const int v = (int)__viaddmin_s16x2_relu(outc[r], res[r * (kResStride / 2)], maxv2);
                            if (PEER) *(int *)(s_out + 64 * 64 + pl * 32 * 32 + (y + r) * 32 + x) = v;
                            else *(int *)(dst + (size_t)r * a.s_c) = v;
                        }
                }
            }
        }
    }

    }

    if (PEER) {
        // copy-out: 16 bytes per thread, eight threads cover one 128-byte luma row of the CTU; same words to every GPU
        __syncthreads();
        const int rows_l = min(64, a.h - ctu_y), cols_l = min(64, a.w - ctu_x);
        for (int i = tid; i < 64 * 8; i += kR2Threads) {
            const int r = i >> 3, c8 = (i & 7) << 3;
            if (r < rows_l && c8 < cols_l)
                xb_store_all(a, (int4 *)(a.cur.y + (size_t)(ctu_y + r) * a.s_l + ctu_x + c8), *(const int4 *)(s_out + r * 64 + c8));
        }
        for (int i = tid; i < 2 * 32 * 4; i += kR2Threads) {
            const int pl = i >> 7, r = (i >> 2) & 31, c8 = (i & 3) << 3;
            if (r < (rows_l >> 1) && c8 < (cols_l >> 1))
                xb_store_all(a, (int4 *)((pl ? a.cur.v : a.cur.u) + (size_t)((ctu_y >> 1) + r) * a.s_c + (ctu_x >> 1) + c8),
                             *(const int4 *)(s_out + 64 * 64 + pl * 32 * 32 + r * 32 + c8));
        }
    }

    // ---- publish per-SCU maps (xevd_set_dec_info) ------------------------------------------------------------------------------
    // PEER: entries are collected per CTU (16 x 16 SCUs) and written as 16-byte words, because narrow stores to peer memory cost one
    // NVLink request each (measured: they doubled the kernel time, profiles/r1)
    int2 *sm_mv = (int2 *)(s_out + 64 * 64 + 2 * 32 * 32);
    uint32_t *sm_scu = (uint32_t *)(sm_mv + 256);
    int16_t *sm_refi = (int16_t *)(sm_scu + 256);
    uint8_t *sm_edge = (uint8_t *)(sm_refi + 256);
    const bool wide_maps = PEER && (a.w_scu & 15) == 0;
    // A group of lanes per CU, one SCU per lane and step: 16 groups of 16 lanes, or fewer and wider groups when the CTU has few (hence
    // large) CUs - a 64x64 CU is written by all 256 threads at once.  (A thread per CU wrote up to 256 entries x 5 maps on its own at the
    // kernel's tail: 4K pictures of 64x64 CUs took 316 us instead of ~125, of 16x16 CUs 86 instead of 76.)
    const int lgg = ncu > 8 ? 4 : (ncu > 4 ? 3 : (ncu > 2 ? 2 : (ncu > 1 ? 1 : 0))), glanes = kR2Threads >> lgg;
    for (int i = tid >> (8 - lgg); i < ncu; i += 1 << lgg) {
        const XB200_CU cu = s_cu[i];
        if (DISP && (cu.flags & kCuOtherKernel)) continue;
        const int sx = cu.x >> 2, sy = cu.y >> 2, lnw = cu.log2w - 2, nw = 1 << lnw, nscu_cu = 1 << (cu.log2w + cu.log2h - 4);
        const bool intra = cu.mode == XB200_MODE_INTRA, ibc = cu.mode == XB200_MODE_IBC;
        uint32_t m = ((uint32_t)(cu.qp_map & 0x7f) << 16) | (1u << 31) | (intra ? 1u << 15 : 0u) | (ibc ? 1u << 26 : 0u);
        if (cu.cbf & 1) m |= 1u << 24;
        if (cu.flags & XB200_CUF_SKIP) m |= 1u << 23;
        const int2 mv = intra ? make_int2(0, 0) : make_int2(((const int *)cu.mv)[0], ((const int *)cu.mv)[1]);
        const int16_t rf = (intra || ibc) ? (int16_t)-1 : *(const int16_t *)cu.refi;
        for (int q = tid & (glanes - 1); q < nscu_cu; q += glanes) {
                const int y = q >> lnw, x = q & (nw - 1);
                const uint8_t e = (uint8_t)(((x & 15) == 0 ? XB200_EDGE_LEFT : 0) | ((y & 15) == 0 ? XB200_EDGE_TOP : 0));
                if (wide_maps) {
                    const int qq = ((sy + y) & 15) * 16 + ((sx + x) & 15);
                    sm_mv[qq] = mv; sm_scu[qq] = m; sm_refi[qq] = rf; sm_edge[qq] = e;
                    continue;
                }
                const int p = (sy + y) * a.w_scu + sx + x;
                const bool fo = a.peer_maps != 0;
                xb_store_all(a, a.map_scu + p, m, fo);
                xb_store_all(a, (int2 *)a.map_mv + p, mv, fo);
                xb_store_all(a, (int2 *)a.map_unrefined_mv + p, mv, fo);
                xb_store_all(a, (int16_t *)a.map_refi + p, rf, fo);
                xb_store_all(a, a.map_edge + p, e, fo);
                if (a.map_order) a.map_order[p] = (uint16_t)i;
            }
    }
    if (wide_maps) {
        __syncthreads();
        const int rows = min(16, a.h_scu - (ctu_y >> 2));
        const int p0 = (ctu_y >> 2) * a.w_scu + (ctu_x >> 2);              // the CTU's 16 SCU columns all exist when w_scu % 16 == 0
        for (int i = tid; i < 16 * 23; i += kR2Threads) {
            const int r = i / 23, k = i - r * 23;
            if (r >= rows) continue;
            const int p = p0 + r * a.w_scu;
            if (k < 8) { const int4 v = ((const int4 *)(sm_mv + r * 16))[k]; xb_store_all(a, (int4 *)((int2 *)a.map_mv + p) + k, v); xb_store_all(a, (int4 *)((int2 *)a.map_unrefined_mv + p) + k, v); }
            else if (k < 16) continue;          // (slots 8..15: the unrefined map shares the store above)
            else if (k < 20) xb_store_all(a, (int4 *)(a.map_scu + p) + (k - 16), ((const int4 *)(sm_scu + r * 16))[k - 16]);
            else if (k < 22) xb_store_all(a, (int4 *)((int16_t *)a.map_refi + p) + (k - 20), ((const int4 *)(sm_refi + r * 16))[k - 20]);
            else xb_store_all(a, (int4 *)(a.map_edge + p), *(const int4 *)(sm_edge + r * 16));
        }
    }
}

}  // namespace xb
