// xb_filters.cuh -- picture-wide passes after CU reconstruction: border padding (and, below, deblocking).
#pragma once
#include "xb_common.cuh"

namespace xb {

// ---- border padding: xevd_picbuf_expand -> picbuf_expand (src_base/xevd_util.c:365-427) ---------------------------
// One CTA per padded row of each plane.  Rows inside the picture only write their 2*pad border samples, rows of
// the top / bottom band write a full replicated row.  4 samples (8 bytes) per thread access; w, pad are
// multiples of 4 for every plane (w % 8 == 0, pads 144 / 72).
struct PadPlane { pel *org; int stride, w, h, pad; };
struct PadArgs { PadPlane pl[3]; int row_start[4]; };

__global__ void __launch_bounds__(128) k_pad(const __grid_constant__ PadArgs a)
{
    const int row = blockIdx.x;
    const int p = row >= a.row_start[2] ? 2 : (row >= a.row_start[1] ? 1 : 0);
    const PadPlane P = a.pl[p];
    const int y = row - a.row_start[p] - P.pad;               // row relative to the picture
    const int ys = min(max(y, 0), P.h - 1);
    xb_grid_wait();
    xb_grid_release();
    const pel *src = P.org + (size_t)ys * P.stride;
    pel *dst = P.org + (size_t)y * P.stride;
    const int padv = P.pad >> 2, wv = P.w >> 2;
    const short4 lft = make_short4(src[0], src[0], src[0], src[0]);
    const short4 rgt = make_short4(src[P.w - 1], src[P.w - 1], src[P.w - 1], src[P.w - 1]);
    for (int v = threadIdx.x; v < padv; v += blockDim.x) {
        ((short4 *)(dst - P.pad))[v] = lft;
        ((short4 *)(dst + P.w))[v] = rgt;
    }
    if (y != ys)
        for (int v = threadIdx.x; v < wv; v += blockDim.x) ((short4 *)dst)[v] = ((const short4 *)src)[v];
}

inline void launch_pad(pel *y, int s_l, int w, int h, int pad_l, pel *u, pel *v, int s_c, int w_c, int h_c, int pad_c, cudaStream_t st)
{
    PadArgs a;
    a.pl[0] = {y, s_l, w, h, pad_l};
    a.pl[1] = {u, s_c, w_c, h_c, pad_c};
    a.pl[2] = {v, s_c, w_c, h_c, pad_c};
    a.row_start[0] = 0;
    a.row_start[1] = h + 2 * pad_l;
    a.row_start[2] = a.row_start[1] + h_c + 2 * pad_c;
    a.row_start[3] = a.row_start[2] + h_c + 2 * pad_c;
    xb_launch_early(k_pad, dim3(a.row_start[3]), dim3(128), 0, st, a);
}

}  // namespace xb

// ---- deblocking, Baseline filter (sps->tool_addb == 0) --------------------------------------------------------------------
// Replaces xevd_deblock / xevdm_deblock + deblock_tree + xevd_deblock_cu_ver/hor + deblock_scu_* (src_base/xevd.c:1057-1243,
// src_base/xevd_df.c:34-545; Main library: src_main/xevdm.c:1935-2103, src_main/xevdm_df.c:38-360) for a picture that is one
// tile and one slice.  Two picture-wide launches: all vertical edges, then all horizontal edges (xevd.c:1918-1975).
// One thread per 4-sample edge segment (SCU granularity).  Luma segments touch disjoint samples, so their order is free.
// Chroma (4:2:0) segments of 4-wide / 4-high CUs are only 2 samples apart and each reads a sample the previous one wrote;
// the reference visits them left-to-right / top-to-bottom (decoding order), so the first segment of such a run walks the
// whole run sequentially.
namespace xb {

__constant__ uint8_t c_df_st[4 * 52] = {   // xevd_tbl_df_st (src_base/xevd_tbl.c:306-324), rows: intra | luma cbf | motion | none
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 12, 12, 12, 12, 12,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 11, 11, 11, 11, 11,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 4, 5, 6, 7, 8, 9, 10, 10, 10, 10, 10,
    0};

struct DbkArgs {
    pel *y, *u, *v;
    int s_l, s_c, w, h, w_scu, h_scu;
    int bd_l, bd_c, qp_u_offset, qp_v_offset;
    const uint32_t *map_scu;
    const int16_t *map_mv;           // Baseline filter: ctx->map_mv (refined / affine sub-block vectors)
    const int16_t *map_umv;          // ADDB: vectors before DMVR refinement, read where map_scu carries the DMVR flag; map_mv elsewhere
    const int8_t *map_refi;
    const uint8_t *map_edge;
    const uint16_t *map_order;   // decoding order of the chroma-owning CU inside its CTU (tool_suco && !tool_addb), else null
    int8_t cq[2][58];            // chroma QP mapping for qp >= 0 (xevd_qp_chroma_dynamic); identity below 0
    // Main-profile filter (tool_addb)
    int alpha_offset, beta_offset, log2_ctu;
    int8_t ref_id[2][XB_MAX_REFS];   // identity of the PICTURE behind (list, refi): get_bs compares pictures, not indices
};

__device__ __forceinline__ int dbk_st(int cls, int q)
{
    const int f = cls * 52 + q;
    return (q < 0 || f >= 4 * 52) ? 0 : c_df_st[f];
}

// Loads that are issued before the first test of a thread: a segment's work is a chain edge flag -> maps of both SCUs -> samples -> stores,
// three dependent memory round trips (long_scoreboard 20.7 per issue, 19 % of the HBM peak in profiles/r1/deblock_ncu_summary.txt).  The
// deblocking kernels fetch all of it at once, for every thread; `volatile` keeps the compiler from sinking a load below the test of its use.
// (Generic addresses of cudaMalloc'ed memory are global-window addresses: no cvta, which cost 5.7 % of the instructions of a pass.)
__device__ __forceinline__ int ld_now32(const void *p)
{
    int v;
    asm volatile("ld.global.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int2 ld_now64(const void *p)
{
    int2 v;
    asm volatile("ld.global.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_now16(const void *p)
{
    unsigned short v;
    asm volatile("ld.global.b16 %0, [%1];" : "=h"(v) : "l"(p));
    return (int16_t)v;
}
__device__ __forceinline__ unsigned ld_now8(const void *p)
{
    unsigned v;
    asm volatile("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// The horizontal-edge pass is launched with programmatic stream serialisation (launch_deblock): its CTAs become resident while the vertical-edge
// pass drains, fetch the maps - which the vertical pass does not write - and only then wait for the vertical pass to complete and become visible.
template <bool VERTICAL>
__device__ __forceinline__ void dbk_pass_order()
{
    if (VERTICAL) xb_grid_release();
    else xb_grid_wait();
}

// xevdm_get_tbl_qp_to_st (xevdm_df.c:38-104): 0 intra, 1 luma cbf, 2 motion differs (or IBC), 3 no filtering.  m = map_scu word, r = both
// reference indices (s8 pair), v = both vectors of the current (0) and the neighbouring (1) SCU
__device__ __forceinline__ int dbk_class_v(uint32_t m0, uint32_t m1, int r0, int r1, int2 v0, int2 v1)
{
    if (((m0 | m1) >> 15) & 1) return 0;
    if (((m0 | m1) >> 24) & 1) return 1;
    if (((m0 | m1) >> 26) & 1) return 2;
    const int8_t r00 = (int8_t)(r0 & 0xff), r01 = (int8_t)(r0 >> 8), r10 = (int8_t)(r1 & 0xff), r11 = (int8_t)(r1 >> 8);
    int a0x = (int16_t)(v0.x & 0xffff), a0y = v0.x >> 16, a1x = (int16_t)(v0.y & 0xffff), a1y = v0.y >> 16;
    int b0x = (int16_t)(v1.x & 0xffff), b0y = v1.x >> 16, b1x = (int16_t)(v1.y & 0xffff), b1y = v1.y >> 16;
    if (r00 < 0) a0x = a0y = 0;
    if (r01 < 0) a1x = a1y = 0;
    if (r10 < 0) b0x = b0y = 0;
    if (r11 < 0) b1x = b1y = 0;
    if (r00 == r10 && r01 == r11)
        return (abs(a0x - b0x) >= 4 || abs(a0y - b0y) >= 4 || abs(a1x - b1x) >= 4 || abs(a1y - b1y) >= 4) ? 2 : 3;
    if (r00 == r11 && r01 == r10)
        return (abs(a0x - b1x) >= 4 || abs(a0y - b1y) >= 4 || abs(a1x - b0x) >= 4 || abs(a1y - b0y) >= 4) ? 2 : 3;
    return 2;
}
__device__ __forceinline__ int dbk_class(const DbkArgs &a, int cur, int nb)
{
    const uint32_t m0 = a.map_scu[cur], m1 = a.map_scu[nb];
    if ((((m0 | m1) >> 15) & 1) || (((m0 | m1) >> 24) & 1) || (((m0 | m1) >> 26) & 1)) return dbk_class_v(m0, m1, 0, 0, make_int2(0, 0), make_int2(0, 0));
    return dbk_class_v(m0, m1, ((const int16_t *)a.map_refi)[cur], ((const int16_t *)a.map_refi)[nb], ((const int2 *)a.map_mv)[cur], ((const int2 *)a.map_mv)[nb]);
}

// deblock_scu_hor / _ver (xevd_df.c:96-134): all intermediates are s16 in the reference; `/` truncates toward zero (T6)
__device__ __forceinline__ void dbk_luma(int &A, int &B, int &C, int &D, int st, int maxv)
{
    const int d = (int16_t)((A - (B << 2) + (C << 2) - D) / 8);
    const int ad = abs(d);
    const int t16 = max(0, (ad - st) << 1);
    int clip = max(0, ad - t16);
    const int d1 = d < 0 ? -clip : clip;
    clip >>= 1;
    const int d2 = xb_clip3(-clip, clip, (A - D) / 4);
    A = xb_clip3(0, maxv, (int16_t)(A - d2));
    B = xb_clip3(0, maxv, (int16_t)(B + d1));
    C = xb_clip3(0, maxv, (int16_t)(C - d1));
    D = xb_clip3(0, maxv, (int16_t)(D + d2));
}
__device__ __forceinline__ void dbk_chroma(int A, int &B, int &C, int D, int st, int maxv)
{
    const int d = (int16_t)((A - (B << 2) + (C << 2) - D) / 8);
    const int ad = abs(d);
    const int t16 = max(0, (ad - st) << 1);
    const int clip = max(0, ad - t16);
    const int d1 = d < 0 ? -clip : clip;
    B = xb_clip3(0, maxv, (int16_t)(B + d1));
    C = xb_clip3(0, maxv, (int16_t)(C - d1));
}

__device__ __forceinline__ bool dbk_has_edge(const DbkArgs &a, int sx, int sy, bool vertical)
{
    if (vertical) return sx > 0 && (a.map_edge[sy * a.w_scu + sx] & XB200_EDGE_LEFT);
    return sy > 0 && (a.map_edge[sy * a.w_scu + sx] & XB200_EDGE_TOP);
}

// chroma edges: the inner leaf boundaries of a local dual tree node are luma edges only (XB200_EDGE_*_NOC; xevdm_df.c:155-160,245-250)
__device__ __forceinline__ bool dbk_has_edge_c(const DbkArgs &a, int sx, int sy, bool vertical)
{
    if (vertical) return sx > 0 && (a.map_edge[sy * a.w_scu + sx] & (XB200_EDGE_LEFT | XB200_EDGE_LEFT_NOC)) == XB200_EDGE_LEFT;
    return sy > 0 && (a.map_edge[sy * a.w_scu + sx] & (XB200_EDGE_TOP | XB200_EDGE_TOP_NOC)) == XB200_EDGE_TOP;
}

// strengths of one segment: luma, Cb, Cr.  QP is the one of the CURRENT (right / lower) SCU only (xevd_df.c:347,446; T7)
__device__ __forceinline__ void dbk_strengths_v(const DbkArgs &a, int cls, int qp, int &st, int &st_u, int &st_v)
{
    st = dbk_st(cls, qp) << (a.bd_l - 8);
    const int qu = xb_clip3(-6 * (a.bd_c - 8), 57, qp + a.qp_u_offset), qv = xb_clip3(-6 * (a.bd_c - 8), 57, qp + a.qp_v_offset);
    st_u = dbk_st(cls, qu < 0 ? qu : a.cq[0][qu]) << (a.bd_c - 8);
    st_v = dbk_st(cls, qv < 0 ? qv : a.cq[1][qv]) << (a.bd_c - 8);
}
__device__ __forceinline__ void dbk_strengths(const DbkArgs &a, int cur, int nb, int &st, int &st_u, int &st_v)
{
    dbk_strengths_v(a, dbk_class(a, cur, nb), (a.map_scu[cur] >> 16) & 0x7f, st, st_u, st_v);
}

// both chroma planes of the 4-sample (2 chroma samples) segment whose right / lower SCU is (cx, cy)
template <bool VERTICAL>
__device__ __forceinline__ void dbk_chroma_segment(const DbkArgs &a, int cx, int cy, int maxc)
{
    const int c2 = cy * a.w_scu + cx, n2 = VERTICAL ? c2 - 1 : c2 - a.w_scu;
    int st, st_u, st_v;
    dbk_strengths(a, c2, n2, st, st_u, st_v);
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int s = k ? st_v : st_u;
        if (!s) continue;
        pel *p = (k ? a.v : a.u) + (size_t)(cy * 2) * a.s_c + cx * 2;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            if (VERTICAL) {
                pel *q = p + (size_t)i * a.s_c;
                int A = q[-2], B = q[-1], C = q[0], D = q[1];
                dbk_chroma(A, B, C, D, s, maxc);
                q[-1] = (pel)B; q[0] = (pel)C;
            } else {
                pel *q = p + i;
                int A = q[-2 * a.s_c], B = q[-a.s_c], C = q[0], D = q[a.s_c];
                dbk_chroma(A, B, C, D, s, maxc);
                q[-a.s_c] = (pel)B; q[0] = (pel)C;
            }
        }
    }
}

// when the reference filters the vertical edge left of SCU (sx, sy): at the visit of the later of the two chroma-owning CUs.  CTUs are
// visited in raster order, CUs inside a CTU in decoding order (map_order)
__device__ __forceinline__ int dbk_visit_time(const DbkArgs &a, int sx, int sy)
{
    const int p = sy * a.w_scu + sx, sh = a.log2_ctu - 2;
    const int t1 = ((sx >> sh) << 16) | a.map_order[p], t0 = (((sx - 1) >> sh) << 16) | a.map_order[p - 1];
    return max(t0, t1);
}

// loop_filter_across_tiles_enabled_flag == 0: CU edges that coincide with a tile boundary lose their flags before the passes run
// (the reference tests map_tidx on both sides of every edge, xevdm_df.c:142,233,877,1088).  blockIdx.y = one interior boundary; column
// and row boundaries are two launches, because the SCU where they cross is one byte that both would read-modify-write.
struct TileEdgeArgs {
    uint8_t *map_edge;
    int w_scu, h_scu, log2_ctu_scu, n_cols, n_rows;
    uint16_t col_bd[XB200_MAX_TILE_COLS + 1], row_bd[XB200_MAX_TILE_ROWS + 1];
};
template <bool COLS>
__global__ void __launch_bounds__(256) k_clear_tile_edges(const __grid_constant__ TileEdgeArgs a)
{
    const int b = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    if (COLS) {
        const int sx = (int)a.col_bd[b + 1] << a.log2_ctu_scu;
        if (sx < a.w_scu && i < a.h_scu) a.map_edge[i * a.w_scu + sx] &= (uint8_t)~(XB200_EDGE_LEFT | XB200_EDGE_LEFT_NOC);
    } else {
        const int sy = (int)a.row_bd[b + 1] << a.log2_ctu_scu;
        if (sy < a.h_scu && i < a.w_scu) a.map_edge[sy * a.w_scu + i] &= (uint8_t)~(XB200_EDGE_TOP | XB200_EDGE_TOP_NOC);
    }
}

template <bool VERTICAL>
__global__ void __launch_bounds__(256) k_deblock(const __grid_constant__ DbkArgs a)
{
    const int sx = blockIdx.x * 32 + (threadIdx.x & 31), sy = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (sx >= a.w_scu || sy >= a.h_scu) return;
    if (VERTICAL) xb_grid_wait();          // the maps and samples come from the kernels before; the horizontal pass waits below, after its map loads
    constexpr unsigned kEdge = VERTICAL ? XB200_EDGE_LEFT : XB200_EDGE_TOP, kNoc = VERTICAL ? XB200_EDGE_LEFT_NOC : XB200_EDGE_TOP_NOC;
    const bool inner = VERTICAL ? sx > 0 : sy > 0, has_next = VERTICAL ? sx + 1 < a.w_scu : sy + 1 < a.h_scu;
    const int cur = sy * a.w_scu + sx, nb = inner ? (VERTICAL ? cur - 1 : cur - a.w_scu) : cur, nx = has_next ? (VERTICAL ? cur + 1 : cur + a.w_scu) : cur;
    // ---- everything the common case reads, in one round trip (border padding makes the sample addresses of picture-edge SCUs valid) ----
    const unsigned e = ld_now8(a.map_edge + cur), e_prev = ld_now8(a.map_edge + nb), e_next = ld_now8(a.map_edge + nx);
    const uint32_t m0 = (uint32_t)ld_now32(a.map_scu + cur), m1 = (uint32_t)ld_now32(a.map_scu + nb);
    const int r0 = ld_now16((const int16_t *)a.map_refi + cur), r1 = ld_now16((const int16_t *)a.map_refi + nb);
    const int2 v0 = ld_now64((const int2 *)a.map_mv + cur), v1 = ld_now64((const int2 *)a.map_mv + nb);
    dbk_pass_order<VERTICAL>();
    pel *p = a.y + (size_t)(sy * 4) * a.s_l + sx * 4;
    pel *pc[2] = {a.u + (size_t)(sy * 2) * a.s_c + sx * 2, a.v + (size_t)(sy * 2) * a.s_c + sx * 2};
    int2 l[4];            // VERTICAL: row i, samples A B | C D around the edge; else row j - 2, the segment's four columns
    int2 c[2][2];         // VERTICAL: plane k, row i, samples A B | C D; else plane k, rows (-2, -1) and (0, 1), the segment's two columns
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (VERTICAL) { l[i].x = ld_now32(p + (size_t)i * a.s_l - 2); l[i].y = ld_now32(p + (size_t)i * a.s_l); }
        else l[i] = ld_now64(p + (ptrdiff_t)(i - 2) * a.s_l);
    }
#pragma unroll
    for (int k = 0; k < 2; k++)
#pragma unroll
        for (int i = 0; i < 2; i++) {
            if (VERTICAL) { c[k][i].x = ld_now32(pc[k] + (size_t)i * a.s_c - 2); c[k][i].y = ld_now32(pc[k] + (size_t)i * a.s_c); }
            else { c[k][i].x = ld_now32(pc[k] + (ptrdiff_t)(2 * i - 2) * a.s_c); c[k][i].y = ld_now32(pc[k] + (ptrdiff_t)(2 * i - 1) * a.s_c); }
        }
    if (!(inner && (e & kEdge))) return;
    int st, st_u, st_v;
    dbk_strengths_v(a, dbk_class_v(m0, m1, r0, r1, v0, v1), (m0 >> 16) & 0x7f, st, st_u, st_v);
    const int maxl = (1 << a.bd_l) - 1, maxc = (1 << a.bd_c) - 1;
    if (st) {
        if (VERTICAL) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                pel *q = p + (size_t)i * a.s_l;
                int A = (int16_t)(l[i].x & 0xffff), B = l[i].x >> 16, C = (int16_t)(l[i].y & 0xffff), D = l[i].y >> 16;
                dbk_luma(A, B, C, D, st, maxl);
                *(int *)(q - 2) = (A & 0xffff) | (B << 16);
                *(int *)q = (C & 0xffff) | (D << 16);
            }
        } else {
            int r[4][4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                r[j][0] = (int16_t)(l[j].x & 0xffff); r[j][1] = l[j].x >> 16; r[j][2] = (int16_t)(l[j].y & 0xffff); r[j][3] = l[j].y >> 16;
            }
#pragma unroll
            for (int i = 0; i < 4; i++) dbk_luma(r[0][i], r[1][i], r[2][i], r[3][i], st, maxl);
#pragma unroll
            for (int j = 0; j < 4; j++)
                *(int2 *)(p + (ptrdiff_t)(j - 2) * a.s_l) = make_int2((r[j][0] & 0xffff) | (r[j][1] << 16), (r[j][2] & 0xffff) | (r[j][3] << 16));
        }
    }
    // chroma: only the head of a run of consecutive segments works; it walks the run in the reference's order
    if ((e & (kEdge | kNoc)) != kEdge) return;
    if ((VERTICAL ? sx - 1 > 0 : sy - 1 > 0) && (e_prev & (kEdge | kNoc)) == kEdge) return;
    if (!(has_next && (e_next & (kEdge | kNoc)) == kEdge)) {
        // a segment without neighbours in its row / column (any CU wider / higher than 4): nobody else touches its samples in this pass
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int s = k ? st_v : st_u;
            if (!s) continue;
#pragma unroll
            for (int i = 0; i < 2; i++) {
                if (VERTICAL) {
                    pel *q = pc[k] + (size_t)i * a.s_c;
                    int A = (int16_t)(c[k][i].x & 0xffff), B = c[k][i].x >> 16, C = (int16_t)(c[k][i].y & 0xffff), D = c[k][i].y >> 16;
                    dbk_chroma(A, B, C, D, s, maxc);
                    q[-1] = (pel)B; q[0] = (pel)C;
                } else {
                    pel *q = pc[k] + i;
                    int A = i ? c[k][0].x >> 16 : (int16_t)(c[k][0].x & 0xffff), B = i ? c[k][0].y >> 16 : (int16_t)(c[k][0].y & 0xffff);
                    int C = i ? c[k][1].x >> 16 : (int16_t)(c[k][1].x & 0xffff), D = i ? c[k][1].y >> 16 : (int16_t)(c[k][1].y & 0xffff);
                    dbk_chroma(A, B, C, D, s, maxc);
                    q[-a.s_c] = (pel)B; q[0] = (pel)C;
                }
            }
        }
        return;
    }
    if (VERTICAL && a.map_order) {
        // SUCO: the reference filters an edge when the LATER of its two CUs is visited (its left edge, then its right edge; xevdm_df.c:272-300),
        // so inside a run the order is not left-to-right.  Only neighbouring edges interact: walk every maximal stretch of decreasing visit
        // times from its right end back to its left end, the stretches themselves left to right.
        int cx = sx;
        while (true) {
            int last = cx;
            while (last + 1 < a.w_scu && dbk_has_edge_c(a, last + 1, sy, true) && dbk_visit_time(a, last, sy) > dbk_visit_time(a, last + 1, sy)) last++;
            for (int x = last; x >= cx; x--) dbk_chroma_segment<true>(a, x, sy, maxc);
            cx = last + 1;
            if (cx >= a.w_scu || !dbk_has_edge_c(a, cx, sy, true)) break;
        }
        return;
    }
    int cx = sx, cy = sy;
    while (true) {
        dbk_chroma_segment<VERTICAL>(a, cx, cy, maxc);
        if (VERTICAL) { cx++; if (cx >= a.w_scu) break; } else { cy++; if (cy >= a.h_scu) break; }
        if (!dbk_has_edge_c(a, cx, cy, VERTICAL)) break;
    }
}

// ---- Main-profile deblocking (sps->tool_addb): H.264-style filter on the 8x8 luma grid -------------------------------------
// Replaces get_bs, deblock_scu_line_luma / _chroma and deblock_addb_cu_hor / _ver (src_main/xevdm_df.c:361-1135) for one tile /
// one slice, TREE_LC, no ATS-inter.  Edges are 8 luma samples apart and a line touches at most 4 samples per side, so every
// segment is independent: one thread per 4-line segment, no ordering constraints.
__constant__ uint8_t c_addb_alpha[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 17, 20, 22, 25, 28, 32, 36, 40, 45,
                                         50, 56, 63, 71, 80, 90, 101, 113, 127, 144, 162, 182, 203, 226, 255, 255};
__constant__ uint8_t c_addb_beta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10,
                                        11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18};
__constant__ uint8_t c_addb_clip[52][5] = {
    {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0},
    {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0},
    {0, 0, 0, 0, 0}, {0, 0, 0, 1, 1}, {0, 0, 0, 1, 1}, {0, 0, 0, 1, 1}, {0, 0, 0, 1, 1}, {0, 0, 1, 1, 1}, {0, 0, 1, 1, 1}, {0, 1, 1, 1, 1},
    {0, 1, 1, 1, 1}, {0, 1, 1, 1, 1}, {0, 1, 1, 1, 1}, {0, 1, 1, 2, 2}, {0, 1, 1, 2, 2}, {0, 1, 1, 2, 2}, {0, 1, 1, 2, 2}, {0, 1, 2, 3, 3},
    {0, 1, 2, 3, 3}, {0, 2, 2, 3, 3}, {0, 2, 2, 4, 4}, {0, 2, 3, 4, 4}, {0, 2, 3, 4, 4}, {0, 3, 3, 5, 5}, {0, 3, 4, 6, 6}, {0, 3, 4, 6, 6},
    {0, 4, 5, 7, 7}, {0, 4, 5, 8, 8}, {0, 4, 6, 9, 9}, {0, 5, 7, 10, 10}, {0, 6, 8, 11, 11}, {0, 6, 8, 13, 13}, {0, 7, 10, 14, 14}, {0, 8, 11, 16, 16},
    {0, 9, 12, 18, 18}, {0, 10, 13, 20, 20}, {0, 11, 15, 23, 23}, {0, 13, 17, 25, 25}};

__device__ __forceinline__ int addb_index(int qp, int offset) { return xb_clip3(0, 51, (qp & 0xff) + (offset & 0xff)); }   // u8 arguments in the reference
__device__ __forceinline__ bool addb_near(int ax, int ay, int bx, int by) { return abs(ax - bx) < 4 && abs(ay - by) < 4; }

// get_bs (xevdm_df.c:361-470).  m = map_scu words, ats = ats_present of either CU, r = reference index pairs, v = vector pairs (the caller picks
// map_unrefined_mv where the DMVR flag of m is set), cross_ctu = the edge lies on a CTU boundary
__device__ __forceinline__ int addb_bs_v(const DbkArgs &a, uint32_t m0, uint32_t m1, bool ats, int r0, int r1, int2 v0, int2 v1, bool cross_ctu)
{
    const bool intra = ((m0 | m1) >> 15) & 1;
    if (intra) return cross_ctu ? 4 : 3;
    if (((m0 | m1) >> 26) & 1) return 3;
    if ((((m0 | m1) >> 24) & 1) || ats) return 2;     // luma cbf or ats_present (xevdm_df.c:415)
    const int8_t r00 = (int8_t)(r0 & 0xff), r01 = (int8_t)(r0 >> 8), r10 = (int8_t)(r1 & 0xff), r11 = (int8_t)(r1 >> 8);
    const int pa0 = r00 >= 0 ? a.ref_id[0][r00] : -1, pa1 = r01 >= 0 ? a.ref_id[1][r01] : -1;
    const int pb0 = r10 >= 0 ? a.ref_id[0][r10] : -1, pb1 = r11 >= 0 ? a.ref_id[1][r11] : -1;
    int a0x = (int16_t)(v0.x & 0xffff), a0y = v0.x >> 16, a1x = (int16_t)(v0.y & 0xffff), a1y = v0.y >> 16;
    int b0x = (int16_t)(v1.x & 0xffff), b0y = v1.x >> 16, b1x = (int16_t)(v1.y & 0xffff), b1y = v1.y >> 16;
    if (r00 < 0) a0x = a0y = 0;
    if (r01 < 0) a1x = a1y = 0;
    if (r10 < 0) b0x = b0y = 0;
    if (r11 < 0) b1x = b1y = 0;
    const bool same = pa0 == pb0 && pa1 == pb1, cross = pa0 == pb1 && pa1 == pb0;
    if (!(same || cross)) return 1;
    const bool s00 = addb_near(a0x, a0y, b0x, b0y), s11 = addb_near(a1x, a1y, b1x, b1y);
    const bool s01 = addb_near(a0x, a0y, b1x, b1y), s10 = addb_near(a1x, a1y, b0x, b0y);
    if (pa0 == pa1) return (s00 && s11 && s01 && s10) ? 0 : 1;
    if (same) return (s00 && s11) ? 0 : 1;
    return (s01 && s10) ? 0 : 1;
}

// deblock_scu_line_luma (xevdm_df.c:584-706): p[i] = sample i+1 before the edge, q[i] = sample i after it
__device__ __forceinline__ void addb_luma(int (&p)[4], int (&q)[4], int bs, int alpha, int beta, int c1, int bd)
{
    if (!(bs && abs(p[0] - q[0]) < alpha && abs(p[1] - p[0]) < beta && abs(q[1] - q[0]) < beta)) return;
    const int maxv = (1 << bd) - 1;
    const int ap = abs(p[0] - p[2]) < beta, aq = abs(q[0] - q[2]) < beta;
    int po0 = p[0], po1 = p[1], po2 = p[2], qo0 = q[0], qo1 = q[1], qo2 = q[2];
    if (bs == 4) {
        const bool small = abs(p[0] - q[0]) < ((alpha >> 2) + 2);
        if (ap && small) {
            po0 = (p[2] + 2 * (p[1] + p[0] + q[0]) + q[1] + 4) >> 3;
            po1 = (p[2] + p[1] + p[0] + q[0] + 2) >> 2;
            po2 = (2 * p[3] + 3 * p[2] + p[1] + p[0] + q[0] + 4) >> 3;
        } else po0 = (2 * p[1] + p[0] + q[1] + 2) >> 2;
        if (aq && small) {
            qo0 = (q[2] + 2 * (q[1] + q[0] + p[0]) + p[1] + 4) >> 3;
            qo1 = (q[2] + q[1] + q[0] + p[0] + 2) >> 2;
            qo2 = (2 * q[3] + 3 * q[2] + q[1] + q[0] + p[0] + 4) >> 3;
        } else qo0 = (2 * q[1] + q[0] + p[1] + 2) >> 2;
    } else {
        const int c0 = (c1 + ((ap + aq) << max(0, bd - 9))) & 0xff;
        const int d0 = xb_clip3(-c0, c0, (4 * (q[0] - p[0]) + p[1] - q[1] + 4) >> 3);
        po0 = xb_clip3(0, maxv, p[0] + d0);
        qo0 = xb_clip3(0, maxv, q[0] - d0);
        if (ap) po1 = (int16_t)(p[1] + xb_clip3(-c1, c1, (((p[2] + p[0] + q[0]) * 3) - 8 * p[1] - q[1]) >> 4));
        if (aq) qo1 = (int16_t)(q[1] + xb_clip3(-c1, c1, (((q[2] + q[0] + p[0]) * 3) - 8 * q[1] - p[1]) >> 4));
    }
    p[0] = xb_clip3(0, maxv, po0); p[1] = xb_clip3(0, maxv, po1); p[2] = xb_clip3(0, maxv, po2); p[3] = xb_clip3(0, maxv, p[3]);
    q[0] = xb_clip3(0, maxv, qo0); q[1] = xb_clip3(0, maxv, qo1); q[2] = xb_clip3(0, maxv, qo2); q[3] = xb_clip3(0, maxv, q[3]);
}
__device__ __forceinline__ void addb_chroma(int (&p)[2], int (&q)[2], int bs, int alpha, int beta, int c0, int bd)
{
    if (!(bs && abs(p[0] - q[0]) < alpha && abs(p[1] - p[0]) < beta && abs(q[1] - q[0]) < beta)) return;
    const int maxv = (1 << bd) - 1;
    int po0, qo0;
    if (bs == 4) {
        po0 = (2 * p[1] + p[0] + q[1] + 2) >> 2;
        qo0 = (2 * q[1] + q[0] + p[1] + 2) >> 2;
    } else {
        const int d0 = xb_clip3(-c0, c0, (4 * (q[0] - p[0]) + p[1] - q[1] + 4) >> 3);
        po0 = p[0] + d0; qo0 = q[0] - d0;
    }
    p[0] = xb_clip3(0, maxv, po0); q[0] = xb_clip3(0, maxv, qo0);
    p[1] = xb_clip3(0, maxv, p[1]); q[1] = xb_clip3(0, maxv, q[1]);
}

// One thread per segment of the 8x8 luma grid: lane <-> every second SCU column (vertical edges) / warp <-> every second SCU row (horizontal).
template <bool VERTICAL>
__global__ void __launch_bounds__(256) k_deblock_addb(const __grid_constant__ DbkArgs a)
{
    const int gi = blockIdx.x * 32 + (threadIdx.x & 31), gj = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int sx = VERTICAL ? 2 * gi : gi, sy = VERTICAL ? gj : 2 * gj;
    if (sx >= a.w_scu || sy >= a.h_scu) return;
    if (VERTICAL) xb_grid_wait();          // the maps and samples come from the kernels before; the horizontal pass waits below, after its map loads
    constexpr unsigned kEdge = VERTICAL ? XB200_EDGE_LEFT : XB200_EDGE_TOP, kNoc = VERTICAL ? XB200_EDGE_LEFT_NOC : XB200_EDGE_TOP_NOC;
    const bool inner = VERTICAL ? sx > 0 : sy > 0;
    const int cur = sy * a.w_scu + sx, nb = inner ? (VERTICAL ? cur - 1 : cur - a.w_scu) : cur;
    const int x = sx << 2, y = sy << 2;
    // ---- everything the segment reads, in one round trip (see ld_now32) ----
    const unsigned e = ld_now8(a.map_edge + cur), e_nb = ld_now8(a.map_edge + nb);
    const uint32_t m0 = (uint32_t)ld_now32(a.map_scu + cur), m1 = (uint32_t)ld_now32(a.map_scu + nb);
    const int r0 = ld_now16((const int16_t *)a.map_refi + cur), r1 = ld_now16((const int16_t *)a.map_refi + nb);
    int2 v0 = ld_now64((const int2 *)a.map_mv + cur), v1 = ld_now64((const int2 *)a.map_mv + nb);
    dbk_pass_order<VERTICAL>();
    pel *base = a.y + (size_t)y * a.s_l + x;
    pel *cb[2] = {a.u + (size_t)(y >> 1) * a.s_c + (x >> 1), a.v + (size_t)(y >> 1) * a.s_c + (x >> 1)};
    int2 l[8];            // VERTICAL: row i: samples p3 p2 p1 p0 (l[2i]) | q0 q1 q2 q3 (l[2i+1]); else rows -4..3, the segment's four columns
    int c[2][4];          // VERTICAL: plane k, row i: p1 p0 (c[k][2i]) | q0 q1 (c[k][2i+1]); else plane k, rows -2..1, the segment's two columns
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (VERTICAL) { l[2 * i] = ld_now64(base + (size_t)i * a.s_l - 4); l[2 * i + 1] = ld_now64(base + (size_t)i * a.s_l); }
        else { l[2 * i] = ld_now64(base + (ptrdiff_t)(2 * i - 4) * a.s_l); l[2 * i + 1] = ld_now64(base + (ptrdiff_t)(2 * i - 3) * a.s_l); }
    }
#pragma unroll
    for (int k = 0; k < 2; k++)
#pragma unroll
        for (int i = 0; i < 2; i++) {
            if (VERTICAL) { c[k][2 * i] = ld_now32(cb[k] + (size_t)i * a.s_c - 2); c[k][2 * i + 1] = ld_now32(cb[k] + (size_t)i * a.s_c); }
            else { c[k][2 * i] = ld_now32(cb[k] + (ptrdiff_t)(2 * i - 2) * a.s_c); c[k][2 * i + 1] = ld_now32(cb[k] + (ptrdiff_t)(2 * i - 1) * a.s_c); }
        }
    if (!(inner && (e & kEdge))) return;
    // xevdm_deblock copies map_mv over map_unrefined_mv wherever the DMVR flag is not set before it walks the tree (xevdm.c:2077-2090)
    if ((m0 >> 25) & 1) v0 = ((const int2 *)a.map_umv)[cur];
    if ((m1 >> 25) & 1) v1 = ((const int2 *)a.map_umv)[nb];
    const int x1 = VERTICAL ? x - 1 : x, y1 = VERTICAL ? y : y - 1;
    const int bs = addb_bs_v(a, m0, m1, ((e | e_nb) & XB200_EDGE_ATS) != 0, r0, r1, v0, v1,
                             (x >> a.log2_ctu) != (x1 >> a.log2_ctu) || (y >> a.log2_ctu) != (y1 >> a.log2_ctu));
    const int qp = (((m0 >> 16) & 0x7f) + ((m1 >> 16) & 0x7f) + 1) >> 1;
    const int scale = a.bd_l - 8;
    {
        const int ia = addb_index(qp, a.alpha_offset), ib = addb_index(qp, a.beta_offset);
        const int alpha = (c_addb_alpha[ia] << scale) & 0xffff, beta = (c_addb_beta[ib] << scale) & 0xff;
        const int c1 = (c_addb_clip[ia][bs] << max(0, a.bd_l - 9)) & 0xff;
        if (VERTICAL) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                int p[4], q[4];
                pel *r = base + (size_t)i * a.s_l;
                const int2 lo = l[2 * i], hi = l[2 * i + 1];
                p[3] = (int16_t)(lo.x & 0xffff); p[2] = lo.x >> 16; p[1] = (int16_t)(lo.y & 0xffff); p[0] = lo.y >> 16;
                q[0] = (int16_t)(hi.x & 0xffff); q[1] = hi.x >> 16; q[2] = (int16_t)(hi.y & 0xffff); q[3] = hi.y >> 16;
                addb_luma(p, q, bs, alpha, beta, c1, a.bd_l);
                *(int2 *)(r - 4) = make_int2((p[3] & 0xffff) | (p[2] << 16), (p[1] & 0xffff) | (p[0] << 16));
                *(int2 *)r = make_int2((q[0] & 0xffff) | (q[1] << 16), (q[2] & 0xffff) | (q[3] << 16));
            }
        } else {
            int o[8][4];     // filtered rows -4..3
#pragma unroll
            for (int i = 0; i < 4; i++) {
                int p[4], q[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int2 rq = l[4 + k], rp = l[3 - k];
                    const int wq = i < 2 ? rq.x : rq.y, wp = i < 2 ? rp.x : rp.y;
                    q[k] = (i & 1) ? wq >> 16 : (int16_t)(wq & 0xffff);
                    p[k] = (i & 1) ? wp >> 16 : (int16_t)(wp & 0xffff);
                }
                addb_luma(p, q, bs, alpha, beta, c1, a.bd_l);
#pragma unroll
                for (int k = 0; k < 4; k++) { o[4 + k][i] = q[k]; o[3 - k][i] = p[k]; }
            }
#pragma unroll
            for (int j = 0; j < 8; j++)
                *(int2 *)(base + (ptrdiff_t)(j - 4) * a.s_l) = make_int2((o[j][0] & 0xffff) | (o[j][1] << 16), (o[j][2] & 0xffff) | (o[j][3] << 16));
        }
    }
    if ((e & (kEdge | kNoc)) != kEdge) return;      // inner leaf boundary of a local dual tree node: a luma edge only (xevdm_df.c:916-920,986-997)
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int qc = xb_clip3(-6 * (a.bd_c - 8), 57, qp + (k ? a.qp_v_offset : a.qp_u_offset));
        const int qm = qc < 0 ? qc : a.cq[k][qc];
        const int ia = addb_index(qm, a.alpha_offset), ib = addb_index(qm, a.beta_offset);
        const int alpha = (c_addb_alpha[ia] << scale) & 0xffff, beta = (c_addb_beta[ib] << scale) & 0xff;
        const int c0 = ((c_addb_clip[ia][bs] + 1) << max(0, a.bd_c - 9)) & 0xff;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            int p[2], q[2];
            if (VERTICAL) {
                p[1] = (int16_t)(c[k][2 * i] & 0xffff); p[0] = c[k][2 * i] >> 16; q[0] = (int16_t)(c[k][2 * i + 1] & 0xffff); q[1] = c[k][2 * i + 1] >> 16;
            } else {
                p[1] = i ? c[k][0] >> 16 : (int16_t)(c[k][0] & 0xffff); p[0] = i ? c[k][1] >> 16 : (int16_t)(c[k][1] & 0xffff);
                q[0] = i ? c[k][2] >> 16 : (int16_t)(c[k][2] & 0xffff); q[1] = i ? c[k][3] >> 16 : (int16_t)(c[k][3] & 0xffff);
            }
            addb_chroma(p, q, bs, alpha, beta, c0, a.bd_c);
            const ptrdiff_t st = VERTICAL ? 1 : a.s_c;
            pel *cc = VERTICAL ? cb[k] + (size_t)i * a.s_c : cb[k] + i;
            cc[0] = (pel)q[0]; cc[st] = (pel)q[1]; cc[-st] = (pel)p[0]; cc[-2 * st] = (pel)p[1];
        }
    }
}

inline void launch_deblock(const DbkArgs &a, bool addb, cudaStream_t st)
{
    // rows of SCUs (warps) per CTA: 1, 2, 4 and 8 measure the same (profiles/r2/ab_log.txt) - the passes are not bound by resident warps
    constexpr int rv = 8, rh = 8;
    if (addb) {
        xb_launch_early(k_deblock_addb<true>, dim3(((a.w_scu + 1) / 2 + 31) / 32, (a.h_scu + rv - 1) / rv), dim3(32 * rv), 0, st, a);
        xb_launch_early(k_deblock_addb<false>, dim3((a.w_scu + 31) / 32, ((a.h_scu + 1) / 2 + rh - 1) / rh), dim3(32 * rh), 0, st, a);
    } else {
        xb_launch_early(k_deblock<true>, dim3((a.w_scu + 31) / 32, (a.h_scu + rv - 1) / rv), dim3(32 * rv), 0, st, a);
        xb_launch_early(k_deblock<false>, dim3((a.w_scu + 31) / 32, (a.h_scu + rh - 1) / rh), dim3(32 * rh), 0, st, a);
    }
}

// ---- output path: DRA + crop + bit-depth conversion into packed planes (xb200_pic_pull) ---------------------------------------------
struct OutArgs {
    const pel *y, *u, *v;
    int s_l, s_c;
    int x0, y0, w, h;               // luma crop window (chroma = half)
    const int *lut_l, *lut_c;       // device LUTs: luma_inv_scale_lut[1024], int_chroma_inv_scale_lut[2][1024]; null = no DRA
    void *out_y, *out_u, *out_v;    // packed, row stride = plane width
    int out8;
};
// one thread per chroma sample: its 2x2 luma samples and the two chroma samples
__global__ void __launch_bounds__(256) k_output(const __grid_constant__ OutArgs a)
{
    const int cw = a.w >> 1, ch = a.h >> 1;
    const int k = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (k >= cw || j >= ch) return;
    const pel *py = a.y + (size_t)(a.y0 + 2 * j) * a.s_l + a.x0 + 2 * k;
    int l[4] = {py[0], py[1], py[a.s_l], py[a.s_l + 1]};
    int c[2] = {a.u[(size_t)((a.y0 >> 1) + j) * a.s_c + (a.x0 >> 1) + k], a.v[(size_t)((a.y0 >> 1) + j) * a.s_c + (a.x0 >> 1) + k]};
    if (a.lut_l) {
        // chroma first, scaled by the factor of the UNMAPPED co-located luma sample (xevdm_dra.c:301-354), then luma (:272-300)
        const int ref = max(l[0], 0);
#pragma unroll
        for (int p = 0; p < 2; p++) {
            const int sv = (int16_t)(c[p] - 512);
            int off = (abs(sv) * a.lut_c[p * 1024 + ref] + 256) >> 9;
            c[p] = (int16_t)(512 + (sv < 0 ? -off : off));
        }
#pragma unroll
        for (int p = 0; p < 4; p++) l[p] = (int16_t)a.lut_l[l[p]];
    }
    if (a.out8) {
        uint8_t *oy = (uint8_t *)a.out_y + (size_t)(2 * j) * a.w + 2 * k;
        oy[0] = (uint8_t)xb_clip3(0, 255, (l[0] + 2) >> 2); oy[1] = (uint8_t)xb_clip3(0, 255, (l[1] + 2) >> 2);
        oy[a.w] = (uint8_t)xb_clip3(0, 255, (l[2] + 2) >> 2); oy[a.w + 1] = (uint8_t)xb_clip3(0, 255, (l[3] + 2) >> 2);
        ((uint8_t *)a.out_u)[(size_t)j * cw + k] = (uint8_t)xb_clip3(0, 255, (c[0] + 2) >> 2);
        ((uint8_t *)a.out_v)[(size_t)j * cw + k] = (uint8_t)xb_clip3(0, 255, (c[1] + 2) >> 2);
    } else {
        int16_t *oy = (int16_t *)a.out_y + (size_t)(2 * j) * a.w + 2 * k;
        oy[0] = (int16_t)l[0]; oy[1] = (int16_t)l[1]; oy[a.w] = (int16_t)l[2]; oy[a.w + 1] = (int16_t)l[3];
        ((int16_t *)a.out_u)[(size_t)j * cw + k] = (int16_t)c[0];
        ((int16_t *)a.out_v)[(size_t)j * cw + k] = (int16_t)c[1];
    }
}

}  // namespace xb
