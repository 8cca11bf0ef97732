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
    k_pad<<<a.row_start[3], 128, 0, st>>>(a);
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
    const int16_t *map_mv;
    const int8_t *map_refi;
    const uint8_t *map_edge;
    int8_t cq[2][58];            // chroma QP mapping for qp >= 0 (xevd_qp_chroma_dynamic); identity below 0
};

__device__ __forceinline__ int dbk_st(int cls, int q)
{
    const int f = cls * 52 + q;
    return (q < 0 || f >= 4 * 52) ? 0 : c_df_st[f];
}

// xevdm_get_tbl_qp_to_st (xevdm_df.c:38-104): 0 intra, 1 luma cbf, 2 motion differs (or IBC), 3 no filtering
__device__ __forceinline__ int dbk_class(const DbkArgs &a, int cur, int nb)
{
    const uint32_t m0 = a.map_scu[cur], m1 = a.map_scu[nb];
    if (((m0 | m1) >> 15) & 1) return 0;
    if (((m0 | m1) >> 24) & 1) return 1;
    if (((m0 | m1) >> 26) & 1) return 2;
    const int16_t r0 = ((const int16_t *)a.map_refi)[cur], r1 = ((const int16_t *)a.map_refi)[nb];
    const int8_t r00 = (int8_t)(r0 & 0xff), r01 = (int8_t)(r0 >> 8), r10 = (int8_t)(r1 & 0xff), r11 = (int8_t)(r1 >> 8);
    const int2 v0 = ((const int2 *)a.map_mv)[cur], v1 = ((const int2 *)a.map_mv)[nb];
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

// strengths of one segment: luma, Cb, Cr.  QP is the one of the CURRENT (right / lower) SCU only (xevd_df.c:347,446; T7)
__device__ __forceinline__ void dbk_strengths(const DbkArgs &a, int cur, int nb, int &st, int &st_u, int &st_v)
{
    const int cls = dbk_class(a, cur, nb);
    const int qp = (a.map_scu[cur] >> 16) & 0x7f;
    st = dbk_st(cls, qp) << (a.bd_l - 8);
    const int qu = xb_clip3(-6 * (a.bd_c - 8), 57, qp + a.qp_u_offset), qv = xb_clip3(-6 * (a.bd_c - 8), 57, qp + a.qp_v_offset);
    st_u = dbk_st(cls, qu < 0 ? qu : a.cq[0][qu]) << (a.bd_c - 8);
    st_v = dbk_st(cls, qv < 0 ? qv : a.cq[1][qv]) << (a.bd_c - 8);
}

template <bool VERTICAL>
__global__ void __launch_bounds__(256) k_deblock(const __grid_constant__ DbkArgs a)
{
    const int sx = blockIdx.x * 32 + (threadIdx.x & 31), sy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (sx >= a.w_scu || sy >= a.h_scu) return;
    if (!dbk_has_edge(a, sx, sy, VERTICAL)) return;
    const int cur = sy * a.w_scu + sx, nb = VERTICAL ? cur - 1 : cur - a.w_scu;
    int st, st_u, st_v;
    dbk_strengths(a, cur, nb, st, st_u, st_v);
    const int maxl = (1 << a.bd_l) - 1, maxc = (1 << a.bd_c) - 1;
    if (st) {
        pel *p = a.y + (size_t)(sy * 4) * a.s_l + sx * 4;
        if (VERTICAL) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                pel *q = p + (size_t)i * a.s_l;
                const int lo = *(const int *)(q - 2), hi = *(const int *)q;           // A B | C D
                int A = (int16_t)(lo & 0xffff), B = lo >> 16, C = (int16_t)(hi & 0xffff), D = hi >> 16;
                dbk_luma(A, B, C, D, st, maxl);
                *(int *)(q - 2) = (A & 0xffff) | (B << 16);
                *(int *)q = (C & 0xffff) | (D << 16);
            }
        } else {
            int r[4][4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int2 v = *(const int2 *)(p + (ptrdiff_t)(j - 2) * a.s_l);
                r[j][0] = (int16_t)(v.x & 0xffff); r[j][1] = v.x >> 16; r[j][2] = (int16_t)(v.y & 0xffff); r[j][3] = v.y >> 16;
            }
#pragma unroll
            for (int i = 0; i < 4; i++) dbk_luma(r[0][i], r[1][i], r[2][i], r[3][i], st, maxl);
#pragma unroll
            for (int j = 0; j < 4; j++)
                *(int2 *)(p + (ptrdiff_t)(j - 2) * a.s_l) = make_int2((r[j][0] & 0xffff) | (r[j][1] << 16), (r[j][2] & 0xffff) | (r[j][3] << 16));
        }
    }
    // chroma: only the head of a run of consecutive segments works; it walks the run in the reference's order
    const int psx = VERTICAL ? sx - 1 : sx, psy = VERTICAL ? sy : sy - 1;
    if (dbk_has_edge(a, psx, psy, VERTICAL)) return;
    int cx = sx, cy = sy;
    while (true) {
        const int c2 = cy * a.w_scu + cx, n2 = VERTICAL ? c2 - 1 : c2 - a.w_scu;
        if (c2 != cur) dbk_strengths(a, c2, n2, st, st_u, st_v);
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
        if (VERTICAL) { cx++; if (cx >= a.w_scu) break; } else { cy++; if (cy >= a.h_scu) break; }
        if (!dbk_has_edge(a, cx, cy, VERTICAL)) break;
    }
}

inline void launch_deblock(const DbkArgs &a, cudaStream_t st)
{
    const dim3 grid((a.w_scu + 31) / 32, (a.h_scu + 7) / 8);
    k_deblock<true><<<grid, 256, 0, st>>>(a);
    k_deblock<false><<<grid, 256, 0, st>>>(a);
}

}  // namespace xb
