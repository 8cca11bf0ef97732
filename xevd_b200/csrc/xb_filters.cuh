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
