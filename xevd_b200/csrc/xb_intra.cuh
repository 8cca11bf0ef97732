// xb_intra.cuh -- intra CUs: CTU wavefront kernel.
//
// Replaces the intra branch of xevd_recon_unit (src_base/xevd.c:732-741): get_nbr_yuv -> xevd_get_nbr_b
// (src_base/xevd_ipred.c:33-93), xevd_ipred_b / xevd_ipred_uv_b (src_base/xevd_ipred.c:95-160,586-676), residual
// (xevd_sub_block_itdq) and xevd_recon.
//
// Intra CUs read reconstructed samples of their left / upper neighbours, so they cannot run in the fully parallel inter
// kernel.  The reference serialises them with a CTU-row wavefront (sync_flag, src_base/xevd.c:1497-1501); the same
// dependency structure is used here: one CTA per CTU, CTUs are handed out in wavefront order (x + 2y) by an atomic ticket (so
// every CTA a waiter depends on has already started), a CTA spins on the `done` flags of those of its left, upper-left, upper
// and upper-right CTUs it really depends on (without IBC: only where a neighbour's intra / HTDF-filtered CUs lie under the
// samples its own border CUs read), then reconstructs its intra CUs one after the other in decoding order, all threads
// cooperating on each CU.  Inter CUs of the picture - and the residual of the intra CUs - have been done by the inter kernel
// before this one starts.
// Neighbour availability is order-derived in the reference (COD bits); it arrives precomputed as the per-SCU masks of
// XB200_CU_EXT.
#pragma once
#include "xb_common.cuh"
#include "xb_itdq.cuh"
#include "xb_recon.cuh"

namespace xb {

constexpr int kIntraThreads = 256;

// Shared memory of the wavefront kernel.  The serial chain CU -> CU is the critical path of an I picture, so everything a CU needs
// is kept on chip: the residual of every wavefront CU of the CTU (computed BEFORE the CTA waits for its neighbour CTUs - it does
// not depend on them), the CTU's reconstructed samples (inter CUs preloaded from the picture, wavefront CUs written as they
// finish, mirrored to global memory), and the row above / column left of the CTU fetched once after the wait.
struct IntraSmem {
    // CTU-raster samples, luma then Cb, Cr (plane stride = plane CTU size).  Sized by the picture's CTU: with 64x64 CTUs the two arrays take
    // 24 KB instead of the 96 KB of 128x128 CTUs, the CTA 91 KB instead of 163 KB, and two CTAs share an SM (118 registers allow it) - the
    // kernel's phases are chains of global-memory round trips that one CTA of eight warps cannot hide.
    static constexpr int plane_elems(int log2_ctu) { return (1 << (2 * log2_ctu)) * 3 / 2; }
    static constexpr int kTopElems = 2 * 128 + 8;                   // per plane: x = -1 .. 2 * S - 1 of the row above
    static constexpr int kLeftElems = 128;                          // per plane: the column left of the CTU
    static constexpr int kTmpElems = 66 * 66 / 2 + 65 * 65 * 2 + 16; // HTDF: ring-extended copy of the CU (int16) + four outputs per 2x2 window (4 x int16)
    static constexpr int kNbElems = 3 * (2 * 128 + 8);              // up[-1..w+h), left[-1..w+h), right[-1..w+h)
    static constexpr int kCuStage = 256;                            // CU descriptors + extension records staged on chip (the rest stay in L2)
    static size_t bytes(int log2_ctu)
    {
        return sizeof(int16_t) * (2 * plane_elems(log2_ctu) + 3 * kTopElems + 3 * kLeftElems + 3 * kNbElems) + sizeof(int) * kTmpElems +
               kCuStage * (sizeof(XB200_CU) + sizeof(XB200_CU_EXT)) + 64;
    }
};

// one plane of the current CTU: on-chip samples + the global picture they mirror
struct PlaneCtx {
    int16_t *rec;       // [Sp][Sp] reconstructed samples of the CTU
    int16_t *top;       // top[x], x = -1 .. 2 * Sp - 1: the row above the CTU
    int16_t *left;      // left[y], y = 0 .. Sp - 1: the column left of the CTU
    int16_t *res;       // [Sp][Sp] residual of the wavefront CUs
    pel *g;             // global plane at the CTU origin
    int gs, Sp;
    // X, Y relative to the CTU; only positions the availability rules allow are ever asked for
    __device__ __forceinline__ int get(int X, int Y) const { return Y < 0 ? top[X] : (X < 0 ? left[Y] : rec[Y * Sp + X]); }
    __device__ __forceinline__ void put(int X, int Y, int v) const { rec[Y * Sp + X] = (int16_t)v; g[(size_t)Y * gs + X] = (pel)v; }
};

// residual of one plane of one CU (all threads of the CTA): blocks larger than 64 (chroma: 32) are cut into sub-blocks gated
// by the nnz_sub bits; the result lands CU-raster in `res` (stride = plane width)
template <bool IQT>
__device__ void cu_plane_residual(const int16_t *__restrict__ coef, int lw, int lh, int lmax, int bits, int qp, int bd,
                                  int16_t *res, int rs, int *tmp, int tid, int nthreads, int ats = -1)
{
    const int pw = 1 << lw, ph = 1 << lh;
    for (int i = tid; i < pw * ph; i += nthreads) res[(i >> lw) * rs + (i & (pw - 1))] = 0;
    const int slw = min(lw, lmax), slh = min(lh, lmax);
    const int nx = 1 << (lw - slw), ny = 1 << (lh - slh);
    const int w = 1 << slw, h = 1 << slh, ts = w + 1;
    Dequant dq;
    dq.init(slw, slh, qp, bd, IQT);
    const int sh1 = IQT ? 7 : 0, sh2 = IQT ? 12 - (bd - 8) : 19 - (bd - 8);
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            if (!((bits >> ((j << 1) | i)) & 1)) continue;       // uniform across the CTA
            const int16_t *src = coef + (j << slh) * pw + (i << slw);
            __syncthreads();
            for (int x = tid; x < w; x += nthreads) {
                int *dst = tmp + x;
                if (ats >= 0) ats_line_dyn(slh, ats & 1, [&](int k) { return dq.apply(src[k * pw + x]); }, [&](int n, int v) { dst[n * ts] = v; }, 7);
                else itx_line_dyn<IQT>(slh, [&](int k) { return dq.apply(src[k * pw + x]); }, [&](int n, int v) { dst[n * ts] = v; }, sh1);
            }
            __syncthreads();
            for (int y = tid; y < h; y += nthreads) {
                const int *srow = tmp + y * ts;
                int16_t *drow = res + ((j << slh) + y) * rs + (i << slw);
                if (ats >= 0) ats_line_dyn(slw, ats >> 1, [&](int k) { return srow[k]; }, [&](int n, int v) { drow[n] = (int16_t)v; }, 20 - bd);
                else itx_line_dyn<false>(slw, [&](int k) { return srow[k]; }, [&](int n, int v) { drow[n] = (int16_t)v; }, sh2);
            }
        }
    __syncthreads();
}

// xevd_get_nbr_b for one plane: up[-1 .. w+h), left[-1 .. h+w); unit = samples per SCU (4 luma, 2 chroma)
// (cx, cy): position of the CU inside the CTU plane
// ---- Main profile (tool_eipd) ------------------------------------------------------------------------------------------
// xevdm_get_nbr (src_main/xevdm_ipred.c:39-150): unavailable units repeat the last filled sample, scanning away from the
// corner.  In closed form: the value at unit k is the sample itself when the unit is available, else the last sample of the
// nearest available unit before it, else the scan's start value.  The corner up[-1] becomes up[0] when the up-left unit is
// unavailable (the reference's loop over the units left of the corner overwrites it, :85-104).
struct NbSrc {
    PlaneCtx pc;
    int cx, cy;                 // CU position inside the CTU plane
    int w, ush, dflt;
    unsigned long long um, lm, rm;
    bool ul;
    __device__ __forceinline__ int up(int i) const
    {
        const int k = i >> ush;
        if ((um >> k) & 1) return pc.get(cx + i, cy - 1);
        const unsigned long long m = um & ((1ull << k) - 1);
        if (m) return pc.get(cx + (((63 - __clzll((long long)m)) + 1) << ush) - 1, cy - 1);
        return ul ? pc.get(cx - 1, cy - 1) : dflt;
    }
    __device__ __forceinline__ int corner() const { return ul ? pc.get(cx - 1, cy - 1) : up(0); }
    __device__ __forceinline__ int side(int i, unsigned long long mask, int col, int start) const
    {
        const int k = i >> ush;
        if ((mask >> k) & 1) return pc.get(cx + col, cy + i);
        const unsigned long long m = mask & ((1ull << k) - 1);
        if (m) return pc.get(cx + col, cy + (((63 - __clzll((long long)m)) + 1) << ush) - 1);
        return start;
    }
};

// element i of the 3 * (w + h) + 3 neighbour samples of one plane
__device__ __forceinline__ void intra_gather_main_elem(const NbSrc &nb, int n, int16_t *up, int16_t *left, int16_t *right, int i)
{
    if (i < n) up[i] = (int16_t)nb.up(i);
    else if (i < 2 * n) left[i - n] = (int16_t)nb.side(i - n, nb.lm, -1, nb.corner());
    else if (i < 3 * n) right[i - 2 * n] = (int16_t)nb.side(i - 2 * n, nb.rm, nb.w, nb.up(nb.w));
    else if (i == 3 * n) up[-1] = (int16_t)nb.corner();
    else if (i == 3 * n + 1) left[-1] = (int16_t)nb.corner();
    else right[-1] = (int16_t)nb.up(nb.w);
}

// plane / bilinear mode tables of xevdm_ipred.c (kept out of local memory: a run-time index into a function-local array puts it on the stack)
__constant__ int c_ipred_pl_mult[6] = {13, 17, 5, 11, 23, 47}, c_ipred_pl_shift[6] = {7, 10, 11, 15, 19, 23}, c_ipred_bi_wc[6] = {-1, 341, 205, 114, 60, 31};
__constant__ int c_inv_size_plus1[8] = {2048, 1365, 819, 455, 241, 124, 63, 32};          // xevd_ipred.c:108
__constant__ short c_ipred_dxdy[33][2] = {                                                   // xevd_tbl_ipred_dxdy, xevd_tbl.c:294-304
    {0, 0}, {0, 0}, {0, 0}, {2816, 372}, {2048, 512}, {1408, 744}, {1024, 1024}, {744, 1408}, {512, 2048}, {372, 2816}, {256, 4096},
    {128, 8192}, {0, 0}, {128, 8192}, {256, 4096}, {372, 2816}, {512, 2048}, {744, 1408}, {1024, 1024}, {1408, 744}, {2048, 512},
    {2816, 372}, {4096, 256}, {8192, 128}, {0, 0}, {8192, 128}, {4096, 256}, {2816, 372}, {2048, 512}, {1408, 744}, {1024, 1024},
    {744, 1408}, {512, 2048}};

// one angular sample: ipred_ang_val (src_base/xevd_ipred.c:377-570); 4-tap filter {32-f, 64-f, 32+f, f}, positions clamped to [-1, w+h-1]
__device__ __forceinline__ int intra_ang_px(const int16_t *up, const int16_t *le, const int16_t *ri, int lr, int ipm, int i, int j, int w, int h, int maxv)
{
    const int mdx = c_ipred_dxdy[ipm][0], mdy = c_ipred_dxdy[ipm][1];
    const bool right_ok = lr >= 2;
    const int dxy = (ipm > 24 || ipm < 12) ? -1 : 1;
    const int16_t *src;
    int pos, frac, step;
    auto proj = [&](int m, int d) { const int t = d * m; const int whole = t >> 10; frac = (t >> 5) - (whole << 5); return whole; };
    if (ipm < 12) {
        const int t = proj(mdx, j + 1);
        if (right_ok && i >= w - t) { const int ty = proj(mdy, w - i); src = ri; pos = j - ty; step = dxy > 0 ? 1 : -1; }
        else { src = up; pos = i + t; step = dxy < 0 ? 1 : -1; }
    } else if (ipm > 24) {
        if (right_ok) {
            const int ty = proj(mdy, w - i);
            if (j < ty) { const int t = proj(mdx, w - i); src = up; pos = i + t; step = dxy < 0 ? 1 : -1; }
            else { src = ri; pos = j - ty; step = dxy > 0 ? 1 : -1; }
        } else { const int ty = proj(mdy, i + 1); src = le; pos = j + ty; step = dxy < 0 ? 1 : -1; }
    } else {
        const int ty = proj(mdy, i + 1);
        if (j < ty) { const int t = proj(mdx, j + 1); src = up; pos = i - t; step = dxy < 0 ? 1 : -1; }
        else if (lr == 2) { const int ty2 = proj(mdy, w - i); src = ri; pos = j + ty2; step = dxy > 0 ? 1 : -1; }
        else { src = le; pos = j - ty; step = dxy < 0 ? 1 : -1; }
    }
    const int hi = w + h - 1;
    const int p0 = xb_clip3(-1, hi, pos - step), p1 = xb_clip3(-1, hi, pos), p2 = xb_clip3(-1, hi, pos + step), p3 = xb_clip3(-1, hi, pos + 2 * step);
    const int v = (int16_t)((src[p0] * (32 - frac) + src[p1] * (64 - frac) + src[p2] * (32 + frac) + src[p3] * frac + 64) >> 7);
    return xb_clip3(0, maxv, v);
}

// xevdm_ipred / xevdm_ipred_uv for one plane (src_main/xevdm_ipred.c:153-305; shared predictors src_base/xevd_ipred.c:110-372).
// ipm: 0 DC, 1 planar, 2 bilinear, 12 vertical, 24 horizontal, others angular.  Same split as above: mode scalars by one warp per plane
// (scr[0..2]), then a per-sample function.  pmax clips the predictor (plane bit depth).
__device__ __forceinline__ void intra_scalars_main(int w, int h, int lw, int lh, int ipm, int lr, const int16_t *up, const int16_t *le, const int16_t *ri,
                                                   int *scr, int lane)
{
    if (ipm == 0) {
        int acc = 0;
        for (int i = lane; i < w; i += 32) acc += up[i];
        for (int i = lane; i < h; i += 32) acc += (lr == 3 ? le[i] + ri[i] : (lr == 2 ? ri[i] : le[i]));
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (lane == 0) {
            const int lhh = lr == 3 ? lh + 1 : lh;
            acc += lr == 3 ? (w + h + h) >> 1 : (w + h) >> 1;
            scr[0] = (acc * c_inv_size_plus1[lw > lhh ? lw - lhh : lhh - lw]) >> (min(lw, lhh) + 12);
        }
    } else if (ipm == 1) {
        const bool fr = lr >= 2;
        const int16_t *sd = fr ? ri : le;
        const int w2 = w >> 1, h2 = h >> 1;
        int ch = 0, cv = 0;
        for (int x = 1 + lane; x <= w2; x += 32) ch += fr ? x * (up[w2 - x] - up[w2 + x]) : x * (up[w2 - 1 + x] - up[w2 - 1 - x]);
        for (int y = 1 + lane; y <= h2; y += 32) cv += y * (sd[h2 - 1 + y] - sd[h2 - 1 - y]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { ch += __shfl_xor_sync(0xffffffffu, ch, d); cv += __shfl_xor_sync(0xffffffffu, cv, d); }
        if (lane == 0) {
            const int iw = lw < 2 ? 0 : lw - 2, ih = lh < 2 ? 0 : lh - 2;
            const int a = (sd[h - 1] + (fr ? up[0] : up[w - 1])) << 4;
            const int b = ((ch << 5) * c_ipred_pl_mult[iw] + (1 << (c_ipred_pl_shift[iw] - 1))) >> c_ipred_pl_shift[iw];
            const int c = ((cv << 5) * c_ipred_pl_mult[ih] + (1 << (c_ipred_pl_shift[ih] - 1))) >> c_ipred_pl_shift[ih];
            scr[0] = a - (h2 - 1) * c - (w2 - 1) * b + 16; scr[1] = b; scr[2] = c;
        }
    } else if (ipm == 2 && lane == 0 && lr != 3) {
        const bool fr = lr == 2;
        const int a = fr ? up[-1] : up[w], b = fr ? ri[h] : le[h];
        const int lmin = min(lw, lh);
        const int c = w == h ? (a + b + 1) >> 1 : (((a << lw) + (b << lh)) * c_ipred_bi_wc[lw > lh ? lw - lh : lh - lw] + (1 << (lmin + 9))) >> (lmin + 10);
        scr[0] = a; scr[1] = b; scr[2] = (c << 1) - a - b;
    }
}
__device__ __forceinline__ int intra_px_main(int ipm, int lr, int x, int y, int w, int h, int lw, int lh, const int16_t *up, const int16_t *le,
                                             const int16_t *ri, const int *scr, int pmax)
{
    const int mul_w = c_inv_size_plus1[lw];
    if (ipm == 12) return up[x];
    if (ipm == 24) return lr == 3 ? (int16_t)(((le[y] * (w - x) + ri[y] * (x + 1) + (w >> 1)) * mul_w) >> 12) : (lr == 2 ? ri[y] : le[y]);
    if (ipm == 0) return scr[0];
    if (ipm == 1) return xb_clip3(0, pmax, (scr[0] + y * scr[2] + (lr >= 2 ? w - 1 - x : x) * scr[1]) >> 5);
    if (ipm == 2) {
        if (lr == 3) {
            const int hb = (le[y] * (w - x) + ri[y] * (x + 1) + (w >> 1)) * mul_w >> 12;
            const int hl = (le[h - 1] * (w - x) + ri[h - 1] * (x + 1) + (w >> 1)) * mul_w >> 12;
            const int vb = (up[x] * (h - 1 - y) + hl * (y + 1) + (h >> 1)) >> lh;
            return (int16_t)((hb + vb + 1) >> 1);
        }
        const bool fr = lr == 2;
        const int sd = fr ? ri[y] : le[y], k = fr ? w - 1 - x : x;
        const int px = (sd << lw) + (k + 1) * (scr[0] - sd), py = (up[x] << lh) + (y + 1) * (scr[1] - up[x]);
        return xb_clip3(0, pmax, (int16_t)(((px << lh) + (py << lw) + k * y * scr[2] + (1 << (lw + lh))) >> (lw + lh + 1)));
    }
    return intra_ang_px(up, le, ri, lr, ipm, x, y, w, h, pmax);
}

// ---- HTDF (Main, tool_htdf): xevdm_htdf (src_main/xevdm_recon.c:153-385) ------------------------------------------------------------
// 2x2 Hadamard windows slide over the CU plus a one-sample ring (neighbours where xevd_get_avail_intra says so, else replicated;
// the row below is always replicated); the three AC terms are shrunk through a table chosen by the slice QP; every sample becomes
// the rounded average of its four windows.  All windows see unfiltered input (the reference overwrites a sample only after its last
// window), so the filter is a per-sample gather.
__constant__ uint8_t c_htdf_thr_log2[5] = {6, 7, 7, 8, 8};
__constant__ __align__(16) uint8_t c_htdf_tbl[5][16] = {
    {0, 0, 2, 6, 10, 14, 19, 23, 28, 32, 36, 41, 45, 49, 53, 57},       {0, 0, 5, 12, 20, 29, 38, 47, 56, 65, 73, 82, 90, 98, 107, 115},
    {0, 0, 1, 4, 9, 16, 24, 32, 41, 50, 59, 68, 77, 86, 94, 103},       {0, 0, 3, 9, 19, 32, 47, 64, 81, 99, 117, 135, 154, 179, 205, 230},
    {0, 0, 0, 2, 6, 11, 18, 27, 38, 51, 64, 96, 128, 160, 192, 224}};

// the 16-entry table row of the CU's QP lives in four registers: a per-lane index into __constant__ memory would be replayed once per
// distinct address, twelve times per sample
struct HtdfTbl { unsigned t0, t1, t2, t3; };
__device__ __forceinline__ int htdf_shrink(int z, const HtdfTbl &tbl, int thr, int shift, int round)
{
    const int av = abs(z);
    if (av >= thr) return z;
    const int idx = ((av + round) & thr) >> shift;
    const int v = (int)((idx < 8 ? __byte_perm(tbl.t0, tbl.t1, idx & 7) : __byte_perm(tbl.t2, tbl.t3, idx & 7)) & 0xff);
    return z < 0 ? -v : v;
}
// true when the CU is filtered (xevdm.c:1383 + xevdm_htdf_skip_condition, xevdm_recon.c:271-297); qp receives the table QP
__device__ __forceinline__ bool htdf_applies(const XbFrameArgs &a, const XB200_CU &cu, int &qp)
{
    if (!a.htdf || cu.mode == XB200_MODE_IBC || !(cu.flags & XB200_CUF_LUMA)) return false;
    const bool intra = cu.mode == XB200_MODE_INTRA;
    if (!intra && !(cu.cbf & 15)) return false;
    qp = a.slice_qp;
    const int w = 1 << cu.log2w, h = 1 << cu.log2h;
    if (qp <= 17 || w * h < 64 || max(w, h) >= 128) return false;
    if (!intra) return min(w, h) < 32;
    if (w == h && w >= 32) qp -= 8;
    return true;
}

// ms: map_scu at the CU's first SCU when the ring takes the constrained-intra test (intra CU under pps.constrained_intra_pred_flag: left /
// right / upper samples only from intra neighbours, xevdm_recon.c:317,338,359), else null.  The maps of the whole picture are final here:
// the inter kernels publish them for every CU before this kernel starts.
__device__ __forceinline__ void cu_htdf(const XbFrameArgs &a, int log2w, int log2h, int av, const PlaneCtx pc, int cx, int cy, int qp, int16_t *t, int tid, int nthreads,
                                        const uint32_t *__restrict__ ms)
{
    const int w = 1 << log2w, h = 1 << log2h, we = w + 2;
    const bool up = av & 1, le = (av >> 1) & 1, ri = (av >> 3) & 1;
    // (1) the CU itself, straight from the on-chip CTU
    for (int idx = tid; idx < w * h; idx += nthreads) {
        const int y = idx >> log2w, x = idx & (w - 1);
        t[(y + 1) * we + x + 1] = pc.rec[(cy + y) * pc.Sp + cx + x];
    }
    //     and the one-sample ring: row above, row below, left and right columns
    for (int idx = tid; idx < 2 * we + 2 * h; idx += nthreads) {
        int i, j;
        if (idx < we) { i = -1; j = idx - 1; }
        else if (idx < 2 * we) { i = h; j = idx - we - 1; }
        else { const int k = idx - 2 * we; i = k >> 1; j = (k & 1) ? w : -1; }
        int si = min(max(i, 0), h - 1), sj = min(max(j, 0), w - 1);          // replicated by default
        auto nb_intra = [&](int off) -> bool { return !ms || ((__ldg(ms + off) >> 15) & 1); };          // MCU_GET_IF
        if (i >= 0 && i < h) {
            if (j < 0 && le) { if (nb_intra(-1 + (i >> 2) * a.w_scu)) sj = -1; }
            else if (j >= w && ri) { if (nb_intra((w >> 2) + (i >> 2) * a.w_scu)) sj = w; }
        }
        else if (i < 0) {
            if (j >= 0 && j < w) { if (up && nb_intra(-a.w_scu + (j >> 2))) si = -1; }
            else if (j < 0) { if ((av >> 5) & 1) { si = -1; sj = -1; } }
            else if ((av >> 6) & 1) { si = -1; sj = w; }
        } else if (j < 0) { if ((av >> 7) & 1) { si = h; sj = -1; } }
        else if (j >= w) { if ((av >> 8) & 1) { si = h; sj = w; } }
        t[(i + 1) * we + j + 1] = (int16_t)pc.get(cx + sj, cy + si);
    }
    __syncthreads();
    // (2) every 2x2 window once: Hadamard, shrink, inverse; its four outputs go to the four samples it covers
    int k = (qp - 20 + 4) >> 3;
    k = min(max(k, 0), 4);
    const int lg = c_htdf_thr_log2[k], shift = lg - 4, round = (1 << shift) >> 1, thr = (1 << lg) - (1 << shift);
    const uint4 tw = *(const uint4 *)c_htdf_tbl[k];
    const HtdfTbl tbl = {tw.x, tw.y, tw.z, tw.w};
    const int maxv = (1 << a.bd_l) - 1;
    const int ww = w + 1, nwin = ww * (h + 1);
    uint2 *win = (uint2 *)(t + ((we * (h + 2) + 3) & ~3));
    const float inv_ww = 1.0f / (float)ww;
    for (int idx = tid; idx < nwin; idx += nthreads) {
        const int i = __float2int_rd(((float)idx + 0.5f) * inv_ww), j = idx - i * ww;      // exact: idx < 4225, ww <= 65
        const int16_t *q = t + i * we + j;
        const int x0 = q[0], x1 = q[1], x2 = q[we], x3 = q[we + 1];
        const int y0 = x0 + x2, y1 = x1 + x3, y2 = x0 - x2, y3 = x1 - x3;
        const int z0 = y0 + y1;
        const int z1 = htdf_shrink(y0 - y1, tbl, thr, shift, round), z2 = htdf_shrink(y2 + y3, tbl, thr, shift, round), z3 = htdf_shrink(y2 - y3, tbl, thr, shift, round);
        const int i0 = z0 + z2, i1 = z1 + z3, i2 = z0 - z2, i3 = z1 - z3;
        const unsigned o0 = (unsigned)((i0 + i1) >> 2) & 0xffffu, o1 = (unsigned)((i0 - i1) >> 2) & 0xffffu;
        const unsigned o2 = (unsigned)((i2 + i3) >> 2) & 0xffffu, o3 = (unsigned)((i2 - i3) >> 2) & 0xffffu;
        win[idx] = make_uint2(o0 | (o1 << 16), o2 | (o3 << 16));
    }
    __syncthreads();
    // (3) a sample is the rounded average of the four windows that cover it (the reference accumulates in s16: sums are mod 2^16)
    for (int idx = tid; idx < w * h; idx += nthreads) {
        const int y = idx >> log2w, x = idx & (w - 1);
        const uint2 a00 = win[y * ww + x], a01 = win[y * ww + x + 1], a10 = win[(y + 1) * ww + x], a11 = win[(y + 1) * ww + x + 1];
        const int acc = (int16_t)((a00.y >> 16) + (a01.y & 0xffffu) + (a10.x >> 16) + (a11.x & 0xffffu));
        pc.put(cx + x, cy + y, xb_clip3(0, maxv, (acc + 2) >> 2));
    }
    __syncthreads();
}

struct IntraSync {
    int *ticket;        // next position of `order` to hand out
    int *err;           // sticky: set when a wait gave up (malformed work list: a dependency that never completes); read by xb200_sync
    int *done;          // [n_ctu] 1 when every CU of the CTU is final
    const int *order;   // CTU addresses sorted by wavefront index x + 2y: CTUs of one index are independent, so the CTAs resident at
                        // any time span many CTU rows (raster order kept only ~2.5 rows busy: 148 resident CTAs / 60 CTUs per row);
                        // every dependency has a smaller index, hence an earlier ticket -> no deadlock
};

template <bool IQT>
__global__ void __launch_bounds__(kIntraThreads, 2)      // two CTAs per SM with CTUs of 64 samples and less (IntraSmem): at most 128 registers
k_recon_intra(const __grid_constant__ XbFrameArgs a, const IntraSync sy)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // the arrays of fixed size first, at compile-time offsets (the CU chain is bound by its instruction count: run-time offsets for all of them
    // cost 9 % on an I picture); only the residual array behind the CTU-sized sample array sits at an offset computed from log2_ctu
    int *s_tmp = (int *)smem_raw;
    XB200_CU *s_cu = (XB200_CU *)(s_tmp + IntraSmem::kTmpElems);
    XB200_CU_EXT *s_ext = (XB200_CU_EXT *)(s_cu + IntraSmem::kCuStage);
    int16_t *s_top = (int16_t *)(s_ext + IntraSmem::kCuStage);
    int16_t *s_left = s_top + 3 * IntraSmem::kTopElems;
    int16_t *s_nb = s_left + 3 * IntraSmem::kLeftElems;
    int16_t *s_rec = s_nb + 3 * IntraSmem::kNbElems;
    int16_t *s_res = s_rec + IntraSmem::plane_elems(a.log2_ctu);
    __shared__ int s_ctu, s_scr12[12], s_inter_all;
    __shared__ unsigned s_req[4], s_has[4];
    const int tid = threadIdx.x;

    // Persistent CTAs: each takes CTU after CTU in wavefront order until the tickets run out.  The host sizes the grid - one CTA per CTU
    // when the dependencies are sparse (P / B pictures: most CTUs can start at once), about as many CTAs as the wavefront is wide when
    // they are dense (I pictures), so that the CTAs of one picture do not fill the device with waiters and several independent pictures
    // (other contexts / streams) can be in flight.  A dependency always has an earlier ticket, and a ticket is only ever held by a
    // resident CTA: no deadlock for any grid size.
    // A wait that does not end within ~2 s (a malformed work list) raises the sticky error word and lets the CTA go on: wrong pixels, no hang.
    auto wait_done = [&](int nc) {
        volatile int *f = sy.done + nc;
        unsigned spins = 0;
        while (*f == 0) { __nanosleep(64); if (++spins > (1u << 24)) { atomicExch(sy.err, 1); break; } }
    };
    // Overlap with the generic inter kernel (a.inter_done): this kernel then runs beside it on a second stream and must not read a CTU's
    // samples before that kernel is done with them - its own CTU before the preload, the four neighbours before their border row / column.
    // Picture samples are read with ld.global.cg throughout, so a line cached before another kernel rewrote it cannot be served.
    // Once the generic kernel has finished every CTU (a.inter_count) the CTA stops looking at flags (s_inter_all, written by thread 0 only).
    auto wait_inter = [&](int nc) {
        if (s_inter_all) return;
        volatile int *f = a.inter_done + nc;
        unsigned spins = 0;
        while (*f == 0) { __nanosleep(64); if (++spins > (1u << 24)) { atomicExch(sy.err, 1); break; } }
    };
    if (tid == 0) s_inter_all = a.inter_done ? 0 : 1;
    for (;;) {
    __syncthreads();                // every thread is done with the previous CTU's shared state
    if (tid == 0) s_ctu = atomicAdd(sy.ticket, 1);
    const int inter_all_seen = s_inter_all;      // read between barriers: thread 0 sets the word further down (a stale 0 only costs one more look at a.inter_count)
    __syncthreads();
    if (s_ctu >= a.n_ctu) return;
    const int ctu = sy.order[s_ctu];
    const int cx = ctu % a.w_ctu, cy = ctu / a.w_ctu;
    const int cu0 = a.ctu_first[ctu], cu1 = a.ctu_first[ctu + 1];
    const int S = 1 << a.log2_ctu, Sc = S >> 1;
    const int ctu_x = cx << a.log2_ctu, ctu_y = cy << a.log2_ctu;

    // any wavefront work here?  (uniform: every thread scans the same descriptors through L1)
    bool any = false;
    for (int i = cu0 + tid; i < cu1; i += kIntraThreads) { int q; any |= xb_wavefront_mode(a.cus[i].mode) || htdf_applies(a, a.cus[i], q); }
    any = __syncthreads_or(any);
    if (any) {
        // CU descriptors and extension records on chip: two dependent L2 round trips per CU would sit on the critical path otherwise
        const int n_stage = min(cu1 - cu0, IntraSmem::kCuStage);
        for (int i = tid; i < n_stage * 2; i += kIntraThreads) ((int4 *)s_cu)[i] = __ldg((const int4 *)(a.cus + cu0) + i);
        __syncthreads();
        for (int i = tid; i < n_stage * 2; i += kIntraThreads) {
            const XB200_CU &c = s_cu[i >> 1];
            if (c.mode == XB200_MODE_INTRA) { uint32_t ei; memcpy(&ei, c.mv[1], 4); ((int4 *)s_ext)[i] = __ldg((const int4 *)(a.ext + ei) + (i & 1)); }
        }
        __syncthreads();
        auto get_cu = [&](int i) -> XB200_CU { return i - cu0 < n_stage ? s_cu[i - cu0] : a.cus[i]; };
        // plane contexts and neighbour arrays are rebuilt from the plane index by arithmetic: arrays of them indexed by a run-time plane
        // would live in local memory, and every sample of the serial CU chain would pay its latency
        pel *const g_y = a.cur.y + (size_t)ctu_y * a.s_l + ctu_x;
        pel *const g_u = a.cur.u + (size_t)(ctu_y >> 1) * a.s_c + (ctu_x >> 1), *const g_v = a.cur.v + (size_t)(ctu_y >> 1) * a.s_c + (ctu_x >> 1);
        auto pc = [&](int pl) -> PlaneCtx {
            PlaneCtx c;
            const int po = pl == 0 ? 0 : (pl == 1 ? S * S : S * S + Sc * Sc);
            c.rec = s_rec + po; c.res = s_res + po; c.Sp = pl ? Sc : S;
            c.top = s_top + pl * IntraSmem::kTopElems + 4;
            c.left = s_left + pl * IntraSmem::kLeftElems;
            c.gs = pl ? a.s_c : a.s_l;
            c.g = pl == 0 ? g_y : (pl == 1 ? g_u : g_v);
            return c;
        };
        auto up = [&](int pl) -> int16_t * { return s_nb + pl * IntraSmem::kNbElems + 4; };
        auto le = [&](int pl) -> int16_t * { return s_nb + pl * IntraSmem::kNbElems + 4 + (2 * 128 + 8); };
        auto ri = [&](int pl) -> int16_t * { return s_nb + pl * IntraSmem::kNbElems + 4 + 2 * (2 * 128 + 8); };
        if (a.inter_done) {
            if (tid == 0 && !inter_all_seen) {       // (the register copy: a shared-memory read here is hoisted above the test of tid and races with the store below)
                if (*(volatile int *)a.inter_count >= a.n_ctu) s_inter_all = 1;
                else wait_inter(ctu);
                __threadfence();
            }
            __syncthreads();
        }
        // ---- before waiting: (1) the CTU's samples as the inter kernel left them (neighbours of intra CUs, HTDF input) -------------
        for (int pl = 0; pl < 3; pl++) {
            const int Sp = pc(pl).Sp, wv = min(Sp, ((a.w - ctu_x) >> (pl ? 1 : 0))), hv = min(Sp, ((a.h - ctu_y) >> (pl ? 1 : 0)));
            for (int i = tid; i < Sp * hv; i += kIntraThreads) {
                const int y = i / Sp, x = i - y * Sp;
                if (x < wv) pc(pl).rec[y * Sp + x] = pc(pl).res[y * Sp + x] = __ldcg(pc(pl).g + (size_t)y * pc(pl).gs + x);
            }
        }
        //      intra / IBC areas hold the RESIDUAL there (parked by the inter kernel), which is why the preload fills both arrays
        // ---- which of the left, upper-left, upper and upper-right CTU this one really depends on -----------------------------------
        // Without HTDF the wavefront kernel changes nothing but the samples of intra CUs, and an intra CU reads the row above
        // it (up[-1 .. w+h)) and the column left of it (left[-1 .. h+w)) only.  Inter CUs are final when this kernel starts, so a
        // neighbour CTU matters only where one of ITS intra CUs lies under those samples: per 4-sample unit, the columns of the
        // neighbour's bottom row / rows of its right column covered by intra CUs against the units this CTU's border CUs read.
        // In P/B pictures with scattered intra CUs most CTUs then start at once instead of queueing on a 128-step wavefront.
        // With HTDF the same holds for the CUs it filters: they read a one-sample ring around the CU (inside the extents above) and
        // are the only other samples this kernel changes - uncoded CUs break the chain.  IBC reaches arbitrarily far: full wavefront.
        const bool prune = !a.ibc;
        auto touched = [&](const XB200_CU &cu) { int q; return cu.mode == XB200_MODE_INTRA || htdf_applies(a, cu, q); };
        if (tid < 4) { s_req[tid] = 0; s_has[tid] = 0; }
        __syncthreads();
        if (prune) {
            const int n = S >> 2;
            auto bits = [](int lo, int hi) -> unsigned { return hi <= lo ? 0u : ((hi - lo >= 32 ? 0u : (1u << (hi - lo))) - 1u) << lo; };   // [lo, hi)
            for (int i = cu0 + tid; i < cu1; i += kIntraThreads) {
                const XB200_CU cu = get_cu(i);
                if (!touched(cu)) continue;
                const int X = (cu.x - ctu_x) >> 2, Y = (cu.y - ctu_y) >> 2, W = 1 << (cu.log2w - 2), H = 1 << (cu.log2h - 2);
                if (Y == 0) {                       // columns X-1 .. X+W+H of the row above
                    if (X == 0) atomicOr(&s_req[1], 1u);
                    atomicOr(&s_req[2], bits(max(X - 1, 0), min(X + W + H + 1, n)));
                    if (X + W + H + 1 > n) atomicOr(&s_req[3], bits(0, min(X + W + H + 1 - n, n)));
                }
                if (X == 0) {                       // rows Y-1 .. Y+H+W of the column to the left (below the CTU: never available)
                    if (Y == 0) atomicOr(&s_req[1], 1u);
                    atomicOr(&s_req[0], bits(max(Y - 1, 0), min(Y + H + W + 1, n)));
                }
            }
            for (int k = 0; k < 4; k++) {
                const int nx = cx + (k == 0 ? -1 : k - 2), ny = cy - (k == 0 ? 0 : 1);
                if (nx < 0 || nx >= a.w_ctu || ny < 0) continue;
                const int nc = ny * a.w_ctu + nx, ox = nx << a.log2_ctu, oy = ny << a.log2_ctu;
                for (int i = a.ctu_first[nc] + tid; i < a.ctu_first[nc + 1]; i += kIntraThreads) {
                    const XB200_CU cu = a.cus[i];
                    if (!touched(cu)) continue;
                    const int X = (cu.x - ox) >> 2, Y = (cu.y - oy) >> 2, W = 1 << (cu.log2w - 2), H = 1 << (cu.log2h - 2);
                    const bool at_right = X + W == n, at_bottom = Y + H == n;
                    if (k == 0) { if (at_right) atomicOr(&s_has[0], bits(Y, Y + H)); }
                    else if (k == 1) { if (at_right && at_bottom) atomicOr(&s_has[1], 1u); }
                    else if (at_bottom) atomicOr(&s_has[k], bits(X, X + W));
                }
            }
            __syncthreads();
        }
        // ---- wait for them ----------------------------------------------------------------------------------------------------------
        if (tid < 4) {
            const int nx = cx + (tid == 0 ? -1 : tid - 2), ny = cy - (tid == 0 ? 0 : 1);
            if (nx >= 0 && nx < a.w_ctu && ny >= 0) {
                if (a.inter_done) wait_inter(ny * a.w_ctu + nx);                  // its border samples are read below whether or not it has wavefront work
                if (!prune || (s_req[tid] & s_has[tid])) wait_done(ny * a.w_ctu + nx);
            }
            __threadfence();
        }
        __syncthreads();
        // ---- the row above and the column left of the CTU (L2 loads: other CTAs wrote them) ------------------------------------------
        for (int pl = 0; pl < 3; pl++) {
            const int sh = pl ? 1 : 0, Sp = pc(pl).Sp, wp = a.w >> sh, hp = a.h >> sh, x0 = ctu_x >> sh, y0 = ctu_y >> sh;
            for (int i = tid; i < 3 * Sp + 1; i += kIntraThreads) {
                if (i <= 2 * Sp) {                                                            // top[x], x = i - 1
                    const int x = i - 1;
                    if (y0 > 0 && x0 + x >= 0 && x0 + x < wp) pc(pl).top[x] = __ldcg(pc(pl).g - pc(pl).gs + x);
                } else {
                    const int y = i - 2 * Sp - 1;
                    if (x0 > 0 && y0 + y < hp) pc(pl).left[y] = __ldcg(pc(pl).g + (size_t)y * pc(pl).gs - 1);
                }
            }
        }
        __syncthreads();
        // ---- the CUs in decoding order --------------------------------------------------------------------------------------------
        const int dflt = 1 << (a.bd_l - 1), maxv = (1 << a.bd_l) - 1;
        for (int i = cu0; i < cu1; i++) {
            const XB200_CU cu = get_cu(i);
            int hq = 0;
            const bool do_htdf = htdf_applies(a, cu, hq);            // uniform
            const int lx = cu.x - ctu_x, ly = cu.y - ctu_y;
            // inter CUs are reconstructed by the inter kernel; only the in-order HTDF pass below is left for them
            if (xb_wavefront_mode(cu.mode)) do {
            const int w = 1 << cu.log2w, h = 1 << cu.log2h, cw = w >> 1, ch = h >> 1;
            // local dual tree: a TREE_L leaf carries luma only, the TREE_C CU after its siblings the chroma of the whole node; every
            // per-plane step of xevd_recon_unit is gated by xevd_check_luma / xevd_check_chroma (xevdm.c:611-640,1344-1391)
            const bool do_l = cu.flags & XB200_CUF_LUMA, do_c = cu.flags & XB200_CUF_CHROMA;          // uniform
            if (cu.mode == XB200_MODE_IBC) {
                // xevdm_IBC_mc (src_main/xevdm_mc.c:2040-2106): whole-sample copy from the already reconstructed part of the CURRENT
                // picture (block vector mv[0], chroma vector = luma >> 1), then xevdm_recon.  Conforming vectors stay inside the
                // current CTU row at or left of this CTU: samples of this CTU come from shared memory, older ones from the picture.
                const int bx = cu.mv[0][0], by = cu.mv[0][1];
                if (!s_inter_all) {         // (uniform) beside the generic kernel: the CTUs under the source block must have their inter CUs
                    if (tid == 0) {
                        const int X0 = max(cu.x + bx, 0), Y0 = max(cu.y + by, 0);
                        for (int yy = Y0 >> a.log2_ctu; yy <= min(Y0 + h - 1, a.h - 1) >> a.log2_ctu; yy++)
                            for (int xx = X0 >> a.log2_ctu; xx <= min(X0 + w - 1, a.w - 1) >> a.log2_ctu; xx++) wait_inter(yy * a.w_ctu + xx);
                        __threadfence();
                    }
                    __syncthreads();
                }
                for (int pl = do_l ? 0 : 1; pl < (do_c ? 3 : 1); pl++) {
                    const int sh = pl ? 1 : 0, pw = w >> sh, ph = h >> sh, lwp = cu.log2w - sh, Sp = pc(pl).Sp;
                    const int ox = lx >> sh, oy = ly >> sh, vx = bx >> sh, vy = by >> sh;
                    const bool coded = ((cu.cbf >> (4 * pl)) & 15) != 0;
                    // the source block is decoded earlier, so it cannot overlap this CU
                    for (int k = tid; k < pw * ph; k += kIntraThreads) {
                        const int y = k >> lwp, x = k & (pw - 1), X = ox + x + vx, Y = oy + y + vy;
                        const int p = (X >= 0 && X < Sp && Y >= 0 && Y < Sp) ? pc(pl).rec[Y * Sp + X] : __ldcg(pc(pl).g + (ptrdiff_t)Y * pc(pl).gs + X);
                        pc(pl).put(ox + x, oy + y, xb_clip3(0, maxv, (int16_t)(p + (coded ? pc(pl).res[(oy + y) * Sp + ox + x] : 0))));
                    }
                }
                __syncthreads();
                break;
            }
            uint32_t ei;
            memcpy(&ei, cu.mv[1], 4);
            const XB200_CU_EXT ex = i - cu0 < n_stage ? s_ext[i - cu0] : a.ext[ei];
            const bool ul = (cu.avail >> 2) & 1;
            // neighbours of all three planes, then prediction + reconstruction
            if (a.eipd) {
                const int lr = cu.avail & 3;
                const int pmax_c = (1 << a.bd_c) - 1;
                static const int8_t kChromaToLuma[5] = {-1, 2, 0, 24, 12};       // IPD_BI_C, DC_C, HOR_C, VER_C -> luma mode ids (xevdm_ipred.c:267-305)
                const int ipm_c = cu.refi[1] == 0 ? cu.refi[0] : kChromaToLuma[cu.refi[1]];
                {
                    const int n0 = 3 * (w + h) + 3, n1 = 3 * (cw + ch) + 3;
                    for (int k = (do_l ? 0 : n0) + tid; k < (do_c ? n0 + 2 * n1 : n0); k += kIntraThreads) {
                        const int pl = k < n0 ? 0 : (k < n0 + n1 ? 1 : 2);
                        NbSrc nb;
                        nb.pc = pc(pl); nb.cx = lx >> (pl ? 1 : 0); nb.cy = ly >> (pl ? 1 : 0);
                        nb.w = pl ? cw : w; nb.ush = pl ? 1 : 2; nb.dflt = dflt;
                        nb.um = ex.u.intra.up; nb.lm = ex.u.intra.left; nb.rm = ex.u.intra.right; nb.ul = ul;
                        intra_gather_main_elem(nb, pl ? cw + ch : w + h, up(pl), le(pl), ri(pl), k - (pl == 0 ? 0 : (pl == 1 ? n0 : n0 + n1)));
                    }
                }
                __syncthreads();
                if (cu.refi[0] <= 2 || ipm_c <= 2) {          // DC / plane / bilinear: per-plane scalars first (angular modes need none)
                    if ((tid >> 5) < 3 && ((tid >> 5) ? do_c : do_l)) {
                        const int pl = tid >> 5;
                        intra_scalars_main(pl ? cw : w, pl ? ch : h, cu.log2w - (pl ? 1 : 0), cu.log2h - (pl ? 1 : 0), pl ? ipm_c : cu.refi[0], lr, up(pl), le(pl), ri(pl),
                                           s_scr12 + 4 * pl, tid & 31);
                    }
                    __syncthreads();
                }
                for (int k = (do_l ? 0 : w * h) + tid; k < (do_c ? w * h + 2 * cw * ch : w * h); k += kIntraThreads) {
                    const int pl = k < w * h ? 0 : (k < w * h + cw * ch ? 1 : 2), kk = k - (pl == 0 ? 0 : (pl == 1 ? w * h : w * h + cw * ch));
                    const int lwp = cu.log2w - (pl ? 1 : 0), lhp = cu.log2h - (pl ? 1 : 0), wp = 1 << lwp, hp = 1 << lhp;
                    const int y = kk >> lwp, x = kk & (wp - 1), ox = lx >> (pl ? 1 : 0), oy = ly >> (pl ? 1 : 0);
                    const int p = intra_px_main(pl ? ipm_c : cu.refi[0], lr, x, y, wp, hp, lwp, lhp, up(pl), le(pl), ri(pl), s_scr12 + 4 * pl, pl ? pmax_c : maxv);
                    const int r = ((cu.cbf >> (4 * pl)) & 15) ? pc(pl).res[(oy + y) * pc(pl).Sp + ox + x] : 0;
                    pc(pl).put(ox + x, oy + y, xb_clip3(0, maxv, (int16_t)(p + r)));
                }
                __syncthreads();
                break;
            }
            {
                // Baseline modes (xevd_get_nbr_b + xevd_ipred_b): every sample is predicted straight from the on-chip neighbours - no
                // gathered copy, no barrier before the samples are written (neighbours lie outside the CU, writes inside it).  The DC
                // value is a warp-local reduction that every warp does for itself.  One barrier per CU instead of three.
                const unsigned long long um = ex.u.intra.up, lm = ex.u.intra.left;
                const int lane = tid & 31;
                int dc0 = 0, dc1 = 0, dc2 = 0;
#pragma unroll
                for (int pl = 0; pl < 3; pl++) {
                    if ((pl ? cu.refi[1] : cu.refi[0]) != 0 || !(pl ? do_c : do_l)) continue;            // uniform
                    const int sh = pl ? 1 : 0, wp = w >> sh, hp = h >> sh, cxp = lx >> sh, cyp = ly >> sh, ush = 2 - sh;
                    const PlaneCtx c = pc(pl);
                    int acc = 0;
                    for (int i = lane; i < hp; i += 32) acc += ((lm >> (i >> ush)) & 1) ? c.get(cxp - 1, cyp + i) : dflt;
                    for (int i = lane; i < wp; i += 32) acc += ((um >> (i >> ush)) & 1) ? c.get(cxp + i, cyp - 1) : dflt;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
                    const int v = (acc + wp) >> (cu.log2w - sh + 1);
                    if (pl == 0) dc0 = v; else if (pl == 1) dc1 = v; else dc2 = v;
                }
                for (int k = (do_l ? 0 : w * h) + tid; k < (do_c ? w * h + 2 * cw * ch : w * h); k += kIntraThreads) {
                    const int pl = k < w * h ? 0 : (k < w * h + cw * ch ? 1 : 2), kk = k - (pl == 0 ? 0 : (pl == 1 ? w * h : w * h + cw * ch));
                    const int sh = pl ? 1 : 0, lwp = cu.log2w - sh, wp = 1 << lwp, ush = 2 - sh;
                    const int y = kk >> lwp, x = kk & (wp - 1), ox = lx >> sh, oy = ly >> sh;
                    const PlaneCtx c = pc(pl);
                    auto nb_up = [&](int i) -> int { return ((um >> (i >> ush)) & 1) ? c.get(ox + i, oy - 1) : dflt; };
                    auto nb_le = [&](int i) -> int { return ((lm >> (i >> ush)) & 1) ? c.get(ox - 1, oy + i) : dflt; };
                    int p;
                    switch (pl ? cu.refi[1] : cu.refi[0]) {
                    case 0: p = pl == 0 ? dc0 : (pl == 1 ? dc1 : dc2); break;
                    case 1: p = nb_le(y); break;
                    case 2: p = nb_up(x); break;
                    case 3: p = y > x ? nb_le(y - x - 1) : (y == x ? (ul ? c.get(ox - 1, oy - 1) : dflt) : nb_up(x - y - 1)); break;
                    default: p = (nb_up(x + y + 1) + nb_le(x + y + 1)) >> 1; break;
                    }
                    const int r = ((cu.cbf >> (4 * pl)) & 15) ? c.res[(oy + y) * c.Sp + ox + x] : 0;
                    c.put(ox + x, oy + y, xb_clip3(0, maxv, (int16_t)(p + r)));          // xevd_recon: s16 wrap, then clip
                }
            }
            __syncthreads();         // the next CU reads these samples from shared memory
            } while (0);
            // one call site: the routine is not inlined, and a second copy of its operands would live in local memory
            if (do_htdf && cu.mode != XB200_MODE_IBC) cu_htdf(a, cu.log2w, cu.log2h, cu.avail_cu, pc(0), lx, ly, hq, (int16_t *)s_tmp, tid, kIntraThreads,
                                                               (a.constrained && cu.mode == XB200_MODE_INTRA) ? a.map_scu + (cu.y >> 2) * a.w_scu + (cu.x >> 2) : nullptr);
        }
    }
    else if (a.ibc && cx > 0) {
        // a CTU without wavefront work still hands on its left neighbour's completion: an IBC vector may reach any CTU to the left in the
        // row, and the CTUs in between must not report `done` before those are final (the waits are transitive along the row)
        if (tid == 0) wait_done(ctu - 1);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(sy.done + ctu, 1);
    }
}

}  // namespace xb
