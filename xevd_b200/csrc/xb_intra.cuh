// xb_intra.cuh -- intra CUs: CTU wavefront kernel.
//
// Replaces the intra branch of xevd_recon_unit (src_base/xevd.c:732-741): get_nbr_yuv -> xevd_get_nbr_b
// (src_base/xevd_ipred.c:33-93), xevd_ipred_b / xevd_ipred_uv_b (src_base/xevd_ipred.c:95-160,586-676), residual
// (xevd_sub_block_itdq) and xevd_recon.
//
// Intra CUs read reconstructed samples of their left / upper neighbours, so they cannot run in the fully parallel inter
// kernel.  The reference serialises them with a CTU-row wavefront (sync_flag, src_base/xevd.c:1497-1501); the same
// dependency structure is used here: one CTA per CTU, CTUs are handed out in raster order by an atomic ticket (so every
// CTA a waiter depends on has already started), a CTA spins on the `done` flags of its left, upper-left, upper and
// upper-right CTU, then reconstructs its intra CUs one after the other in decoding order, all threads cooperating on each
// CU.  Inter CUs of the picture have been reconstructed by the inter kernel before this one starts.
// Neighbour availability is order-derived in the reference (COD bits); it arrives precomputed as the per-SCU masks of
// XB200_CU_EXT.
#pragma once
#include "xb_common.cuh"
#include "xb_itdq.cuh"
#include "xb_recon.cuh"

namespace xb {

constexpr int kIntraThreads = 256;

struct IntraSmem {
    // residual of the current CU, CU-raster: luma up to 128x128, chroma 2 x 64x64
    static constexpr int kResElems = 128 * 128 + 2 * 64 * 64;
    static constexpr int kTmpElems = 64 * 65;          // pass-1 buffer of one transform block
    static constexpr int kNbElems = 2 * (2 * 128 + 8); // up[-1..w+h), left[-1..w+h)
    static size_t bytes() { return sizeof(int16_t) * kResElems + sizeof(int) * kTmpElems + sizeof(int16_t) * 3 * kNbElems + 64; }
};

// residual of one plane of one CU (all threads of the CTA): blocks larger than 64 (chroma: 32) are cut into sub-blocks gated
// by the nnz_sub bits; the result lands CU-raster in `res` (stride = plane width)
template <bool IQT>
__device__ void cu_plane_residual(const int16_t *__restrict__ coef, int lw, int lh, int lmax, int bits, int qp, int bd,
                                  int16_t *res, int *tmp, int tid, int nthreads)
{
    const int pw = 1 << lw, ph = 1 << lh;
    for (int i = tid; i < pw * ph; i += nthreads) res[i] = 0;
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
                itx_line_dyn<IQT>(slh, [&](int k) { return dq.apply(src[k * pw + x]); }, [&](int n, int v) { dst[n * ts] = v; }, sh1);
            }
            __syncthreads();
            for (int y = tid; y < h; y += nthreads) {
                const int *srow = tmp + y * ts;
                int16_t *drow = res + ((j << slh) + y) * pw + (i << slw);
                itx_line_dyn<false>(slw, [&](int k) { return srow[k]; }, [&](int n, int v) { drow[n] = (int16_t)v; }, sh2);
            }
        }
    __syncthreads();
}

// xevd_get_nbr_b for one plane: up[-1 .. w+h), left[-1 .. h+w); unit = samples per SCU (4 luma, 2 chroma)
__device__ void intra_gather(const pel *rec, int s, int w, int h, int unit, unsigned long long up_mask, unsigned long long left_mask,
                             bool up_left, int dflt, int16_t *up, int16_t *left, int tid, int nthreads)
{
    const int n = w + h;
    const int ush = unit == 4 ? 2 : 1;
    for (int i = tid; i < 2 * n + 1; i += nthreads) {
        if (i == 2 * n) {
            // L2 loads: the line may sit stale in this SM's L1 from before a neighbouring CTA wrote it
            const int v = up_left ? __ldcg(rec - s - 1) : dflt;
            up[-1] = (int16_t)v; left[-1] = (int16_t)v;
        } else if (i < n) {
            up[i] = (int16_t)(((up_mask >> (i >> ush)) & 1) ? __ldcg(rec - s + i) : dflt);
        } else {
            const int k = i - n;
            left[k] = (int16_t)(((left_mask >> (k >> ush)) & 1) ? __ldcg(rec + (ptrdiff_t)k * s - 1) : dflt);
        }
    }
}

// xevd_ipred_b + xevd_recon for one plane; modes IPD_DC_B 0, HOR 1, VER 2, UL 3, UR 4
__device__ void intra_pred_recon(pel *rec, int s, int w, int h, int lw, int mode, const int16_t *up, const int16_t *left,
                                 const int16_t *res, bool coded, int maxv, int *scratch, int tid, int nthreads)
{
    if (mode == 0) {
        // DC = (sum(left[0..h)) + sum(up[0..w)) + w) >> (log2 w + 1): warp 0 reduces
        if (tid < 32) {
            int acc = 0;
            for (int i = tid; i < h; i += 32) acc += left[i];
            for (int i = tid; i < w; i += 32) acc += up[i];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
            if (tid == 0) *scratch = (acc + w) >> (lw + 1);
        }
        __syncthreads();
    }
    const int dc = mode == 0 ? *scratch : 0;
    for (int i = tid; i < w * h; i += nthreads) {
        const int y = i >> lw, x = i & (w - 1);
        int p;
        switch (mode) {
        case 0: p = dc; break;
        case 1: p = left[y]; break;
        case 2: p = up[x]; break;
        case 3: p = y > x ? left[y - x - 1] : (y == x ? up[-1] : up[x - y - 1]); break;
        default: p = (up[x + y + 1] + left[x + y + 1]) >> 1; break;
        }
        const int r = coded ? res[i] : 0;
        rec[(size_t)y * s + x] = (pel)xb_clip3(0, maxv, (int16_t)(p + r));      // xevd_recon: s16 wrap, then clip
    }
}

struct IntraSync {
    int *ticket;        // next CTU to hand out
    int *done;          // [n_ctu] 1 when every CU of the CTU is final
};

template <bool IQT>
__global__ void __launch_bounds__(kIntraThreads)
k_recon_intra(const __grid_constant__ XbFrameArgs a, const IntraSync sy)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int16_t *s_res = (int16_t *)smem_raw;
    int *s_tmp = (int *)(s_res + IntraSmem::kResElems);
    int16_t *s_nb = (int16_t *)(s_tmp + IntraSmem::kTmpElems);
    __shared__ int s_ctu, s_scratch;
    const int tid = threadIdx.x;

    if (tid == 0) s_ctu = atomicAdd(sy.ticket, 1);
    __syncthreads();
    const int ctu = s_ctu;
    if (ctu >= a.n_ctu) return;
    const int cx = ctu % a.w_ctu, cy = ctu / a.w_ctu;
    const int cu0 = a.ctu_first[ctu], cu1 = a.ctu_first[ctu + 1];

    // any intra CU here?  (uniform: every thread scans the same descriptors through L1)
    bool any = false;
    for (int i = cu0 + tid; i < cu1; i += kIntraThreads) any |= a.cus[i].mode == XB200_MODE_INTRA;
    any = __syncthreads_or(any);
    if (any) {
        if (tid < 4) {
            // left, upper-left, upper, upper-right
            const int nx = cx + (tid == 0 ? -1 : tid - 2), ny = cy - (tid == 0 ? 0 : 1);
            if (nx >= 0 && nx < a.w_ctu && ny >= 0) {
                volatile int *f = sy.done + ny * a.w_ctu + nx;
                while (*f == 0) __nanosleep(64);
            }
            __threadfence();
        }
        __syncthreads();
        for (int i = cu0; i < cu1; i++) {
            const XB200_CU cu = a.cus[i];
            if (cu.mode != XB200_MODE_INTRA) continue;              // uniform
            const int w = 1 << cu.log2w, h = 1 << cu.log2h, cw = w >> 1, ch = h >> 1;
            uint32_t ei;
            memcpy(&ei, cu.mv[1], 4);
            const XB200_CU_EXT ex = a.ext[ei];
            const bool ul = (cu.avail >> 2) & 1;
            const int dflt = 1 << (a.bd_l - 1), maxv = (1 << a.bd_l) - 1;
            // residual of the three planes (CU-raster in shared memory)
            int16_t *res[3] = {s_res, s_res + w * h, s_res + w * h + cw * ch};
            const int16_t *coef = a.coef + cu.coef_off;
            for (int pl = 0; pl < 3; pl++) {
                const int bits = (cu.cbf >> (4 * pl)) & 15;
                if (!bits) continue;
                const int lw = cu.log2w - (pl ? 1 : 0), lh = cu.log2h - (pl ? 1 : 0);
                cu_plane_residual<IQT>(coef, lw, lh, pl ? 5 : 6, bits, pl == 0 ? cu.qp_y : (pl == 1 ? cu.qp_u : cu.qp_v), a.bd_l, res[pl], s_tmp,
                                       tid, kIntraThreads);
                coef += ((1 << (lw + lh)) + 7) & ~7;
            }
            // neighbours of all three planes, then prediction + reconstruction
            int16_t *up[3], *le[3];
            for (int pl = 0; pl < 3; pl++) { up[pl] = s_nb + pl * IntraSmem::kNbElems + 4; le[pl] = up[pl] + (2 * 128 + 8); }
            intra_gather(a.cur.y + (size_t)cu.y * a.s_l + cu.x, a.s_l, w, h, 4, ex.u.intra.up, ex.u.intra.left, ul, dflt, up[0], le[0], tid, kIntraThreads);
            intra_gather(a.cur.u + (size_t)(cu.y >> 1) * a.s_c + (cu.x >> 1), a.s_c, cw, ch, 2, ex.u.intra.up, ex.u.intra.left, ul, dflt, up[1], le[1],
                         tid, kIntraThreads);
            intra_gather(a.cur.v + (size_t)(cu.y >> 1) * a.s_c + (cu.x >> 1), a.s_c, cw, ch, 2, ex.u.intra.up, ex.u.intra.left, ul, dflt, up[2], le[2],
                         tid, kIntraThreads);
            __syncthreads();
            intra_pred_recon(a.cur.y + (size_t)cu.y * a.s_l + cu.x, a.s_l, w, h, cu.log2w, cu.refi[0], up[0], le[0], res[0], (cu.cbf & 0x00f) != 0, maxv,
                             &s_scratch, tid, kIntraThreads);
            __syncthreads();
            intra_pred_recon(a.cur.u + (size_t)(cu.y >> 1) * a.s_c + (cu.x >> 1), a.s_c, cw, ch, cu.log2w - 1, cu.refi[1], up[1], le[1], res[1],
                             (cu.cbf & 0x0f0) != 0, maxv, &s_scratch, tid, kIntraThreads);
            __syncthreads();
            intra_pred_recon(a.cur.v + (size_t)(cu.y >> 1) * a.s_c + (cu.x >> 1), a.s_c, cw, ch, cu.log2w - 1, cu.refi[1], up[2], le[2], res[2],
                             (cu.cbf & 0xf00) != 0, maxv, &s_scratch, tid, kIntraThreads);
            __syncthreads();         // the next CU may read these samples (global writes are visible block-wide after the barrier)
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(sy.done + ctu, 1);
}

}  // namespace xb
