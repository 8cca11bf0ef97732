// xb_common.cuh -- shared device-side declarations of the B200 reconstruction path.
//
// Layout in HBM: pictures are padded 16-bit planes exactly like the reference's
// XEVD_PIC (pad 144 luma / 72 chroma, xevd_util.c:153-230); per-SCU maps follow XEVD_PIC.map_mv /
// map_refi and ctx->map_scu.  CU work items, the coefficient stream and the per-CTU index are the
// flat arrays of include/xevd_b200.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <utility>
#include "../../include/xevd_b200.h"

#define XB_MAX_REFS 17          // XEVD_MAX_NUM_REF_PICS (21) is never reached per list; 17 = 16 + 1

typedef int16_t pel;

struct XbPlanes {
    pel *y, *u, *v;             // sample (0,0)
};

// everything one picture-level kernel needs, passed by value (__grid_constant__)
struct XbFrameArgs {
    XbPlanes cur;
    const pel *ref_y[2][XB_MAX_REFS];
    const pel *ref_u[2][XB_MAX_REFS];
    const pel *ref_v[2][XB_MAX_REFS];
    int ref_poc[2][XB_MAX_REFS];
    const CUtensorMap *ref_tmap[2 * XB_MAX_REFS];   // per reference picture: 3 TMA descriptors (Y, U, V) in device memory
    int s_l, s_c;               // strides in pels (same geometry for every picture of a sequence)
    int w, h;                   // luma size
    int bd_l, bd_c;
    int log2_ctu, w_ctu, n_ctu;
    int peer_maps;              // debug: 0 = keep the per-SCU maps local (XB200_PEER_NOMAPS=1)
    int n_peer;                 // band mode over NVLink: every output store is repeated into the same picture on n_peer other GPUs
    long long peer_delta[7];    //   byte distance from this GPU's picture allocation to the peer's mapping of its twin (same layout)
    int ctu_row0;               // band mode: first CTU row of this launch (the CU arrays / ctu_first index CTUs relative to it)
    int main_tables, iqt, eipd, ats, htdf, slice_qp, dmvr, poc, affine, ibc;
    int constrained;              // pps.constrained_intra_pred_flag (HTDF ring of intra CUs)
    int dispatch;                 // 1: the throughput kernel reconstructs the CUs it has code for, the generic kernel the ATS / DMVR / affine CUs
    int *inter_count;             // (with inter_done) number of CTUs the generic kernel has finished: once it equals n_ctu no flag needs a look
    int *inter_done;              // null, or one flag per CTU of this launch: the generic inter kernel sets it when the CTU's inter CUs are final, the
                                  // wavefront kernel - running beside it on a second stream - waits for the flags of the CTUs it is about to read
    const XB200_CU *cus;
    const uint32_t *ctu_first;
    const int16_t *coef;
    const XB200_CU_EXT *ext;
    int16_t *map_mv;            // [scu][2][2]
    int16_t *map_unrefined_mv;  // [scu][2][2] vectors before DMVR refinement (equal to map_mv elsewhere)
    int8_t *map_refi;           // [scu][2]
    uint32_t *map_scu;
    uint8_t *map_edge;          // XB200_EDGE_* per SCU: CU boundaries + the 64-sample transform split of larger CUs
    uint16_t *map_order;        // index (inside its CTU, decoding order) of the CU that owns the SCU's CHROMA samples; null unless tool_suco
    int w_scu, h_scu;
};

// CUs that read samples of the CURRENT picture (intra prediction, intra block copy) are reconstructed by the CTU wavefront
// kernel (xb_intra.cuh); everything else by the fully parallel inter kernels
__host__ __device__ __forceinline__ bool xb_wavefront_mode(int mode) { return mode == XB200_MODE_INTRA || mode == XB200_MODE_IBC; }

// store to the local picture and to its twins on the peer GPUs (P2P stores over NVLink; no-op loop when n_peer == 0)
template <typename T>
__device__ __forceinline__ void xb_store_all(const XbFrameArgs &a, T *dst, T v, bool fan_out = true)
{
    *dst = v;
    if (fan_out)
        for (int k = 0; k < a.n_peer; k++) *(T *)((char *)dst + a.peer_delta[k]) = v;
}

// Programmatic dependent launch (stream serialisation relaxed by the launch attribute below): the CTAs of a kernel launched through
// xb_launch_early become resident while the kernel before it in the stream drains - once every CTA of that kernel has started - and do their
// index arithmetic and shared-memory set-up; xb_grid_wait() returns when the kernel before has completed and its stores are visible, so it
// stands before the first global access.  Both instructions are no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void xb_grid_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void xb_grid_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t xb_launch_early(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

__device__ __forceinline__ int xb_clip3(int lo, int hi, int v) { return max(lo, min(hi, v)); }
__device__ __forceinline__ int xb_clip16(int v) { return max(-32768, min(32767, v)); }

// The library is a single translation unit (libxevd_b200.cu includes every .cuh), so __constant__
// tables are plain definitions.
// interpolation taps [main_tables][phase][tap]: src_base/xevd_mc.c:80-134, src_main/xevdm_mc.c:121-175
__constant__ int16_t c_mc_l[2][16][8] = {
    {{0, 0, 0, 64, 0, 0, 0, 0}, {0}, {0}, {0}, {0, 1, -5, 52, 20, -5, 1, 0}, {0}, {0}, {0},
     {0, 2, -10, 40, 40, -10, 2, 0}, {0}, {0}, {0}, {0, 1, -5, 20, 52, -5, 1, 0}, {0}, {0}, {0}},
    {{0, 0, 0, 64, 0, 0, 0, 0},        {0, 1, -3, 63, 4, -2, 1, 0},      {-1, 2, -5, 62, 8, -3, 1, 0},
     {-1, 3, -8, 60, 13, -4, 1, 0},    {-1, 4, -10, 58, 17, -5, 1, 0},   {-1, 4, -11, 52, 26, -8, 3, -1},
     {-1, 3, -9, 47, 31, -10, 4, -1},  {-1, 4, -11, 45, 34, -10, 4, -1}, {-1, 4, -11, 40, 40, -11, 4, -1},
     {-1, 4, -10, 34, 45, -11, 4, -1}, {-1, 4, -10, 31, 47, -9, 3, -1},  {-1, 3, -8, 26, 52, -11, 4, -1},
     {0, 1, -5, 17, 58, -10, 4, -1},   {0, 1, -4, 13, 60, -8, 3, -1},    {0, 1, -3, 8, 62, -5, 2, -1},
     {0, 1, -2, 4, 63, -3, 1, 0}}};
__constant__ int16_t c_mc_c[2][32][4] = {
    {{0, 64, 0, 0},   {0}, {0}, {0}, {-2, 58, 10, -2}, {0}, {0}, {0}, {-4, 52, 20, -4}, {0}, {0}, {0},
     {-6, 46, 30, -6}, {0}, {0}, {0}, {-8, 40, 40, -8}, {0}, {0}, {0}, {-6, 30, 46, -6}, {0}, {0}, {0},
     {-4, 20, 52, -4}, {0}, {0}, {0}, {-2, 10, 58, -2}, {0}, {0}, {0}},
    {{0, 64, 0, 0},    {-1, 63, 2, 0},   {-2, 62, 4, 0},   {-2, 60, 7, -1},  {-2, 58, 10, -2}, {-3, 57, 12, -2},
     {-4, 56, 14, -2}, {-4, 55, 15, -2}, {-4, 54, 16, -2}, {-5, 53, 18, -2}, {-6, 52, 20, -2}, {-6, 49, 24, -3},
     {-6, 46, 28, -4}, {-5, 44, 29, -4}, {-4, 42, 30, -4}, {-4, 39, 33, -4}, {-4, 36, 36, -4}, {-4, 33, 39, -4},
     {-4, 30, 42, -4}, {-4, 29, 44, -5}, {-4, 28, 46, -6}, {-3, 24, 49, -6}, {-2, 20, 52, -6}, {-2, 18, 53, -5},
     {-2, 16, 54, -4}, {-2, 15, 55, -4}, {-2, 14, 56, -4}, {-2, 12, 57, -3}, {-2, 10, 58, -2}, {-1, 7, 60, -2},
     {0, 4, 62, -2},   {0, 2, 63, -1}}};
// xevd_tbl_dq_scale_b / xevd_tbl_dq_scale (src_base/xevd_tbl.c:255-256), [iqt][qp % 6]
__constant__ int c_dq_scale[2][6] = {{40, 45, 51, 57, 64, 71}, {40, 45, 51, 57, 64, 72}};
