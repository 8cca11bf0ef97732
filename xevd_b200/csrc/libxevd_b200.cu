// libxevd_b200.cu -- the C ABI (include/xevd_b200.h) over the sm_100a kernels.  Single translation unit.
//
// Host side of the boundary: device pictures (PICBUF_ALLOCATOR replacement), argument marshalling,
// pinned staging for the host-pointer entry points, launches on one CUDA stream per context.
// No CPU implementation of any kernel exists here: without a device every call fails.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <new>
#include <atomic>

#include "xb_common.cuh"
#include "xb_itdq.cuh"
#include "xb_recon.cuh"
#include "xb_recon2.cuh"
#include "xb_intra.cuh"
#include "xb_filters.cuh"
#include "xb_micro.cuh"
#include "xb_alf.cuh"

struct xb200_pic {
    int w, h, w_c, h_c, s_l, s_c, pad_l, pad_c, w_scu, h_scu, poc;
    size_t luma_elems, chroma_elems;
    pel *buf;                    // one allocation: Y | U | V padded planes
    pel *y, *u, *v;              // sample (0,0)
    CUtensorMap *d_tmaps;        // device copy of 3 TMA descriptors (whole padded Y, U, V planes)
    int16_t *map_mv;
    int16_t *map_unrefined_mv;   // Main: vectors before DMVR refinement (mctx->map_unrefined_mv); equal to map_mv elsewhere
    int8_t *map_refi;
    uint32_t *map_scu;
    uint8_t *map_edge;
    uint16_t *map_order;
    size_t alloc_bytes;
    int n_peer;                  // twins of this picture on other GPUs, opened over CUDA IPC (band mode with P2P stores)
    void *peer_base[7];
};

struct Staging {                 // one slot of the host->device staging ring
    void *pinned = nullptr, *dev = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;  // recorded after the kernel that consumes the slot
    bool busy = false;
};

// Only ONE context of the process runs its wavefront kernel beside the generic kernel (recon_frame_dev): wavefront CTAs spin on flags the
// generic kernel raises, and the waiters of several contexts together could occupy every SM before the kernels they wait for get a CTA in.
// The first context that needs it owns the right until it is destroyed; the others launch the two kernels one after the other.
static std::atomic<xb200_ctx *> g_overlap_owner{nullptr};

struct xb200_ctx {
    int device;
    cudaStream_t stream;
    bool own_stream;
    bool wavefront_alone;        // XB200_WAVEFRONT_ALONE=1: one wavefront CTA per SM
    long long launches;
    char err[256];
    Staging ring[3];
    int ring_pos;
    int sm_count;
    int *d_sync;                 // wavefront state of the intra kernel: [0] ticket, [1..] per-CTU done flags
    int *d_err;                  // sticky error word of the wavefront kernel (a wait that gave up), read by xb200_sync
    bool wave_used;              // a wavefront kernel has been launched since the last xb200_sync
    int sync_cap;
    int *d_order;                // CTU addresses in wavefront order (x + 2y) for the current picture geometry
    int order_w, order_n;
    int8_t chroma_qp[2][58];     // xevd_qp_chroma_dynamic for the sequence
    pel *alf_copy;               // pre-ALF copy of the picture being filtered
    size_t alf_cap;
    uint8_t *alf_flags_pinned, *alf_flags_dev;
    int tile_cols, tile_rows, tile_across;   // PPS tile grid (xb200_set_tiles); 1 x 1 = one tile
    uint16_t tile_col_bd[XB200_MAX_TILE_COLS + 1], tile_row_bd[XB200_MAX_TILE_ROWS + 1];
    void *alf_tab_dev, *alf_tab_pinned;      // ALF: the 100 permuted luma filters of the current APS
    cudaEvent_t alf_tab_done;
    bool alf_tab_valid;
    int16_t alf_tab_coef[25][13];            // the coefficients the device table was built from
    int alf_flags_cap;
    cudaEvent_t alf_flags_done;
    unsigned char *out_buf;      // output path: packed planes produced by k_output, then copied to the caller
    size_t out_cap;
    int *d_dra;                  // DRA LUTs on the device (3 x 1024 ints)
    bool peer_maps;              // XB200_PEER_NOMAPS=1 (read once at creation) keeps the per-SCU maps local in band mode (debug)
    bool force_generic;          // XB200_FORCE_GENERIC=1: route everything through the generic kernel (debug / A-B tests)
    bool no_overlap;             // XB200_NO_OVERLAP=1: the wavefront kernel starts after the generic inter kernel instead of beside it (A/B, debugging)
    cudaStream_t stream_w;       // second stream: the wavefront kernel when it runs beside the generic inter kernel
    cudaEvent_t ev_fork, ev_join;
    int v2_variant;              // XB200_V2_VARIANT=rounds | slots: pin the prediction-stage variant of the throughput kernel (tests run both); else per picture
};

#define CK(ctx, call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), "%s: %s", #call, cudaGetErrorString(e_)); \
            return XB200_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)

extern "C" {

int xb200_abi_version(void) { return XB200_ABI_VERSION; }

int xb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return XB200_ERR_NO_DEVICE;
    return n;
}

xb200_ctx *xb200_create(int device, int *err)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        if (err) *err = XB200_ERR_NO_DEVICE;
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { if (err) *err = XB200_ERR_NO_DEVICE; return nullptr; }
    xb200_ctx *c = new (std::nothrow) xb200_ctx();
    if (!c) { if (err) *err = XB200_ERR_OUT_OF_MEMORY; return nullptr; }
    c->device = device;
    c->own_stream = true;
    c->launches = 0;
    c->err[0] = 0;
    c->ring_pos = 0;
    c->stream_w = nullptr; c->ev_fork = c->ev_join = nullptr;
    c->d_sync = nullptr; c->sync_cap = 0; c->d_err = nullptr; c->wave_used = false;
    c->d_order = nullptr; c->order_w = c->order_n = 0;
    c->out_buf = nullptr; c->out_cap = 0; c->d_dra = nullptr;
    c->tile_cols = c->tile_rows = 1; c->tile_across = 0;
    c->tile_col_bd[0] = c->tile_row_bd[0] = 0; c->tile_col_bd[1] = c->tile_row_bd[1] = 0xffff;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        if (err) *err = XB200_ERR_CUDA;
        return nullptr;
    }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    // opt in to large dynamic shared memory for the CTU kernels (CTU 128 needs ~150 KB); a kernel that cannot get its shared memory would
    // fail at its first launch with a less readable error, so creation fails instead
    cudaError_t fe = cudaSuccess;
#define XB_SMEM(fn, bytes) do { if (fe == cudaSuccess) fe = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
                                if (fe == cudaSuccess) fe = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); } while (0)
    XB_SMEM((xb::k_recon_inter<false>), (int)xb::ReconSmem::bytes(7));
    XB_SMEM((xb::k_recon_inter<true>), (int)xb::ReconSmem::bytes(7));
    XB_SMEM((xb::k_recon_intra<false>), (int)xb::IntraSmem::bytes(7));
    XB_SMEM((xb::k_recon_intra<true>), (int)xb::IntraSmem::bytes(7));
    XB_SMEM((xb::k_itdq_blocks<false>), xb::itdq_blocks_smem(1, 6, 32));          // the widest case: 2-wide blocks (row stride w + 4)
    XB_SMEM((xb::k_itdq_blocks<true>), xb::itdq_blocks_smem(1, 6, 32));
    XB_SMEM((xb::k_recon_inter_v2<false>), xb::R2Layout::make(1, 256).total);
    XB_SMEM((xb::k_recon_inter_v2<true>), xb::R2Layout::make(2, 256).total);
    XB_SMEM((xb::k_recon_inter_v2<false, true>), xb::R2Layout::make(1, 256, true).total);
    XB_SMEM((xb::k_recon_inter_v2<true, true>), xb::R2Layout::make(2, 256, true).total);
    XB_SMEM((xb::k_recon_inter_v2<false, false, true>), xb::R2Layout::make(1, 256).total);
    XB_SMEM((xb::k_recon_inter_v2<true, false, true>), xb::R2Layout::make(2, 256).total);
    XB_SMEM((xb::k_recon_inter_v2<false, false, true, true>), xb::R2Layout::make(1, 256).total);
    XB_SMEM((xb::k_recon_inter_v2<true, false, true, true>), xb::R2Layout::make(2, 256).total);
    XB_SMEM((xb::k_recon_inter_v2<false, false, false, true>), xb::R2Layout::make(1, 256).total);
    XB_SMEM((xb::k_recon_inter_v2<true, false, false, true>), xb::R2Layout::make(2, 256).total);
    XB_SMEM((xb::k_recon_inter_v2<false, false, false, false, true>), xb::R2Layout::make(1, 256, false, true).total);
    XB_SMEM((xb::k_recon_inter_v2<true, false, false, false, true>), xb::R2Layout::make(2, 256, false, true).total);
    XB_SMEM((xb::k_recon_inter_v2<false, false, true, false, true>), xb::R2Layout::make(1, 256, false, true).total);
    XB_SMEM((xb::k_recon_inter_v2<true, false, true, false, true>), xb::R2Layout::make(2, 256, false, true).total);
    XB_SMEM((xb::k_recon_inter_v2<false, false, true, true, true>), xb::R2Layout::make(1, 256, false, true).total);
    XB_SMEM((xb::k_recon_inter_v2<true, false, true, true, true>), xb::R2Layout::make(2, 256, false, true).total);
    XB_SMEM((xb::k_recon_inter_v2<false, false, false, true, true>), xb::R2Layout::make(1, 256, false, true).total);
    XB_SMEM((xb::k_recon_inter_v2<true, false, false, true, true>), xb::R2Layout::make(2, 256, false, true).total);
#undef XB_SMEM
    if (fe != cudaSuccess) {
        fprintf(stderr, "[xb200] cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed: %s\n", cudaGetErrorString(fe));
        cudaStreamDestroy(c->stream);
        delete c;
        if (err) *err = XB200_ERR_CUDA;
        return nullptr;
    }
    {   // xevd_tbl_qp_chroma_adjust_base (src_base/xevd_tbl.c:345-355): the default when the SPS carries no table
        static const int8_t base[58] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
                                        29, 29, 30, 31, 32, 32, 33, 33, 34, 34, 35, 35, 36, 36, 36, 37, 37, 37, 38, 38, 39, 39, 40, 40, 40, 41, 41, 41};
        memcpy(c->chroma_qp[0], base, 58); memcpy(c->chroma_qp[1], base, 58);
    }
    { const char *e = getenv("XB200_FORCE_GENERIC"); c->force_generic = e && e[0] == '1'; }
    { const char *e = getenv("XB200_NO_OVERLAP"); c->no_overlap = e && e[0] == '1'; }
    { const char *e = getenv("XB200_V2_VARIANT"); c->v2_variant = !e ? 0 : (!strcmp(e, "rounds") ? 1 : (!strcmp(e, "slots") ? 2 : 0)); }
    { const char *e = getenv("XB200_PEER_NOMAPS"); c->peer_maps = !(e && e[0] == '1'); }
    { const char *e = getenv("XB200_WAVEFRONT_ALONE"); c->wavefront_alone = e && e[0] == '1'; }
    {   // packed IDP.2A tap tables for the throughput kernel, derived from the interpolation tables
        int16_t hl[2][16][8], hc[2][32][4];
        cudaMemcpyFromSymbol(hl, c_mc_l, sizeof(hl));
        cudaMemcpyFromSymbol(hc, c_mc_c, sizeof(hc));
        int t5[2][16 * 9], t3[2][32 * 6];
        for (int m = 0; m < 2; m++) {
            for (int ph = 0; ph < 16; ph++) xb::build_taps8(hl[m][ph], t5[m] + ph * 9);
            for (int ph = 0; ph < 32; ph++) xb::build_taps4(hc[m][ph], t3[m] + ph * 6);
        }
        cudaMemcpyToSymbol(xb::c_taps5, t5, sizeof(t5));
        cudaMemcpyToSymbol(xb::c_taps3, t3, sizeof(t3));
    }
    {   // inverse DST-7 / DCT-8 kernels for tool_ats: the reference generates them at start-up with double-precision sin / cos
        // (xevdm_init_multi_tbl / xevd_init_multi_inv_tbl, src_main/xevdm_itdq.c:81-159); same expression here
        static int16_t m[2][1360];
        const double PI = 3.14159265358979323846;
        int off = 0;
        for (int lg = 2; lg <= 5; lg++) {
            const int n = 1 << lg;
            const double sc = sqrt((double)n) * 64.0;
            for (int k = 0; k < n; k++)
                for (int j = 0; j < n; j++) {
                    double v = cos(PI * (k + 0.5) * (j + 0.5) / (n + 0.5)) * sqrt(2.0 / (n + 0.5));
                    m[0][off + j * n + k] = (int16_t)(sc * v + (v > 0 ? 0.5 : -0.5));
                    v = sin(PI * (k + 0.5) * (j + 1) / (n + 0.5)) * sqrt(2.0 / (n + 0.5));
                    m[1][off + j * n + k] = (int16_t)(sc * v + (v > 0 ? 0.5 : -0.5));
                }
            off += n * n;
        }
        cudaMemcpyToSymbol(xb::g_ats_inv, m, sizeof(m));
    }
    if (getenv("XB200_DEBUG")) {
        for (int nl = 1; nl <= 2; nl++)
            for (int mc = 16; mc <= 256; mc *= 4) {
                int nb = 0;
                const int sm = xb::R2Layout::make(nl, mc).total, smw = xb::R2Layout::make(nl, mc, false, true).total;
                int nbw = 0;
                if (nl == 1) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, xb::k_recon_inter_v2<false>, xb::kR2Threads, sm);
                               cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbw, xb::k_recon_inter_v2<false, false, false, false, true>, xb::kR2Threads, smw); }
                else { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, xb::k_recon_inter_v2<true>, xb::kR2Threads, sm);
                       cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbw, xb::k_recon_inter_v2<true, false, false, false, true>, xb::kR2Threads, smw); }
                fprintf(stderr, "[xb200] k_recon_inter_v2 lists=%d max_cu=%d: rounds %d B dynamic smem, %d CTAs/SM; warp slots %d B, %d CTAs/SM\n", nl, mc, sm, nb, smw, nbw);
            }
    }
    if (err) *err = XB200_OK;
    return c;
}

void xb200_destroy(xb200_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto &s : c->ring) {
        if (s.pinned) cudaFreeHost(s.pinned);
        if (s.dev) cudaFree(s.dev);
        if (s.done) cudaEventDestroy(s.done);
    }
    if (c->d_sync) cudaFree(c->d_sync);
    if (c->d_err) cudaFree(c->d_err);
    if (c->d_order) cudaFree(c->d_order);
    if (c->out_buf) cudaFree(c->out_buf);
    if (c->d_dra) cudaFree(c->d_dra);
    if (c->alf_copy) cudaFree(c->alf_copy);
    if (c->alf_flags_pinned) cudaFreeHost(c->alf_flags_pinned);
    if (c->alf_flags_dev) cudaFree(c->alf_flags_dev);
    if (c->alf_tab_dev) cudaFree(c->alf_tab_dev);
    if (c->alf_tab_pinned) cudaFreeHost(c->alf_tab_pinned);
    if (c->alf_tab_done) cudaEventDestroy(c->alf_tab_done);
    if (c->alf_flags_done) cudaEventDestroy(c->alf_flags_done);
    { xb200_ctx *me = c; g_overlap_owner.compare_exchange_strong(me, nullptr); }
    if (c->stream_w) cudaStreamDestroy(c->stream_w);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char *xb200_last_error(xb200_ctx *c) { return c ? c->err : "no context"; }
int xb200_sync(xb200_ctx *c)
{
    if (!c) return XB200_ERR_INVALID_ARGUMENT;
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaGetLastError());
    if (c->wave_used && c->d_err) {       // did a CTU of the wavefront kernel give up waiting for a neighbour (malformed work list)?
        int e = 0;
        c->wave_used = false;
        CK(c, cudaMemcpy(&e, c->d_err, sizeof(int), cudaMemcpyDeviceToHost));
        if (e) {
            cudaMemset(c->d_err, 0, sizeof(int));
            snprintf(c->err, sizeof(c->err), "wavefront kernel: a CTU waited for a neighbour that never completed (inconsistent ctu_first / CU list)");
            return XB200_ERR_INVALID_ARGUMENT;
        }
    }
    return XB200_OK;
}
void *xb200_stream(xb200_ctx *c) { return c ? (void *)c->stream : nullptr; }
int xb200_set_stream(xb200_ctx *c, void *s)
{
    if (!c) return XB200_ERR_INVALID_ARGUMENT;
    if (c->own_stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); c->own_stream = false; }
    c->stream = (cudaStream_t)s;
    return XB200_OK;
}
long long xb200_launch_count(xb200_ctx *c) { return c ? c->launches : 0; }

// ---- page-locked host memory for the producer side (coefficient stream, CU arrays, output planes) ---------------------
void *xb200_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void xb200_host_free(void *p) { if (p) cudaFreeHost(p); }
int xb200_host_register(void *p, size_t bytes)
{
    if (!p || !bytes) return XB200_ERR_INVALID_ARGUMENT;
    if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return XB200_ERR_CUDA; }
    return XB200_OK;
}
int xb200_host_unregister(void *p)
{
    if (!p) return XB200_ERR_INVALID_ARGUMENT;
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return XB200_ERR_CUDA; }
    return XB200_OK;
}

// ---- pictures ---------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// TMA descriptors over the whole padded planes: box 40x23 (luma) / 24x11 (chroma) = the 8-sample-aligned superset of
// the interpolation window of a 16x16 tile (8-tap: +7, 4-tap: +3)
static int make_tensor_maps(xb200_ctx *c, xb200_pic *p)
{
    static PFN_encodeTiled encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) {
            snprintf(c->err, sizeof(c->err), "cuTensorMapEncodeTiled unavailable");
            return XB200_ERR_CUDA;
        }
        encode = (PFN_encodeTiled)fn;
    }
    // maps[0]: luma plane (2-D).  maps[1]: Cb and Cr as one 3-D tensor (x, y, plane) - the planes have the same geometry and lie
    // chroma_elems apart, so one TMA instruction fetches a tile's two chroma windows.  maps[2], maps[3]: Cr / Cb alone (2-D; xb200_mc_blocks_dev).
    alignas(64) CUtensorMap maps[4];
    for (int pl = 0; pl < 4; pl++) {
        const bool luma = pl == 0, both = pl == 1;
        void *base = luma ? (void *)p->buf : (void *)(p->buf + p->luma_elems + (pl == 2 ? p->chroma_elems : 0));
        const cuuint64_t dims[3] = {(cuuint64_t)(luma ? p->s_l : p->s_c), (cuuint64_t)(luma ? p->h + 2 * p->pad_l : p->h_c + 2 * p->pad_c), 2};
        const cuuint64_t strides[2] = {(cuuint64_t)(luma ? p->s_l : p->s_c) * 2, (cuuint64_t)p->chroma_elems * 2};
        const cuuint32_t box[3] = {luma ? (cuuint32_t)xb::kBoxLW : (cuuint32_t)xb::kBoxCW, luma ? (cuuint32_t)xb::kBoxLH : (cuuint32_t)xb::kBoxCH, 2};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&maps[pl], CU_TENSOR_MAP_DATA_TYPE_UINT16, both ? 3 : 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            snprintf(c->err, sizeof(c->err), "cuTensorMapEncodeTiled failed (%d) plane %d", (int)r, pl);
            return XB200_ERR_CUDA;
        }
    }
    CK(c, cudaMemcpyAsync(p->d_tmaps, maps, sizeof(maps), cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return XB200_OK;
}

xb200_pic *xb200_pic_alloc(xb200_ctx *c, int w, int h, int *err)
{
    if (!c || w <= 0 || h <= 0 || (w & 7) || (h & 7)) { if (err) *err = XB200_ERR_INVALID_ARGUMENT; return nullptr; }
    cudaSetDevice(c->device);
    xb200_pic *p = new (std::nothrow) xb200_pic();
    if (!p) { if (err) *err = XB200_ERR_OUT_OF_MEMORY; return nullptr; }
    p->w = w; p->h = h; p->w_c = w >> 1; p->h_c = h >> 1;
    p->pad_l = 144; p->pad_c = 72;                       // PIC_PAD_SIZE_L / _C (xevd_def.h:211-212)
    // strides: the reference's (w + 2*pad), chroma rounded up to 8 samples so every row is 16-byte aligned (TMA)
    p->s_l = w + 2 * p->pad_l; p->s_c = (p->w_c + 2 * p->pad_c + 7) & ~7;
    p->w_scu = (w + 3) >> 2; p->h_scu = (h + 3) >> 2;
    p->poc = 0;
    p->luma_elems = (size_t)p->s_l * (h + 2 * p->pad_l);
    p->chroma_elems = (size_t)p->s_c * (p->h_c + 2 * p->pad_c);
    const size_t nscu = (size_t)p->w_scu * p->h_scu;
    const size_t pix_bytes = (p->luma_elems + 2 * p->chroma_elems) * sizeof(pel);
    const size_t pix_al = (pix_bytes + 255) & ~(size_t)255;
    const size_t total = pix_al + ((nscu * 26 + 255) & ~(size_t)255) + 4 * sizeof(CUtensorMap) + 256;
    if (cudaMalloc((void **)&p->buf, total) != cudaSuccess) {
        snprintf(c->err, sizeof(c->err), "cudaMalloc(%zu) failed", total);
        delete p;
        if (err) *err = XB200_ERR_OUT_OF_MEMORY;
        return nullptr;
    }
    p->alloc_bytes = total;
    p->n_peer = 0;
    cudaMemsetAsync(p->buf, 0, total, c->stream);
    p->y = p->buf + (size_t)p->pad_l * p->s_l + p->pad_l;
    p->u = p->buf + p->luma_elems + (size_t)p->pad_c * p->s_c + p->pad_c;
    p->v = p->buf + p->luma_elems + p->chroma_elems + (size_t)p->pad_c * p->s_c + p->pad_c;
    unsigned char *m = (unsigned char *)p->buf + pix_al;
    p->map_mv = (int16_t *)m;
    p->map_scu = (uint32_t *)(m + nscu * 8);
    p->map_refi = (int8_t *)(m + nscu * 12);
    p->map_edge = (uint8_t *)(m + nscu * 14);
    p->map_unrefined_mv = (int16_t *)(m + nscu * 16);
    p->map_order = (uint16_t *)(m + nscu * 24);
    p->d_tmaps = (CUtensorMap *)(m + ((nscu * 26 + 255) & ~(size_t)255));
    if (make_tensor_maps(c, p) != XB200_OK) {
        cudaFree(p->buf);
        delete p;
        if (err) *err = XB200_ERR_CUDA;
        return nullptr;
    }
    if (err) *err = XB200_OK;
    return p;
}

void xb200_pic_free(xb200_ctx *c, xb200_pic *p)
{
    if (!p) return;
    if (c) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
    for (int k = 0; k < p->n_peer; k++) cudaIpcCloseMemHandle(p->peer_base[k]);
    cudaFree(p->buf);
    delete p;
}

int xb200_pic_info(xb200_pic *p, XB200_PIC_INFO *i)
{
    if (!p || !i) return XB200_ERR_INVALID_ARGUMENT;
    i->w_l = p->w; i->h_l = p->h; i->w_c = p->w_c; i->h_c = p->h_c;
    i->s_l = p->s_l; i->s_c = p->s_c; i->pad_l = p->pad_l; i->pad_c = p->pad_c;
    i->dev_y = p->y; i->dev_u = p->u; i->dev_v = p->v;
    i->dev_map_mv = p->map_mv; i->dev_map_refi = p->map_refi; i->dev_map_scu = p->map_scu;
    i->w_scu = p->w_scu; i->h_scu = p->h_scu; i->poc = p->poc; i->dev_map_edge = p->map_edge; i->dev_map_unrefined_mv = p->map_unrefined_mv;
    return XB200_OK;
}

int xb200_pic_set_poc(xb200_pic *p, int poc) { if (!p) return XB200_ERR_INVALID_ARGUMENT; p->poc = poc; return XB200_OK; }

int xb200_pic_upload(xb200_ctx *c, xb200_pic *p, const xb200_pel *y, int sy, const xb200_pel *u, int su, const xb200_pel *v, int sv)
{
    if (!c || !p || !y || !u || !v) return XB200_ERR_INVALID_ARGUMENT;
    CK(c, cudaMemcpy2DAsync(p->y, p->s_l * 2, y, sy * 2, p->w * 2, p->h, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemcpy2DAsync(p->u, p->s_c * 2, u, su * 2, p->w_c * 2, p->h_c, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemcpy2DAsync(p->v, p->s_c * 2, v, sv * 2, p->w_c * 2, p->h_c, cudaMemcpyHostToDevice, c->stream));
    return XB200_OK;
}

int xb200_pic_download(xb200_ctx *c, xb200_pic *p, xb200_pel *y, int sy, xb200_pel *u, int su, xb200_pel *v, int sv)
{
    if (!c || !p || !y || !u || !v) return XB200_ERR_INVALID_ARGUMENT;
    CK(c, cudaMemcpy2DAsync(y, sy * 2, p->y, p->s_l * 2, p->w * 2, p->h, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpy2DAsync(u, su * 2, p->u, p->s_c * 2, p->w_c * 2, p->h_c, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpy2DAsync(v, sv * 2, p->v, p->s_c * 2, p->w_c * 2, p->h_c, cudaMemcpyDeviceToHost, c->stream));
    return XB200_OK;
}

int xb200_pic_download_padded(xb200_ctx *c, xb200_pic *p, xb200_pel *y, xb200_pel *u, xb200_pel *v)
{
    if (!c || !p || !y || !u || !v) return XB200_ERR_INVALID_ARGUMENT;
    // host layout = the reference's (stride w + 2*pad); the device chroma stride may be wider
    const int hs_c = p->w_c + 2 * p->pad_c;
    CK(c, cudaMemcpyAsync(y, p->buf, p->luma_elems * 2, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpy2DAsync(u, hs_c * 2, p->buf + p->luma_elems, p->s_c * 2, hs_c * 2, p->h_c + 2 * p->pad_c, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpy2DAsync(v, hs_c * 2, p->buf + p->luma_elems + p->chroma_elems, p->s_c * 2, hs_c * 2, p->h_c + 2 * p->pad_c,
                            cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return XB200_OK;
}

int xb200_pic_download_maps(xb200_ctx *c, xb200_pic *p, int16_t *map_mv, int8_t *map_refi, uint32_t *map_scu)
{
    if (!c || !p) return XB200_ERR_INVALID_ARGUMENT;
    const size_t n = (size_t)p->w_scu * p->h_scu;
    if (map_mv) CK(c, cudaMemcpyAsync(map_mv, p->map_mv, n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (map_refi) CK(c, cudaMemcpyAsync(map_refi, p->map_refi, n * 2, cudaMemcpyDeviceToHost, c->stream));
    if (map_scu) CK(c, cudaMemcpyAsync(map_scu, p->map_scu, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return XB200_OK;
}

int xb200_pic_download_unrefined_mv(xb200_ctx *c, xb200_pic *p, int16_t *map_unrefined_mv)
{
    if (!c || !p || !map_unrefined_mv) return XB200_ERR_INVALID_ARGUMENT;
    CK(c, cudaMemcpyAsync(map_unrefined_mv, p->map_unrefined_mv, (size_t)p->w_scu * p->h_scu * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return XB200_OK;
}

int xb200_pic_download_edge_map(xb200_ctx *c, xb200_pic *p, uint8_t *map_edge)
{
    if (!c || !p || !map_edge) return XB200_ERR_INVALID_ARGUMENT;
    CK(c, cudaMemcpyAsync(map_edge, p->map_edge, (size_t)p->w_scu * p->h_scu, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return XB200_OK;
}

// ---- reconstruction ------------------------------------------------------------------------------------------------
static int fill_args(xb200_ctx *c, const XB200_PARAMS *prm, xb200_pic *cur, xb200_pic *const *l0, int n0, xb200_pic *const *l1, int n1,
                     XbFrameArgs &a)
{
    if (!c || !prm || !cur) return XB200_ERR_INVALID_ARGUMENT;
    if (prm->chroma_format_idc != 1) return XB200_ERR_UNSUPPORTED;
    if (prm->w != cur->w || prm->h != cur->h) return XB200_ERR_INVALID_ARGUMENT;
    if (prm->log2_ctu < 5 || prm->log2_ctu > 7) return XB200_ERR_INVALID_ARGUMENT;
    if (n0 < 0 || n1 < 0 || n0 > XB_MAX_REFS || n1 > XB_MAX_REFS) return XB200_ERR_INVALID_ARGUMENT;
    if (prm->bit_depth_luma < 8 || prm->bit_depth_luma > 14) return XB200_ERR_UNSUPPORTED;
    a.inter_done = a.inter_count = nullptr;
    memset(&a, 0, sizeof(a));
    a.cur.y = cur->y; a.cur.u = cur->u; a.cur.v = cur->v;
    for (int l = 0; l < 2; l++) {
        xb200_pic *const *lst = l ? l1 : l0;
        const int n = l ? n1 : n0;
        for (int i = 0; i < n; i++) {
            if (!lst || !lst[i] || lst[i]->w != cur->w || lst[i]->h != cur->h) return XB200_ERR_INVALID_ARGUMENT;
            a.ref_y[l][i] = lst[i]->y; a.ref_u[l][i] = lst[i]->u; a.ref_v[l][i] = lst[i]->v;
            a.ref_poc[l][i] = lst[i]->poc;
            a.ref_tmap[l * XB_MAX_REFS + i] = lst[i]->d_tmaps;
        }
    }
    a.s_l = cur->s_l; a.s_c = cur->s_c; a.w = cur->w; a.h = cur->h;
    a.bd_l = prm->bit_depth_luma; a.bd_c = prm->bit_depth_chroma;
    a.log2_ctu = prm->log2_ctu;
    a.w_ctu = (cur->w + (1 << a.log2_ctu) - 1) >> a.log2_ctu;
    a.n_ctu = a.w_ctu * ((cur->h + (1 << a.log2_ctu) - 1) >> a.log2_ctu);
    a.n_peer = 0;
    if (prm->ctu_rows > 0) {        // band mode: this launch covers CTU rows [ctu_row0, ctu_row0 + ctu_rows)
        if (prm->ctu_row0 < 0 || (prm->ctu_row0 + prm->ctu_rows) * a.w_ctu > a.n_ctu) return XB200_ERR_INVALID_ARGUMENT;
        a.ctu_row0 = prm->ctu_row0;
        a.n_ctu = prm->ctu_rows * a.w_ctu;
        a.n_peer = cur->n_peer;
        a.peer_maps = c->peer_maps ? 1 : 0;
        for (int k = 0; k < cur->n_peer; k++) a.peer_delta[k] = (long long)((char *)cur->peer_base[k] - (char *)cur->buf);
    }
    a.main_tables = prm->tool_admvp ? 1 : 0;
    a.iqt = prm->tool_iqt ? 1 : 0;
    a.eipd = prm->tool_eipd ? 1 : 0;
    a.ats = prm->tool_ats ? 1 : 0;
    a.htdf = prm->tool_htdf ? 1 : 0;
    a.ibc = prm->tool_ibc ? 1 : 0;
    a.constrained = prm->constrained_intra_pred ? 1 : 0;
    a.dmvr = prm->tool_dmvr ? 1 : 0;
    a.poc = prm->poc;
    a.affine = prm->tool_affine ? 1 : 0;
    a.slice_qp = prm->slice_qp;
    a.map_mv = cur->map_mv; a.map_unrefined_mv = cur->map_unrefined_mv; a.map_refi = cur->map_refi; a.map_scu = cur->map_scu; a.map_edge = cur->map_edge;
    a.map_order = prm->tool_suco ? cur->map_order : nullptr;
    a.w_scu = cur->w_scu; a.h_scu = cur->h_scu;
    return XB200_OK;
}

int xb200_recon_frame_dev(xb200_ctx *c, const XB200_PARAMS *prm, xb200_pic *cur,
                          xb200_pic *const *l0, int n0, xb200_pic *const *l1, int n1,
                          const void *d_cus, int n_cu, const void *d_ctu_first, int n_ctu,
                          const void *d_ext, int n_ext, const void *d_coef, size_t n_coef, int has_intra, int max_cu_per_ctu)
{
    XbFrameArgs a;
    int r = fill_args(c, prm, cur, l0, n0, l1, n1, a);
    if (r < 0) return r;
    (void)n_ext; (void)n_coef;
    if (n_ctu != a.n_ctu || n_cu < 0 || !d_cus || !d_ctu_first) return XB200_ERR_INVALID_ARGUMENT;
    if (has_intra && prm->ctu_rows > 0) return XB200_ERR_UNSUPPORTED;          // the wavefront crosses band boundaries
    // throughput kernel (xb_recon2.cuh): 64x64 CTUs, Baseline or IQT transform; with ATS / DMVR / affine enabled it takes the CTUs that hold
    // no such CU and the generic kernel (xb_recon.cuh) the others
    // The throughput kernel sizes its on-chip work lists for at most 256 CUs per CTU (a 64x64 CTU of 4x4 CUs).  Local dual tree nodes
    // can exceed that (four 4x4 luma leaves + one chroma CU per 8x8 node: up to 320): such pictures go through the generic kernel.
    const bool many_cus = max_cu_per_ctu > 256 || (max_cu_per_ctu <= 0 && (has_intra & XB200_HAS_DUAL_TREE));
    const bool fast = a.log2_ctu == 6 && !c->force_generic && !many_cus;
    const bool mixed = fast && (a.ats || a.dmvr || a.affine || (has_intra & XB200_HAS_DUAL_TREE));     // dual-tree CUs need per-plane owners
    a.dispatch = mixed ? 1 : 0;
    if (a.n_peer > 0 && (!fast || mixed || a.iqt)) return XB200_ERR_UNSUPPORTED;        // peer stores exist in the throughput kernel only; use the all-gather exchange
    if (((uintptr_t)d_coef & 15) || ((uintptr_t)d_cus & 15)) return XB200_ERR_INVALID_ARGUMENT;   // 16-byte vector / bulk-copy access
    if (cur->poc != prm->poc) cur->poc = prm->poc;
    a.cus = (const XB200_CU *)d_cus;
    a.ctu_first = (const uint32_t *)d_ctu_first;
    a.coef = (const int16_t *)d_coef;
    a.ext = (const XB200_CU_EXT *)d_ext;
    cudaSetDevice(c->device);
    // With Main tools and wavefront work in one picture the wavefront kernel runs BESIDE the generic inter kernel on a second stream: it
    // keeps a fraction of the SMs busy (a 126-step dependency chain), the generic kernel the rest.  The generic kernel raises one flag per
    // CTU, the wavefront kernel waits for the flags of the CTUs it reads - its own, its four neighbours, the CTUs under an IBC source block
    // (xb_intra.cuh).  Not with constrained intra prediction (reads map_scu of neighbours through the read-only path).
    bool overlap = has_intra && mixed && !a.constrained && !c->no_overlap && a.n_peer == 0;
    if (overlap) {
        xb200_ctx *none = nullptr;
        overlap = g_overlap_owner.compare_exchange_strong(none, c) || none == c;
    }
    if (has_intra) {
        // synchronisation words of the wavefront kernel: ticket, one `done` flag per CTU, one `inter_done` flag per CTU
        if (c->sync_cap < 2 * a.n_ctu + 2) {
            CK(c, cudaStreamSynchronize(c->stream));
            if (c->d_sync) cudaFree(c->d_sync);
            c->d_sync = nullptr; c->sync_cap = 0;
            CK(c, cudaMalloc((void **)&c->d_sync, sizeof(int) * (2 * a.n_ctu + 2)));
            c->sync_cap = 2 * a.n_ctu + 2;
        }
        if (c->order_w != a.w_ctu || c->order_n != a.n_ctu) {
            // wavefront order: sort CTU addresses by x + 2y (a stable counting pass per index)
            const int wc = a.w_ctu, hc = a.n_ctu / a.w_ctu;
            int *h = (int *)malloc(sizeof(int) * a.n_ctu);
            if (!h) return XB200_ERR_OUT_OF_MEMORY;
            int k = 0;
            for (int d = 0; d <= (wc - 1) + 2 * (hc - 1); d++)
                for (int y = 0; y < hc; y++) { const int x = d - 2 * y; if (x >= 0 && x < wc) h[k++] = y * wc + x; }
            cudaError_t e = cudaStreamSynchronize(c->stream);
            if (c->d_order) cudaFree(c->d_order);
            c->d_order = nullptr; c->order_w = c->order_n = 0;
            if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_order, sizeof(int) * a.n_ctu);
            if (e == cudaSuccess) e = cudaMemcpy(c->d_order, h, sizeof(int) * a.n_ctu, cudaMemcpyHostToDevice);
            free(h);
            CK(c, e);
            c->order_w = a.w_ctu; c->order_n = a.n_ctu;
        }
        CK(c, cudaMemsetAsync(c->d_sync, 0, sizeof(int) * (2 * a.n_ctu + 2), c->stream));
        if (!c->d_err) { CK(c, cudaMalloc((void **)&c->d_err, sizeof(int))); CK(c, cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream)); }
        if (overlap && !c->stream_w) {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            CK(c, cudaStreamCreateWithPriority(&c->stream_w, cudaStreamNonBlocking, hi));
            CK(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
            CK(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        }
    }
    a.inter_done = overlap ? c->d_sync + 1 + a.n_ctu : nullptr;
    a.inter_count = overlap ? c->d_sync + 1 + 2 * a.n_ctu : nullptr;
    if (fast) {
        int max_cu = max_cu_per_ctu > 0 ? (max_cu_per_ctu > 256 ? 256 : max_cu_per_ctu) : 256;
        max_cu = (max_cu + 15) & ~15;
        const bool bi = n1 > 0;
        const bool peer = a.n_peer > 0;
        // warp-slot variant (three CTAs per SM, no block-wide step in the prediction stages) for pictures of large CUs; the round variant
        // shares its per-tile passes between two tiles, which wins when most CUs are 8x8 and smaller: more than 32 CUs per CTU on average
        // (xb_recon2.cuh; A/B in profiles/r2: quadtree pictures of ~20 CUs per CTU 160 vs 164 us for the slots, 8x8 pictures 160 vs 146 against them)
        const bool ws = !peer && (c->v2_variant == 2 || (c->v2_variant == 0 && (long long)n_cu <= 32LL * a.n_ctu));
        const xb::R2Layout L = xb::R2Layout::make(bi ? 2 : 1, max_cu, peer, ws);
        const dim3 grid(a.w_ctu, a.n_ctu / a.w_ctu);
        // launched early: resident while the kernel before (the previous picture's padding, an expansion kernel) drains; waits before its first global access
#define XB_V2(BI_, IQT_, DISP_) do { if (ws) xb_launch_early(xb::k_recon_inter_v2<BI_, false, IQT_, DISP_, true>, grid, dim3(xb::kR2Threads), L.total, c->stream, a, max_cu); \
                                     else xb_launch_early(xb::k_recon_inter_v2<BI_, false, IQT_, DISP_, false>, grid, dim3(xb::kR2Threads), L.total, c->stream, a, max_cu); } while (0)
        if (peer) {
            if (bi) xb::k_recon_inter_v2<true, true><<<grid, xb::kR2Threads, L.total, c->stream>>>(a, max_cu);
            else    xb::k_recon_inter_v2<false, true><<<grid, xb::kR2Threads, L.total, c->stream>>>(a, max_cu);
        } else if (a.iqt && mixed) { if (bi) XB_V2(true, true, true); else XB_V2(false, true, true); }
        else if (a.iqt)            { if (bi) XB_V2(true, true, false); else XB_V2(false, true, false); }
        else if (mixed)            { if (bi) XB_V2(true, false, true); else XB_V2(false, false, true); }
        else                       { if (bi) XB_V2(true, false, false); else XB_V2(false, false, false); }
#undef XB_V2
        if (mixed) { c->launches++; CK(c, cudaGetLastError()); }
    }
    if (overlap) {              // the wavefront stream continues from here: the throughput kernel is complete, the generic kernel is not
        CK(c, cudaEventRecord(c->ev_fork, c->stream));
        CK(c, cudaStreamWaitEvent(c->stream_w, c->ev_fork, 0));
    }
    if (!fast || mixed) {
        const size_t smem = xb::ReconSmem::bytes(a.log2_ctu);
        if (a.iqt) xb::k_recon_inter<true><<<a.n_ctu, xb::kReconThreads, smem, c->stream>>>(a);
        else       xb::k_recon_inter<false><<<a.n_ctu, xb::kReconThreads, smem, c->stream>>>(a);
    }
    c->launches++;
    CK(c, cudaGetLastError());
    if (has_intra) {
        // intra CUs: CTU wavefront over the picture the inter kernels complete
        xb::IntraSync sy{c->d_sync, c->d_err, c->d_sync + 1, c->d_order};
        // Two wavefront CTAs share an SM with CTUs of 64 samples and less (91 KB each): +30..50 % pictures/s with several pictures in flight and
        // -22 % on a P picture's wavefront pass, but a lone I picture's ~40 chained CTAs then also pair up on SMs and its critical path gets
        // 3..9 % longer.  XB200_WAVEFRONT_ALONE=1 asks for more than half an SM's shared memory, i.e. one CTA per SM (the latency setting).
        size_t sm = xb::IntraSmem::bytes(a.log2_ctu);
        if (c->wavefront_alone && sm < (size_t)116 * 1024) sm = (size_t)116 * 1024;
        // persistent CTAs (xb_intra.cuh): dense dependencies (I pictures) -> about as many CTAs as the x + 2y wavefront is wide
        const int h_ctu = a.n_ctu / a.w_ctu;
        int grid = a.n_ctu;
        if (has_intra & XB200_HAS_DENSE_WAVEFRONT) { const int wide = ((a.w_ctu + 1) / 2 < h_ctu ? (a.w_ctu + 1) / 2 : h_ctu) + 8; if (wide < grid) grid = wide; }
        // beside the generic kernel: at most one wavefront CTA per two SMs, so that CTAs of the kernel it waits for always find a free SM
        if (overlap && grid > c->sm_count / 2) grid = c->sm_count / 2;
        cudaStream_t ws = overlap ? c->stream_w : c->stream;
        if (a.iqt) xb::k_recon_intra<true><<<grid, xb::kIntraThreads, sm, ws>>>(a, sy);
        else       xb::k_recon_intra<false><<<grid, xb::kIntraThreads, sm, ws>>>(a, sy);
        if (overlap) {
            CK(c, cudaEventRecord(c->ev_join, c->stream_w));
            CK(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        }
        c->wave_used = true;
        c->launches++;
        CK(c, cudaGetLastError());
    }
    return XB200_OK;
}

// grab the next staging slot, make sure its previous consumer has finished, and size it
static int stage_acquire(xb200_ctx *c, size_t bytes, Staging **out)
{
    Staging &s = c->ring[c->ring_pos];
    c->ring_pos = (c->ring_pos + 1) % 3;
    if (!s.done) CK(c, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    if (s.busy) { CK(c, cudaEventSynchronize(s.done)); s.busy = false; }
    if (s.cap < bytes) {
        // a failed allocation must leave the slot empty, not pointing at freed memory with a larger capacity
        if (s.pinned) cudaFreeHost(s.pinned);
        if (s.dev) cudaFree(s.dev);
        s.pinned = nullptr; s.dev = nullptr; s.cap = 0;
        const size_t want = bytes + bytes / 4 + 4096;
        CK(c, cudaMallocHost(&s.pinned, want));
        if (cudaMalloc(&s.dev, want) != cudaSuccess) {
            cudaFreeHost(s.pinned); s.pinned = nullptr; s.dev = nullptr;
            snprintf(c->err, sizeof(c->err), "cudaMalloc(%zu) failed for the staging ring", want);
            return XB200_ERR_OUT_OF_MEMORY;
        }
        s.cap = want;
    }
    *out = &s;
    return XB200_OK;
}

// sparse -> dense coefficient stream on the device: one CTA per chunk of XB200_SPARSE_CHUNK int16 (zero it, then scatter the chunk's
// non-zero levels).  The dense stream never crosses PCIe; writing and re-reading it in HBM costs ~10 us per 4K picture.
__global__ void __launch_bounds__(256) k_expand_coef(const uint32_t *__restrict__ entries, const uint32_t *__restrict__ chunk_first, int16_t *__restrict__ dense, size_t n_coef)
{
    const size_t base = (size_t)blockIdx.x * XB200_SPARSE_CHUNK;
    const int n = (int)min((size_t)XB200_SPARSE_CHUNK, n_coef - base);
    int16_t *d = dense + base;
    for (int i = threadIdx.x; i < (n >> 3); i += blockDim.x) ((int4 *)d)[i] = make_int4(0, 0, 0, 0);
    for (int i = (n & ~7) + threadIdx.x; i < n; i += blockDim.x) d[i] = 0;
    __syncthreads();
    const uint32_t e0 = chunk_first[blockIdx.x], e1 = chunk_first[blockIdx.x + 1];
    for (uint32_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const uint32_t v = __ldg(entries + e);
        const int pos = (int)(v & 0xffffu);
        if (pos < n) d[pos] = (int16_t)(v >> 16);
    }
}

static int recon_frame_host(xb200_ctx *c, const XB200_PARAMS *prm, xb200_pic *cur,
                            xb200_pic *const *l0, int n0, xb200_pic *const *l1, int n1,
                            const XB200_CU *cus, int n_cu, const uint32_t *ctu_first, int n_ctu,
                            const XB200_CU_EXT *ext, int n_ext, const int16_t *coef, size_t n_coef,
                            const uint32_t *sp_entries, size_t n_entries, const uint32_t *sp_chunk_first)
{
    const bool sparse = sp_entries != nullptr || sp_chunk_first != nullptr;
    const size_t n_chunks = (n_coef + XB200_SPARSE_CHUNK - 1) / XB200_SPARSE_CHUNK;
    if (!c || !prm || !cur || !cus || !ctu_first || n_cu < 0 || n_ctu <= 0) return XB200_ERR_INVALID_ARGUMENT;
    if (sparse) {
        if (!sp_chunk_first || (n_entries && !sp_entries)) return XB200_ERR_INVALID_ARGUMENT;
        if (sp_chunk_first[0] != 0 || sp_chunk_first[n_chunks] != n_entries) { snprintf(c->err, sizeof(c->err), "chunk_first does not span the entries"); return XB200_ERR_INVALID_ARGUMENT; }
        for (size_t k = 0; k < n_chunks; k++)
            if (sp_chunk_first[k + 1] < sp_chunk_first[k]) { snprintf(c->err, sizeof(c->err), "chunk_first is not monotonic at chunk %zu", k); return XB200_ERR_INVALID_ARGUMENT; }
    }
    if (prm->log2_ctu < 5 || prm->log2_ctu > 7) return XB200_ERR_INVALID_ARGUMENT;
    cudaSetDevice(c->device);
    // Everything the kernels turn into an address is checked here, where the lists are still host memory and every CU is visited anyway:
    // a malformed work list must fail the call, not write outside the picture or overrun a CTA's shared memory.  (The _dev entry point
    // takes device-resident lists from a producer that is trusted to have done the same.)
    {
        const int ctu = 1 << prm->log2_ctu, w_ctu = (prm->w + ctu - 1) >> prm->log2_ctu;
        const int row0 = prm->ctu_rows > 0 ? prm->ctu_row0 : 0;
        auto bad = [&](int i, const char *what) {
            snprintf(c->err, sizeof(c->err), "CU %d (mode %d, %dx%d at %d,%d): %s", i, cus[i].mode, 1 << cus[i].log2w, 1 << cus[i].log2h, cus[i].x, cus[i].y, what);
            return XB200_ERR_INVALID_ARGUMENT;
        };
        if (ctu_first[0] != 0 || ctu_first[n_ctu] != (uint32_t)n_cu) { snprintf(c->err, sizeof(c->err), "ctu_first[0] / ctu_first[n_ctu] do not span the CU list"); return XB200_ERR_INVALID_ARGUMENT; }
        size_t run = n_cu > 0 ? cus[0].coef_off : 0;
        for (int t = 0; t < n_ctu; t++) {
            if (ctu_first[t + 1] < ctu_first[t] || ctu_first[t + 1] > (uint32_t)n_cu) { snprintf(c->err, sizeof(c->err), "ctu_first is not monotonic at CTU %d", t); return XB200_ERR_INVALID_ARGUMENT; }
            const int cx = (t % w_ctu) << prm->log2_ctu, cy = (t / w_ctu + row0) << prm->log2_ctu;
            for (int i = (int)ctu_first[t]; i < (int)ctu_first[t + 1]; i++) {
                const XB200_CU &u = cus[i];
                if (u.log2w < 2 || u.log2w > 7 || u.log2h < 2 || u.log2h > 7) return bad(i, "size outside 4..128");
                const int w = 1 << u.log2w, h = 1 << u.log2h;
                if ((u.x & 3) || (u.y & 3) || u.x < cx || u.y < cy || u.x + w > cx + ctu || u.y + h > cy + ctu) return bad(i, "not inside the CTU it is listed under");
                if (u.x >= prm->w || u.y >= prm->h) return bad(i, "outside the picture");
                if (u.mode != XB200_MODE_INTRA && u.mode != XB200_MODE_INTER && u.mode != XB200_MODE_IBC && u.mode != XB200_MODE_AFFINE) return bad(i, "unknown mode");
                if (u.mode == XB200_MODE_INTRA || u.mode == XB200_MODE_AFFINE) {
                    uint32_t ei;
                    memcpy(&ei, u.mv[1], 4);
                    if (!ext || ei >= (uint32_t)n_ext) return bad(i, "extension record index outside the XB200_CU_EXT array");
                }
                if (u.mode == XB200_MODE_IBC) {
                    // xevdm_IBC_mc copies from the part of the picture decoded so far; the wavefront kernel orders CTUs on their left / upper neighbours
                    const int rx = u.x + u.mv[0][0], ry = u.y + u.mv[0][1];
                    if (rx < 0 || ry < 0 || rx + w > prm->w || ry + h > prm->h) return bad(i, "block vector leaves the picture");
                    if (ry + h > cy + ctu || rx + w > cx + ctu) return bad(i, "block vector reaches below / right of the CU's CTU");
                }
                // coefficient blocks: decoding order, planes without coefficients absent, every block padded to 8 entries
                int tlw = u.log2w, tlh = u.log2h;
                const int ai = (u.mode == XB200_MODE_INTER || u.mode == XB200_MODE_AFFINE) && prm->tool_ats ? XB200_ATS_INTER_IDX(u.ats) : 0;
                if (ai > 4) return bad(i, "ats_inter_idx > 4");
                if (ai) { const int sh = ai >= 3 ? 2 : 1; if (ai == 2 || ai == 4) tlh -= sh; else tlw -= sh; if (tlw < 2 || tlh < 2) return bad(i, "ats_inter split of a CU that is too small"); }
                const size_t n = (size_t)1 << (tlw + tlh);
                size_t len = 0;
                if (u.cbf & 0x00f) len += (n + 7) & ~(size_t)7;
                if (u.cbf & 0x0f0) len += ((n >> 2) + 7) & ~(size_t)7;
                if (u.cbf & 0xf00) len += ((n >> 2) + 7) & ~(size_t)7;
                if (len) {
                    if (u.coef_off != run) return bad(i, "coef_off does not continue the coefficient stream (blocks must follow the decoding order without gaps)");
                    if ((!coef && !sparse) || run + len > n_coef) return bad(i, "coefficient blocks run past the end of the stream");
                    run += len;
                } else if (u.coef_off != run && u.coef_off != 0) return bad(i, "coef_off of a CU without coefficients must continue the stream");
            }
        }
    }
    // one staging blob: [cus][ctu_first][ext][coef], each 256-byte aligned
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t b_cu = al((size_t)n_cu * sizeof(XB200_CU)), b_first = al((size_t)(n_ctu + 1) * 4);
    const size_t b_ext = al((size_t)(n_ext > 0 ? n_ext : 1) * sizeof(XB200_CU_EXT)), b_coef = al(n_coef * 2 + 2);
    // sparse form: [chunk_first][entries] travel, the dense stream exists in the device half of the slot only
    const size_t b_cf = sparse ? al((n_chunks + 1) * 4) : 0, b_en = sparse ? al(n_entries * 4 + 4) : 0;
    Staging *s;
    int r = stage_acquire(c, b_cu + b_first + b_ext + b_coef + b_cf + b_en, &s);
    if (r < 0) return r;
    unsigned char *hp = (unsigned char *)s->pinned, *dp = (unsigned char *)s->dev;
    // Page-locked caller memory (xb200_host_alloc, cudaHostRegister, ...) is DMA'd straight to the device slot; pageable
    // memory goes through the context's pinned staging buffer first.
    auto h2d = [&](size_t off, const void *src, size_t bytes) -> cudaError_t {
        if (!src || !bytes) return cudaSuccess;
        cudaPointerAttributes at;
        const bool pinned = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost;
        if (!pinned) { cudaGetLastError(); memcpy(hp + off, src, bytes); src = hp + off; }
        return cudaMemcpyAsync(dp + off, src, bytes, cudaMemcpyHostToDevice, c->stream);
    };
    int has_intra = 0, max_cu = 0, any_l1 = 0, n_wave = 0;
    for (int i = 0; i < n_cu; i++) {
        const bool intra = xb_wavefront_mode(cus[i].mode);
        n_wave += intra || (prm->tool_htdf && (cus[i].cbf & 15));
        // local dual tree (TREE_L / TREE_C CUs, src_main/xevdm.c:1828-1846): intra-only nodes; inter CUs always carry all three planes and
        // IBC needs luma (xevdm.c:1113-1122); the cbf bits of a plane the CU does not carry must be clear (its coefficient block is absent)
        const int pl = cus[i].flags & (XB200_CUF_LUMA | XB200_CUF_CHROMA);
        if (pl != (XB200_CUF_LUMA | XB200_CUF_CHROMA)) {
            if (!pl || !intra || (cus[i].mode == XB200_MODE_IBC && !(pl & XB200_CUF_LUMA))) {
                snprintf(c->err, sizeof(c->err), "CU %d: planes 0x%x with mode %d (single-plane CUs are intra / luma IBC only)", i, pl, cus[i].mode);
                return XB200_ERR_INVALID_ARGUMENT;
            }
            if ((!(pl & XB200_CUF_LUMA) && (cus[i].cbf & 0x00f)) || (!(pl & XB200_CUF_CHROMA) && (cus[i].cbf & 0xff0))) {
                snprintf(c->err, sizeof(c->err), "CU %d: cbf 0x%x names a plane the CU does not carry (flags 0x%x)", i, cus[i].cbf, cus[i].flags);
                return XB200_ERR_INVALID_ARGUMENT;
            }
            has_intra |= XB200_HAS_DUAL_TREE;
        }
        has_intra |= (intra || (prm->tool_htdf && (cus[i].cbf & 15))) ? XB200_HAS_INTRA : 0;       // HTDF-filtered inter CUs are finished by the wavefront kernel too
        if (intra) continue;
        any_l1 |= cus[i].refi[1] >= 0;
        // a reference index outside the lists the caller supplied would dereference a missing picture on the device
        if (cus[i].refi[0] >= n0 || cus[i].refi[1] >= n1 || (cus[i].refi[0] < 0 && cus[i].refi[1] < 0)) {
            snprintf(c->err, sizeof(c->err), "CU %d (mode %d at %d,%d): reference indices %d / %d outside the lists (%d / %d pictures)", i, cus[i].mode, cus[i].x, cus[i].y,
                     cus[i].refi[0], cus[i].refi[1], n0, n1);
            return XB200_ERR_INVALID_ARGUMENT;
        }
    }
    if (2 * n_wave > n_cu) has_intra |= XB200_HAS_DENSE_WAVEFRONT;       // I picture (or nearly): size the wavefront grid for the wavefront, not for the picture
    if (!any_l1) n1 = 0;          // P picture: no CU predicts from list 1 (selects the single-list kernel)
    for (int i = 0; i < n_ctu; i++) { const int d = (int)(ctu_first[i + 1] - ctu_first[i]); if (d > max_cu) max_cu = d; }
    CK(c, h2d(0, cus, (size_t)n_cu * sizeof(XB200_CU)));
    CK(c, h2d(b_cu, ctu_first, (size_t)(n_ctu + 1) * 4));
    if (n_ext > 0) CK(c, h2d(b_cu + b_first, ext, (size_t)n_ext * sizeof(XB200_CU_EXT)));
    if (!sparse) CK(c, h2d(b_cu + b_first + b_ext, coef, n_coef * 2));
    else if (n_coef) {
        const size_t o_cf = b_cu + b_first + b_ext + b_coef, o_en = o_cf + b_cf;
        CK(c, h2d(o_cf, sp_chunk_first, (n_chunks + 1) * 4));
        CK(c, h2d(o_en, sp_entries, n_entries * 4));
        k_expand_coef<<<(unsigned)n_chunks, 256, 0, c->stream>>>((const uint32_t *)(dp + o_en), (const uint32_t *)(dp + o_cf), (int16_t *)(dp + b_cu + b_first + b_ext), n_coef);
        c->launches++;
        CK(c, cudaGetLastError());
    }
    r = xb200_recon_frame_dev(c, prm, cur, l0, n0, l1, n1, dp, n_cu, dp + b_cu, n_ctu, dp + b_cu + b_first, n_ext,
                              dp + b_cu + b_first + b_ext, n_coef, has_intra, max_cu);
    CK(c, cudaEventRecord(s->done, c->stream));
    s->busy = true;
    return r;
}

int xb200_recon_frame(xb200_ctx *c, const XB200_PARAMS *prm, xb200_pic *cur,
                      xb200_pic *const *l0, int n0, xb200_pic *const *l1, int n1,
                      const XB200_CU *cus, int n_cu, const uint32_t *ctu_first, int n_ctu,
                      const XB200_CU_EXT *ext, int n_ext, const int16_t *coef, size_t n_coef)
{
    return recon_frame_host(c, prm, cur, l0, n0, l1, n1, cus, n_cu, ctu_first, n_ctu, ext, n_ext, coef, n_coef, nullptr, 0, nullptr);
}

int xb200_recon_frame_sparse(xb200_ctx *c, const XB200_PARAMS *prm, xb200_pic *cur,
                             xb200_pic *const *l0, int n0, xb200_pic *const *l1, int n1,
                             const XB200_CU *cus, int n_cu, const uint32_t *ctu_first, int n_ctu,
                             const XB200_CU_EXT *ext, int n_ext, const uint32_t *entries, size_t n_entries, const uint32_t *chunk_first, size_t n_coef)
{
    if (!chunk_first) return XB200_ERR_INVALID_ARGUMENT;
    return recon_frame_host(c, prm, cur, l0, n0, l1, n1, cus, n_cu, ctu_first, n_ctu, ext, n_ext, nullptr, n_coef, entries, n_entries, chunk_first);
}

// ---- in-loop filters / padding ----------------------------------------------------------------------------------------
int xb200_pad(xb200_ctx *c, xb200_pic *p)
{
    if (!c || !p) return XB200_ERR_INVALID_ARGUMENT;
    cudaSetDevice(c->device);
    xb::launch_pad(p->y, p->s_l, p->w, p->h, p->pad_l, p->u, p->v, p->s_c, p->w_c, p->h_c, p->pad_c, c->stream);
    c->launches++;
    CK(c, cudaGetLastError());
    return XB200_OK;
}

int xb200_alf(xb200_ctx *c, const XB200_PARAMS *prm, xb200_pic *p, const XB200_ALF *alf, const uint8_t *ctb_flag_luma)
{
    if (!c || !prm || !p || !alf) return XB200_ERR_INVALID_ARGUMENT;
    if (prm->w != p->w || prm->h != p->h || prm->log2_ctu < 5 || prm->log2_ctu > 7 || prm->bit_depth_luma < 8 || prm->bit_depth_luma > 14) {
        snprintf(c->err, sizeof(c->err), "xb200_alf: picture %dx%d / params %dx%d, log2_ctu %d, bit depth %d", p->w, p->h, prm->w, prm->h, prm->log2_ctu, prm->bit_depth_luma);
        return XB200_ERR_INVALID_ARGUMENT;
    }
    if (prm->chroma_format_idc != 1) return XB200_ERR_UNSUPPORTED;
    if (!alf->enable[0] && !alf->enable[1] && !alf->enable[2]) return XB200_OK;   // alf_process :1172
    cudaSetDevice(c->device);
    // the pre-ALF copy has the layout of the picture's own buffer (padded planes back to back), so plane pointers translate by one offset
    const size_t n_copy = p->luma_elems + 2 * p->chroma_elems;
    if (c->alf_cap < n_copy) {
        CK(c, cudaStreamSynchronize(c->stream));
        if (c->alf_copy) cudaFree(c->alf_copy);
        c->alf_copy = nullptr; c->alf_cap = 0;
        CK(c, cudaMalloc(&c->alf_copy, n_copy * sizeof(pel)));
        c->alf_cap = n_copy;
    }
    xb::AlfArgs a;
    a.sy = c->alf_copy + (p->y - p->buf); a.su = c->alf_copy + (p->u - p->buf); a.sv = c->alf_copy + (p->v - p->buf);
    a.dy = p->y; a.du = p->u; a.dv = p->v;
    a.s_l = p->s_l; a.s_c = p->s_c; a.w = p->w; a.h = p->h; a.log2_ctu = prm->log2_ctu; a.bd = prm->bit_depth_luma;
    a.w_ctu = (p->w + (1 << prm->log2_ctu) - 1) >> prm->log2_ctu;
    a.ctb_flag = nullptr;
    a.n_tile_cols = c->tile_cols; a.n_tile_rows = c->tile_rows; a.tile_across = c->tile_across;
    memcpy(a.tile_col_bd, c->tile_col_bd, sizeof(a.tile_col_bd)); memcpy(a.tile_row_bd, c->tile_row_bd, sizeof(a.tile_row_bd));
    memcpy(a.coef_c, alf->coef_chroma, sizeof(a.coef_c));
    memcpy(a.enable, alf->enable, 3);
    if (alf->enable[0]) {
        // the 100 luma filters a 4x4 block can select: 25 classes x 4 transposes, coefficients already permuted, 16 int16 each; the table
        // lives on the device and is uploaded again only when the coefficients change (a new APS)
        if (!c->alf_tab_dev) {
            CK(c, cudaMalloc(&c->alf_tab_dev, xb::kAlfTabBytes));
            CK(c, cudaMallocHost(&c->alf_tab_pinned, xb::kAlfTabBytes));
            CK(c, cudaEventCreateWithFlags(&c->alf_tab_done, cudaEventDisableTiming));
            c->alf_tab_valid = false;
        }
        if (!c->alf_tab_valid || memcmp(c->alf_tab_coef, alf->coef_luma, sizeof(c->alf_tab_coef)) != 0) {
            static const uint8_t perm[4][13] = {{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12}, {9, 4, 10, 8, 1, 5, 11, 7, 3, 0, 2, 6, 12},
                                                {0, 3, 2, 1, 8, 7, 6, 5, 4, 9, 10, 11, 12}, {9, 8, 10, 4, 3, 7, 11, 5, 1, 0, 2, 6, 12}};
            if (c->alf_tab_valid) CK(c, cudaEventSynchronize(c->alf_tab_done));      // the previous upload has left the pinned buffer
            int16_t *tab = (int16_t *)c->alf_tab_pinned;
            for (int cls = 0; cls < 25; cls++)
                for (int tr = 0; tr < 4; tr++) {
                    int16_t *f = tab + (cls * 4 + tr) * 16;
                    for (int i = 0; i < 13; i++) f[i] = alf->coef_luma[cls][perm[tr][i]];
                    f[13] = f[14] = f[15] = 0;
                }
            CK(c, cudaMemcpyAsync(c->alf_tab_dev, c->alf_tab_pinned, xb::kAlfTabBytes, cudaMemcpyHostToDevice, c->stream));
            CK(c, cudaEventRecord(c->alf_tab_done, c->stream));
            memcpy(c->alf_tab_coef, alf->coef_luma, sizeof(c->alf_tab_coef));
            c->alf_tab_valid = true;
        }
        a.ftab = (const int4 *)c->alf_tab_dev;
    }
    if (ctb_flag_luma && alf->enable[0]) {
        const int n = a.w_ctu * ((p->h + (1 << prm->log2_ctu) - 1) >> prm->log2_ctu);
        if (c->alf_flags_cap < n) {
            CK(c, cudaStreamSynchronize(c->stream));
            if (c->alf_flags_pinned) cudaFreeHost(c->alf_flags_pinned);
            if (c->alf_flags_dev) cudaFree(c->alf_flags_dev);
            c->alf_flags_pinned = c->alf_flags_dev = nullptr; c->alf_flags_cap = 0;
            CK(c, cudaMallocHost(&c->alf_flags_pinned, n));
            CK(c, cudaMalloc(&c->alf_flags_dev, n));
            if (!c->alf_flags_done) CK(c, cudaEventCreateWithFlags(&c->alf_flags_done, cudaEventDisableTiming));
            c->alf_flags_cap = n;
        } else {
            CK(c, cudaEventSynchronize(c->alf_flags_done));    // previous upload has left the pinned buffer
        }
        memcpy(c->alf_flags_pinned, ctb_flag_luma, n);
        CK(c, cudaMemcpyAsync(c->alf_flags_dev, c->alf_flags_pinned, n, cudaMemcpyHostToDevice, c->stream));
        CK(c, cudaEventRecord(c->alf_flags_done, c->stream));
        a.ctb_flag = c->alf_flags_dev;
    }
    {   // sample rows of the enabled planes, whole rows (a row with its padding is a multiple of 16 bytes and the rows of a plane are contiguous)
        xb::AlfCopyArgs ca;
        const pel *rows[3] = {p->y - p->pad_l, p->u - p->pad_c, p->v - p->pad_c};
        for (int pl = 0; pl < 3; pl++) {
            ca.src[pl] = (const int4 *)rows[pl];
            ca.dst[pl] = (int4 *)(c->alf_copy + (rows[pl] - p->buf));
            ca.n[pl] = alf->enable[pl] ? (unsigned)((pl ? (size_t)p->s_c * p->h_c : (size_t)p->s_l * p->h) * sizeof(pel) / 16) : 0u;
        }
        xb_launch_early(xb::k_alf_copy, dim3(c->sm_count * 8), dim3(256), 0, c->stream, ca);
        c->launches++;
    }
    const dim3 grid((p->w + xb::kAlfT - 1) / xb::kAlfT, (p->h + xb::kAlfT - 1) / xb::kAlfT);
    xb_launch_early(xb::k_alf, grid, dim3(256), 0, c->stream, a);
    c->launches++;
    CK(c, cudaGetLastError());
    return XB200_OK;
}

int xb200_set_tiles(xb200_ctx *c, int n_cols, const uint16_t *col_bd, int n_rows, const uint16_t *row_bd, int across)
{
    if (!c || n_cols < 1 || n_rows < 1 || n_cols > XB200_MAX_TILE_COLS || n_rows > XB200_MAX_TILE_ROWS || !col_bd || !row_bd) return XB200_ERR_INVALID_ARGUMENT;
    if (col_bd[0] != 0 || row_bd[0] != 0) return XB200_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < n_cols; i++) if (col_bd[i + 1] <= col_bd[i]) return XB200_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < n_rows; i++) if (row_bd[i + 1] <= row_bd[i]) return XB200_ERR_INVALID_ARGUMENT;
    c->tile_cols = n_cols; c->tile_rows = n_rows; c->tile_across = across != 0;
    memcpy(c->tile_col_bd, col_bd, sizeof(uint16_t) * (size_t)(n_cols + 1));
    memcpy(c->tile_row_bd, row_bd, sizeof(uint16_t) * (size_t)(n_rows + 1));
    return XB200_OK;
}

int xb200_set_chroma_qp_table(xb200_ctx *c, const int32_t *tbl)
{
    if (!c || !tbl) return XB200_ERR_INVALID_ARGUMENT;
    for (int k = 0; k < 2; k++)
        for (int i = 0; i < 58; i++) {
            if (tbl[k * 58 + i] < -128 || tbl[k * 58 + i] > 127) return XB200_ERR_INVALID_ARGUMENT;
            c->chroma_qp[k][i] = (int8_t)tbl[k * 58 + i];
        }
    return XB200_OK;
}

int xb200_pic_upload_maps(xb200_ctx *c, xb200_pic *p, const int16_t *map_mv, const int8_t *map_refi, const uint32_t *map_scu, const uint8_t *map_edge)
{
    if (!c || !p) return XB200_ERR_INVALID_ARGUMENT;
    const size_t n = (size_t)p->w_scu * p->h_scu;
    if (map_mv) CK(c, cudaMemcpyAsync(p->map_mv, map_mv, n * 8, cudaMemcpyHostToDevice, c->stream));
    if (map_mv) CK(c, cudaMemcpyAsync(p->map_unrefined_mv, map_mv, n * 8, cudaMemcpyHostToDevice, c->stream));
    if (map_refi) CK(c, cudaMemcpyAsync(p->map_refi, map_refi, n * 2, cudaMemcpyHostToDevice, c->stream));
    if (map_scu) CK(c, cudaMemcpyAsync(p->map_scu, map_scu, n * 4, cudaMemcpyHostToDevice, c->stream));
    if (map_edge) CK(c, cudaMemcpyAsync(p->map_edge, map_edge, n, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return XB200_OK;
}

int xb200_deblock(xb200_ctx *c, const XB200_PARAMS *prm, xb200_pic *cur, xb200_pic *const *l0, int n0,
                  xb200_pic *const *l1, int n1, const uint8_t *edge_flags)
{
    if (!c || !prm || !cur) return XB200_ERR_INVALID_ARGUMENT;
    if (prm->chroma_format_idc != 1) return XB200_ERR_UNSUPPORTED;
    if (n0 < 0 || n1 < 0 || n0 > XB_MAX_REFS || n1 > XB_MAX_REFS) return XB200_ERR_INVALID_ARGUMENT;
    cudaSetDevice(c->device);
    if (edge_flags) {
        CK(c, cudaMemcpyAsync(cur->map_edge, edge_flags, (size_t)cur->w_scu * cur->h_scu, cudaMemcpyHostToDevice, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));      // the caller's buffer may be pageable and short-lived
    }
    xb::DbkArgs a;
    a.y = cur->y; a.u = cur->u; a.v = cur->v; a.s_l = cur->s_l; a.s_c = cur->s_c; a.w = cur->w; a.h = cur->h;
    a.w_scu = cur->w_scu; a.h_scu = cur->h_scu;
    a.bd_l = prm->bit_depth_luma; a.bd_c = prm->bit_depth_chroma; a.qp_u_offset = prm->qp_u_offset; a.qp_v_offset = prm->qp_v_offset;
    // ADDB compares the vectors before DMVR refinement where the DMVR flag is set and map_mv (affine sub-block vectors) elsewhere
    // (xevdm.c:2009-2041,2077-2090, T7); the Baseline-filter walkers read ctx->map_mv throughout (xevdm_df.c:111-124,1143-1166)
    a.map_scu = cur->map_scu; a.map_mv = cur->map_mv; a.map_umv = cur->map_unrefined_mv; a.map_refi = cur->map_refi; a.map_edge = cur->map_edge;
    a.map_order = (prm->tool_suco && !prm->tool_addb) ? cur->map_order : nullptr;
    memcpy(a.cq, c->chroma_qp, sizeof(a.cq));
    a.alpha_offset = prm->deblock_alpha_offset; a.beta_offset = prm->deblock_beta_offset; a.log2_ctu = prm->log2_ctu;
    {   // picture identity of every reference index: first position of the same picture in (list 0 ++ list 1)
        xb200_pic *all[2 * XB_MAX_REFS];
        int n = 0;
        for (int l = 0; l < 2; l++)
            for (int i = 0; i < (l ? n1 : n0); i++) {
                xb200_pic *p = (l ? l1 : l0) ? (l ? l1 : l0)[i] : nullptr;
                int id = -1;
                for (int k = 0; k < n; k++) if (all[k] == p) { id = k; break; }
                if (id < 0) { all[n] = p; id = n++; }
                a.ref_id[l][i] = (int8_t)id;
            }
    }
    if (c->tile_cols * c->tile_rows > 1 && !c->tile_across) {
        // loop_filter_across_tiles_enabled_flag == 0: the edges that separate two tiles are not filtered
        xb::TileEdgeArgs t;
        t.map_edge = cur->map_edge; t.w_scu = cur->w_scu; t.h_scu = cur->h_scu; t.log2_ctu_scu = prm->log2_ctu - 2;
        t.n_cols = c->tile_cols; t.n_rows = c->tile_rows;
        memcpy(t.col_bd, c->tile_col_bd, sizeof(t.col_bd)); memcpy(t.row_bd, c->tile_row_bd, sizeof(t.row_bd));
        if (c->tile_cols > 1) { xb::k_clear_tile_edges<true><<<dim3((cur->h_scu + 255) / 256, c->tile_cols - 1), 256, 0, c->stream>>>(t); c->launches++; }
        if (c->tile_rows > 1) { xb::k_clear_tile_edges<false><<<dim3((cur->w_scu + 255) / 256, c->tile_rows - 1), 256, 0, c->stream>>>(t); c->launches++; }
    }
    xb::launch_deblock(a, prm->tool_addb != 0, c->stream);
    c->launches += 2;
    CK(c, cudaGetLastError());
    return XB200_OK;
}

// ---- output path ----------------------------------------------------------------------------------------------------------
int xb200_pic_pull(xb200_ctx *c, xb200_pic *p, const XB200_DRA *dra, int out_bits, int crop_l, int crop_r, int crop_t, int crop_b,
                   void *y, int sy, void *u, int su, void *v, int sv)
{
    if (!c || !p || !y || !u || !v || (out_bits != 8 && out_bits != 16)) return XB200_ERR_INVALID_ARGUMENT;
    if (crop_l < 0 || crop_r < 0 || crop_t < 0 || crop_b < 0 || ((crop_l | crop_r | crop_t | crop_b) & 1)) return XB200_ERR_INVALID_ARGUMENT;
    const int w = p->w - crop_l - crop_r, h = p->h - crop_t - crop_b;
    if (w <= 0 || h <= 0) return XB200_ERR_INVALID_ARGUMENT;
    cudaSetDevice(c->device);
    const size_t bps = out_bits == 8 ? 1 : 2, ny = (size_t)w * h * bps, nc = (size_t)(w >> 1) * (h >> 1) * bps;
    if (c->out_cap < ny + 2 * nc) {
        CK(c, cudaStreamSynchronize(c->stream));
        if (c->out_buf) cudaFree(c->out_buf);
        c->out_buf = nullptr; c->out_cap = 0;
        CK(c, cudaMalloc(&c->out_buf, ny + 2 * nc));
        c->out_cap = ny + 2 * nc;
    }
    xb::OutArgs a;
    a.y = p->y; a.u = p->u; a.v = p->v; a.s_l = p->s_l; a.s_c = p->s_c;
    a.x0 = crop_l; a.y0 = crop_t; a.w = w; a.h = h;
    a.lut_l = a.lut_c = nullptr;
    if (dra) {
        if (!c->d_dra) CK(c, cudaMalloc((void **)&c->d_dra, 3 * 1024 * sizeof(int)));
        CK(c, cudaMemcpyAsync(c->d_dra, dra, 3 * 1024 * sizeof(int), cudaMemcpyHostToDevice, c->stream));   // luma LUT, then the two chroma LUTs
        CK(c, cudaStreamSynchronize(c->stream));        // the caller's struct may be short-lived
        a.lut_l = c->d_dra; a.lut_c = c->d_dra + 1024;
    }
    a.out_y = c->out_buf; a.out_u = c->out_buf + ny; a.out_v = c->out_buf + ny + nc;
    a.out8 = out_bits == 8;
    const dim3 grid(((w >> 1) + 255) / 256, h >> 1);
    xb::k_output<<<grid, 256, 0, c->stream>>>(a);
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpy2DAsync(y, (size_t)sy * bps, a.out_y, (size_t)w * bps, (size_t)w * bps, h, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpy2DAsync(u, (size_t)su * bps, a.out_u, (size_t)(w >> 1) * bps, (size_t)(w >> 1) * bps, h >> 1, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpy2DAsync(v, (size_t)sv * bps, a.out_v, (size_t)(w >> 1) * bps, (size_t)(w >> 1) * bps, h >> 1, cudaMemcpyDeviceToHost, c->stream));
    return XB200_OK;
}

// ---- peer pictures (CUDA IPC): band mode with the exchange fused into the kernel's stores --------------------------------
int xb200_pic_export(xb200_ctx *c, xb200_pic *p, void *handle64)
{
    if (!c || !p || !handle64) return XB200_ERR_INVALID_ARGUMENT;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaSetDevice(c->device);
    CK(c, cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, p->buf));
    return XB200_OK;
}
int xb200_pic_open_peer(xb200_ctx *c, xb200_pic *p, const void *handle64)
{
    if (!c || !p || !handle64 || p->n_peer >= 7) return XB200_ERR_INVALID_ARGUMENT;
    cudaSetDevice(c->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void *base = nullptr;
    CK(c, cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    p->peer_base[p->n_peer++] = base;
    return XB200_OK;
}

// ---- band exchange -------------------------------------------------------------------------------------------------------
static void band_geometry(xb200_pic *p, int y0, int rows, int &yc0, int &rc, int &s0, int &sr)
{
    yc0 = y0 >> 1; rc = (rows + 1) >> 1;                    // chroma rows
    s0 = y0 >> 2; sr = (rows + 3) >> 2;                     // SCU rows
    (void)p;
}
size_t xb200_band_bytes(xb200_pic *p, int rows)
{
    if (!p || rows <= 0) return 0;
    int yc0, rc, s0, sr;
    band_geometry(p, 0, rows, yc0, rc, s0, sr);
    const size_t b = (size_t)rows * p->w * 2 + 2 * (size_t)rc * p->w_c * 2 + (size_t)sr * p->w_scu * (8 + 8 + 4 + 2 + 1);
    return (b + 255) & ~(size_t)255;
}
static int band_copy(xb200_ctx *c, xb200_pic *p, int y0, int rows, unsigned char *buf, bool pack)
{
    if (!c || !p || !buf || y0 < 0 || rows <= 0 || y0 + rows > p->h || (y0 & 3)) return XB200_ERR_INVALID_ARGUMENT;
    cudaSetDevice(c->device);
    int yc0, rc, s0, sr;
    band_geometry(p, y0, rows, yc0, rc, s0, sr);
    if (yc0 + rc > p->h_c) rc = p->h_c - yc0;
    if (s0 + sr > p->h_scu) sr = p->h_scu - s0;
    size_t off = 0;
    auto plane = [&](pel *base, int stride, int w, int r0, int nr) -> cudaError_t {
        pel *pp = base + (size_t)r0 * stride;
        cudaError_t e = pack ? cudaMemcpy2DAsync(buf + off, (size_t)w * 2, pp, (size_t)stride * 2, (size_t)w * 2, nr, cudaMemcpyDeviceToDevice, c->stream)
                             : cudaMemcpy2DAsync(pp, (size_t)stride * 2, buf + off, (size_t)w * 2, (size_t)w * 2, nr, cudaMemcpyDeviceToDevice, c->stream);
        off += (size_t)w * 2 * nr;
        return e;
    };
    auto map = [&](void *base, int bytes_per_scu) -> cudaError_t {
        unsigned char *pp = (unsigned char *)base + (size_t)s0 * p->w_scu * bytes_per_scu;
        const size_t n = (size_t)sr * p->w_scu * bytes_per_scu;
        cudaError_t e = pack ? cudaMemcpyAsync(buf + off, pp, n, cudaMemcpyDeviceToDevice, c->stream) : cudaMemcpyAsync(pp, buf + off, n, cudaMemcpyDeviceToDevice, c->stream);
        off += n;
        return e;
    };
    CK(c, plane(p->y, p->s_l, p->w, y0, rows));
    CK(c, plane(p->u, p->s_c, p->w_c, yc0, rc));
    CK(c, plane(p->v, p->s_c, p->w_c, yc0, rc));
    CK(c, map(p->map_mv, 8)); CK(c, map(p->map_unrefined_mv, 8)); CK(c, map(p->map_scu, 4)); CK(c, map(p->map_refi, 2)); CK(c, map(p->map_edge, 1));
    return XB200_OK;
}
int xb200_band_pack(xb200_ctx *c, xb200_pic *p, int y0, int rows, void *d_dst) { return band_copy(c, p, y0, rows, (unsigned char *)d_dst, true); }
int xb200_band_unpack(xb200_ctx *c, xb200_pic *p, int y0, int rows, const void *d_src) { return band_copy(c, p, y0, rows, (unsigned char *)d_src, false); }

// ---- batched leaf kernels ---------------------------------------------------------------------------------------------
int xb200_itdq_blocks_dev(xb200_ctx *c, const void *d_in, void *d_out, int n, int log2w, int log2h, int qp, int bit_depth, int iqt)
{
    if (!c || !d_in || !d_out || n <= 0 || log2w < 1 || log2w > 6 || log2h < 1 || log2h > 6) return XB200_ERR_INVALID_ARGUMENT;
    if (qp < 0 || qp > 81 || bit_depth < 8 || bit_depth > 14) return XB200_ERR_INVALID_ARGUMENT;
    if (((uintptr_t)d_in | (uintptr_t)d_out) & 15) return XB200_ERR_INVALID_ARGUMENT;       // 16-byte vector access
    cudaSetDevice(c->device);
    int r = xb::launch_itdq_blocks((const int16_t *)d_in, (int16_t *)d_out, n, log2w, log2h, qp, bit_depth, iqt, c->stream);
    if (r < 0) return r;
    c->launches += r;
    CK(c, cudaGetLastError());
    return XB200_OK;
}

int xb200_mc_blocks_dev(xb200_ctx *c, xb200_pic *ref, int plane, const void *d_mv, void *d_out, int n, int w, int h, int bit_depth, int main_tables)
{
    if (!c || !ref || !d_mv || !d_out || n <= 0 || plane < 0 || plane > 2) return XB200_ERR_INVALID_ARGUMENT;
    cudaSetDevice(c->device);
    if (bit_depth < 8 || bit_depth > 14 || ((uintptr_t)d_out & 3)) return XB200_ERR_INVALID_ARGUMENT;
    int r = xb::launch_mc_blocks(ref->d_tmaps + (plane == 0 ? 0 : (plane == 1 ? 3 : 2)), plane ? ref->pad_c : ref->pad_l, plane, (const int *)d_mv, (pel *)d_out, n, w, h,
                                 bit_depth, main_tables != 0, c->stream);
    if (r < 0) return r;
    c->launches += r;
    CK(c, cudaGetLastError());
    return XB200_OK;
}

}  // extern "C"
