"""Host-side mirror of the reference's picture-level seam, bound to the CUDA library through the C ABI.

    reference                                         here
    ------------------------------------------------------------------------------------------
    xevd_platform_init (table wiring)                 Context()
    PICBUF_ALLOCATOR.fn_alloc / fn_free               Context.pic_alloc / DevicePicture.free
    xevd_ctu_row_rec_mt (recon of a whole picture)    Context.recon_frame
    ctx->fn_deblock                                   Context.deblock
    ctx->fn_picbuf_expand                             Context.pad
    xevd_pull (host-readable XEVD_IMGB planes)        DevicePicture.download

Every method calls libxevd_b200.so; there is no Python or CPU implementation behind it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .frame import CuList, HostPicture


class XevdB200Error(RuntimeError):
    def __init__(self, code: int, what: str, detail: str = ""):
        super().__init__(f"{what} failed with {code}" + (f" ({detail})" if detail else ""))
        self.code = code


class DevicePicture:
    def __init__(self, ctx: "Context", handle: int, w: int, h: int):
        self.ctx, self.handle, self.w, self.h = ctx, handle, w, h
        self.poc = 0

    def info(self) -> abi.PicInfo:
        i = abi.PicInfo()
        self.ctx._chk(self.ctx.lib.xb200_pic_info(self.handle, C.byref(i)), "xb200_pic_info")
        return i

    def set_poc(self, poc: int):
        self.poc = poc
        self.ctx.lib.xb200_pic_set_poc(self.handle, poc)
        return self

    def upload(self, pic: HostPicture, padded: bool = True):
        """copy a host picture to the device (interior samples), then replicate borders on the device"""
        L = self.ctx.lib
        self.ctx._chk(L.xb200_pic_upload(self.ctx.handle, self.handle,
                                         pic.y.ctypes.data, pic.s_l, pic.u.ctypes.data, pic.s_c, pic.v.ctypes.data, pic.s_c),
                      "xb200_pic_upload")
        self.set_poc(pic.poc)
        if padded:
            self.ctx.pad(self)
        self.ctx.sync()
        return self

    def download(self, out: HostPicture | None = None, maps: bool = False) -> HostPicture:
        out = out or HostPicture(self.w, self.h, self.poc)
        L = self.ctx.lib
        self.ctx._chk(L.xb200_pic_download(self.ctx.handle, self.handle,
                                           out.y.ctypes.data, out.s_l, out.u.ctypes.data, out.s_c, out.v.ctypes.data, out.s_c),
                      "xb200_pic_download")
        self.ctx.sync()
        if maps:
            self.ctx._chk(L.xb200_pic_download_maps(self.ctx.handle, self.handle, out.map_mv.ctypes.data,
                                                    out.map_refi.ctypes.data, out.map_scu.ctypes.data), "xb200_pic_download_maps")
            self.ctx._chk(L.xb200_pic_download_unrefined_mv(self.ctx.handle, self.handle, out.map_unrefined_mv.ctypes.data),
                          "xb200_pic_download_unrefined_mv")
        return out

    def pull(self, dra=None, out_bits: int = 16, crop=(0, 0, 0, 0)):
        """what xevd_pull hands out: optional DRA, crop (left, right, top, bottom in luma samples), 16- or 8-bit planes"""
        cl, cr, ct, cb = crop
        w, h = self.w - cl - cr, self.h - ct - cb
        dt = np.uint8 if out_bits == 8 else np.int16
        y, u, v = np.zeros((h, w), dt), np.zeros((h // 2, w // 2), dt), np.zeros((h // 2, w // 2), dt)
        self.ctx._chk(self.ctx.lib.xb200_pic_pull(self.ctx.handle, self.handle, C.byref(dra) if dra is not None else None, out_bits, cl, cr, ct, cb,
                                                  y.ctypes.data, w, u.ctypes.data, w // 2, v.ctypes.data, w // 2), "xb200_pic_pull")
        self.ctx.sync()
        return y, u, v

    def download_edge_map(self) -> np.ndarray:
        out = np.zeros(((self.w + 3) >> 2) * ((self.h + 3) >> 2), np.uint8)
        self.ctx._chk(self.ctx.lib.xb200_pic_download_edge_map(self.ctx.handle, self.handle, out.ctypes.data), "xb200_pic_download_edge_map")
        return out

    def upload_maps(self, pic: HostPicture, edge: np.ndarray | None = None):
        e = np.ascontiguousarray(edge, np.uint8) if edge is not None else None
        self.ctx._chk(self.ctx.lib.xb200_pic_upload_maps(self.ctx.handle, self.handle, pic.map_mv.ctypes.data, pic.map_refi.ctypes.data,
                                                         pic.map_scu.ctypes.data, e.ctypes.data if e is not None else None),
                      "xb200_pic_upload_maps")
        return self

    def download_padded(self) -> HostPicture:
        out = HostPicture(self.w, self.h, self.poc)
        self.ctx._chk(self.ctx.lib.xb200_pic_download_padded(self.ctx.handle, self.handle, out.buf_y.ctypes.data,
                                                             out.buf_u.ctypes.data, out.buf_v.ctypes.data),
                      "xb200_pic_download_padded")
        return out

    def free(self):
        if self.handle:
            self.ctx.lib.xb200_pic_free(self.ctx.handle, self.handle)
            self.handle = None


def _handles(pics):
    arr = (C.c_void_p * max(1, len(pics)))()
    for i, p in enumerate(pics):
        arr[i] = p.handle
    return arr


class Context:
    """one decoder instance on one GPU (one CUDA stream)"""

    def __init__(self, device: int = 0, lib_path=None):
        self.lib = abi.load_library(lib_path)          # lib_path: another build of the library (A/B timing, tools/ab_v2.py)
        err = C.c_int(0)
        self.handle = self.lib.xb200_create(device, C.byref(err))
        if not self.handle:
            raise XevdB200Error(err.value, "xb200_create", "no usable CUDA device" if err.value == abi.XB200_ERR_NO_DEVICE else "")
        self.device = device

    def _chk(self, code: int, what: str):
        if code < 0:
            raise XevdB200Error(code, what, (self.lib.xb200_last_error(self.handle) or b"").decode())
        return code

    def close(self):
        if self.handle:
            self.lib.xb200_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        self._chk(self.lib.xb200_sync(self.handle), "xb200_sync")

    @property
    def stream(self) -> int:
        return self.lib.xb200_stream(self.handle) or 0

    def set_stream(self, cuda_stream: int):
        self._chk(self.lib.xb200_set_stream(self.handle, cuda_stream), "xb200_set_stream")

    @property
    def launches(self) -> int:
        return int(self.lib.xb200_launch_count(self.handle))

    def pic_alloc(self, w: int, h: int) -> DevicePicture:
        err = C.c_int(0)
        hnd = self.lib.xb200_pic_alloc(self.handle, w, h, C.byref(err))
        if not hnd:
            raise XevdB200Error(err.value, "xb200_pic_alloc", (self.lib.xb200_last_error(self.handle) or b"").decode())
        return DevicePicture(self, hnd, w, h)

    def set_tiles(self, col_bd=None, row_bd=None, across: bool = False):
        """tile grid for the loop filters: column / row boundaries in CTUs (None: one tile)"""
        cb = np.ascontiguousarray(col_bd if col_bd is not None else [0, 0xffff], np.uint16)
        rb = np.ascontiguousarray(row_bd if row_bd is not None else [0, 0xffff], np.uint16)
        self._chk(self.lib.xb200_set_tiles(self.handle, len(cb) - 1, cb.ctypes.data, len(rb) - 1, rb.ctypes.data, int(across)), "xb200_set_tiles")

    def set_chroma_qp_table(self, tbl: np.ndarray):
        t = np.ascontiguousarray(tbl, np.int32)
        assert t.shape == (2, 58)
        self._chk(self.lib.xb200_set_chroma_qp_table(self.handle, t.ctypes.data), "xb200_set_chroma_qp_table")

    def deblock(self, prm: abi.Params, cur: DevicePicture, refs_l0=(), refs_l1=(), edge_flags: np.ndarray | None = None):
        e = np.ascontiguousarray(edge_flags, np.uint8) if edge_flags is not None else None
        self._chk(self.lib.xb200_deblock(self.handle, C.byref(prm), cur.handle, _handles(refs_l0), len(refs_l0), _handles(refs_l1),
                                         len(refs_l1), e.ctypes.data if e is not None else None), "xb200_deblock")

    def alf(self, prm: abi.Params, pic: DevicePicture, alf: abi.AlfParams, ctb_flag_luma=None):
        """mctx->fn_alf (xevd_alf, src_main/xevdm.c:2105): adaptive loop filter in place"""
        f = np.ascontiguousarray(ctb_flag_luma, np.uint8) if ctb_flag_luma is not None else None
        self._chk(self.lib.xb200_alf(self.handle, C.byref(prm), pic.handle, C.byref(alf), f.ctypes.data if f is not None else None), "xb200_alf")

    # -- band exchange (intra-picture multi-GPU sharding) ------------------------------------------------
    def band_bytes(self, pic: DevicePicture, rows: int) -> int:
        return int(self.lib.xb200_band_bytes(pic.handle, rows))

    def band_pack(self, pic: DevicePicture, y0: int, rows: int, d_dst: int):
        self._chk(self.lib.xb200_band_pack(self.handle, pic.handle, y0, rows, d_dst), "xb200_band_pack")

    def band_unpack(self, pic: DevicePicture, y0: int, rows: int, d_src: int):
        self._chk(self.lib.xb200_band_unpack(self.handle, pic.handle, y0, rows, d_src), "xb200_band_unpack")

    def pad(self, pic: DevicePicture):
        self._chk(self.lib.xb200_pad(self.handle, pic.handle), "xb200_pad")

    # -- host-buffer entry point (the drop-in call; copies inputs to the device itself) ----------------
    def recon_frame(self, prm: abi.Params, cur: DevicePicture, refs_l0, refs_l1, cl: CuList):
        cus = np.ascontiguousarray(cl.cus)
        first = np.ascontiguousarray(cl.ctu_first)
        ext = np.ascontiguousarray(cl.ext)
        coef = np.ascontiguousarray(cl.coef)
        cur.set_poc(prm.poc)
        self._chk(self.lib.xb200_recon_frame(self.handle, C.byref(prm), cur.handle,
                                             _handles(refs_l0), len(refs_l0), _handles(refs_l1), len(refs_l1),
                                             cus.ctypes.data, len(cus), first.ctypes.data, len(first) - 1,
                                             ext.ctypes.data, len(ext), coef.ctypes.data, coef.size), "xb200_recon_frame")

    def recon_frame_sparse(self, prm: abi.Params, cur: DevicePicture, refs_l0, refs_l1, cl: CuList, sparse=None):
        """xb200_recon_frame with the coefficient stream in sparse form (entries, chunk_first) - see frame.sparse_coef"""
        from .frame import sparse_coef
        entries, chunk_first = sparse if sparse is not None else sparse_coef(cl.coef)
        cus = np.ascontiguousarray(cl.cus)
        first = np.ascontiguousarray(cl.ctu_first)
        ext = np.ascontiguousarray(cl.ext)
        cur.set_poc(prm.poc)
        self._chk(self.lib.xb200_recon_frame_sparse(self.handle, C.byref(prm), cur.handle,
                                                    _handles(refs_l0), len(refs_l0), _handles(refs_l1), len(refs_l1),
                                                    cus.ctypes.data, len(cus), first.ctypes.data, len(first) - 1,
                                                    ext.ctypes.data, len(ext), entries.ctypes.data, entries.size, chunk_first.ctypes.data, cl.coef.size),
                  "xb200_recon_frame_sparse")

    # -- device-resident entry point (inputs already in HBM; pointers are raw device addresses) --------------
    def recon_frame_dev(self, prm: abi.Params, cur: DevicePicture, refs_l0, refs_l1, d_cus: int, n_cu: int,
                        d_first: int, n_ctu: int, d_ext: int, n_ext: int, d_coef: int, n_coef: int, has_intra: int = 0,
                        max_cu_per_ctu: int = 0):
        cur.set_poc(prm.poc)
        self._chk(self.lib.xb200_recon_frame_dev(self.handle, C.byref(prm), cur.handle,
                                                 _handles(refs_l0), len(refs_l0), _handles(refs_l1), len(refs_l1),
                                                 d_cus, n_cu, d_first, n_ctu, d_ext, n_ext, d_coef, n_coef, int(has_intra),
                                                 int(max_cu_per_ctu)),
                  "xb200_recon_frame_dev")

    def itdq_blocks_dev(self, d_in: int, d_out: int, n: int, log2w: int, log2h: int, qp: int, bit_depth: int, iqt: bool = False):
        self._chk(self.lib.xb200_itdq_blocks_dev(self.handle, d_in, d_out, n, log2w, log2h, qp, bit_depth, int(iqt)),
                  "xb200_itdq_blocks_dev")

    def mc_blocks_dev(self, ref: DevicePicture, plane: int, d_mv: int, d_out: int, n: int, w: int, h: int, bit_depth: int,
                      main_tables: bool = False):
        self._chk(self.lib.xb200_mc_blocks_dev(self.handle, ref.handle, plane, d_mv, d_out, n, w, h, bit_depth, int(main_tables)),
                  "xb200_mc_blocks_dev")
