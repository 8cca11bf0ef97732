"""ctypes front-end of the CPU oracle (oracle/liboracle.so) and of the compiled reference
(oracle/_ref/libxevd_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
The product package (xevd_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from xevd_b200.abi import CU_DTYPE, EXT_DTYPE, Params
from xevd_b200.frame import CuList, HostPicture

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "liboracle.so"
REF_SO = HERE / "_ref" / "libxevd_ref.so"


class OrcPic(C.Structure):
    """struct ORC_PIC (oracle/orc_common.h)"""

    _fields_ = [
        ("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p),
        ("s_l", C.c_int), ("s_c", C.c_int),
        ("w_l", C.c_int), ("h_l", C.c_int), ("w_c", C.c_int), ("h_c", C.c_int),
        ("pad_l", C.c_int), ("pad_c", C.c_int), ("poc", C.c_int),
        ("map_mv", C.c_void_p), ("map_refi", C.c_void_p), ("map_scu", C.c_void_p),
        ("w_scu", C.c_int), ("h_scu", C.c_int),
        ("map_unrefined_mv", C.c_void_p),
    ]


def orc_pic(p: HostPicture) -> OrcPic:
    o = OrcPic()
    o.y, o.u, o.v = p.y.ctypes.data, p.u.ctypes.data, p.v.ctypes.data
    o.s_l, o.s_c = p.s_l, p.s_c
    o.w_l, o.h_l, o.w_c, o.h_c = p.w, p.h, p.w_c, p.h_c
    o.pad_l, o.pad_c, o.poc = p.pad_l, p.pad_c, p.poc
    o.map_mv, o.map_refi, o.map_scu = p.map_mv.ctypes.data, p.map_refi.ctypes.data, p.map_scu.ctypes.data
    o.w_scu, o.h_scu = p.w_scu, p.h_scu
    o.map_unrefined_mv = p.map_unrefined_mv.ctypes.data
    return o


def build(force: bool = False) -> None:
    """compile the C restatement (and the reference harness when /root/reference is present).
    make is a no-op when the libraries are newer than their sources."""
    subprocess.run(["make", "-s", "-C", str(HERE), "-j8", "all"] + (["-B"] if force else []), check=True,
                   stdout=subprocess.DEVNULL)


_built = False


def _ensure_built():
    global _built
    if not _built:
        try:
            build()
        except (OSError, subprocess.CalledProcessError):
            if not ORACLE_SO.exists():
                raise
        _built = True


def _ptr_array(pics):
    keep = [orc_pic(p) for p in pics]
    arr = (C.POINTER(OrcPic) * max(1, len(keep)))()
    for i, k in enumerate(keep):
        arr[i] = C.pointer(k)
    return arr, keep


class _Backend:
    """common call surface of liboracle.so (prefix orc_) and libxevd_ref.so (prefix ref_)"""

    def __init__(self, path: Path, prefix: str):
        if not path.exists():
            raise FileNotFoundError(path)
        self.lib = C.CDLL(str(path))
        self.prefix = prefix
        L = self.lib
        for name in ("mc_luma", "mc_chroma"):
            f = getattr(L, prefix + name)
            f.restype = None
            f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        f = getattr(L, prefix + "itdq_block")
        f.restype = None
        f.argtypes = [C.c_void_p] + [C.c_int] * 5
        f = getattr(L, prefix + "recon_frame")
        f.restype = C.c_int
        f.argtypes = [C.POINTER(Params), C.POINTER(OrcPic), C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                      C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        f = getattr(L, prefix + "pad")
        f.restype = None
        f.argtypes = [C.POINTER(OrcPic)]

    # -- leaf kernels -----------------------------------------------------------------------
    def mc(self, plane: np.ndarray, origin_xy, gmv, ori_mv, w, h, bit_depth, chroma=False, main_tables=False):
        """plane: 2-D int16 array; origin_xy: array index of sample (0,0); gmv absolute (1/16 or 1/32 pel)"""
        assert plane.dtype == np.int16 and plane.flags.c_contiguous
        base = plane.ctypes.data + 2 * (origin_xy[1] * plane.shape[1] + origin_xy[0])
        # the reference's SIMD kernels store whole vectors: narrow blocks overrun w*h (harmless inside the
        # decoder, where pred[] is MAX_CU_DIM), so give the output slack
        buf = np.zeros(h * w + 256, np.int16)
        fn = getattr(self.lib, self.prefix + ("mc_chroma" if chroma else "mc_luma"))
        fn(base, plane.shape[1], gmv[0], gmv[1], ori_mv[0], ori_mv[1], buf.ctypes.data, w, w, h, bit_depth, int(main_tables))
        return buf[:h * w].reshape(h, w).copy()

    def itdq_block(self, coef: np.ndarray, qp: int, bit_depth: int, iqt=False):
        h, w = coef.shape
        c = np.ascontiguousarray(coef, np.int16).copy()
        getattr(self.lib, self.prefix + "itdq_block")(c.ctypes.data, int(np.log2(w)), int(np.log2(h)), qp, bit_depth, int(iqt))
        return c

    # -- picture level ------------------------------------------------------------------------
    def recon_frame(self, prm: Params, cur: HostPicture, refs_l0, refs_l1, cl: CuList):
        a0, k0 = _ptr_array(refs_l0)
        a1, k1 = _ptr_array(refs_l1)
        cur_o = orc_pic(cur)
        cus = np.ascontiguousarray(cl.cus)
        ext = np.ascontiguousarray(cl.ext)
        coef = np.ascontiguousarray(cl.coef)
        r = getattr(self.lib, self.prefix + "recon_frame")(
            C.byref(prm), C.byref(cur_o), a0, len(refs_l0), a1, len(refs_l1),
            cus.ctypes.data, len(cus), ext.ctypes.data, coef.ctypes.data)
        if r < 0:
            raise RuntimeError(f"{self.prefix}recon_frame failed: {r}")
        return cur

    def set_tiles(self, col_bd=None, row_bd=None, across: bool = False):
        """PPS tile grid for deblock_frame / alf_frame: column / row boundaries in CTUs (None: one tile)"""
        cb = np.ascontiguousarray(col_bd if col_bd is not None else [0, 0xffff], np.uint16)
        rb = np.ascontiguousarray(row_bd if row_bd is not None else [0, 0xffff], np.uint16)
        fn = getattr(self.lib, self.prefix + "set_tiles")
        fn.restype = None
        fn.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        fn(len(cb) - 1, cb.ctypes.data, len(rb) - 1, rb.ctypes.data, int(across))

    def deblock_frame(self, prm: Params, pic: HostPicture, cl: CuList, chroma_qp_tbl: np.ndarray, tool_addb: bool = False,
                      ref_ids=((0, 1, 2, 3), (0, 1, 2, 3))):
        """both deblocking passes in place on `pic` (uses pic.map_scu / map_mv / map_refi).  ref_ids[l][i] identifies the
        PICTURE behind reference index i of list l (the Main-profile filter compares pictures, not indices)."""
        o = orc_pic(pic)
        cus = np.ascontiguousarray(cl.cus)
        tbl = np.ascontiguousarray(chroma_qp_tbl, np.int32)
        assert tbl.shape == (2, 58)
        r0, r1 = np.array(ref_ids[0], np.int32), np.array(ref_ids[1], np.int32)
        if self.prefix == "ref_":
            fn = self.lib.ref_deblock_frame
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(Params), C.POINTER(OrcPic), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
            r = fn(C.byref(prm), C.byref(o), cus.ctypes.data, len(cus), tbl.ctypes.data, int(tool_addb), r0.ctypes.data, len(r0), r1.ctypes.data, len(r1))
        elif tool_addb:
            fn = self.lib.orc_deblock_frame_addb
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(Params), C.POINTER(OrcPic), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
            r = fn(C.byref(prm), C.byref(o), cus.ctypes.data, len(cus), tbl.ctypes.data, r0.ctypes.data, r1.ctypes.data)
        else:
            fn = self.lib.orc_deblock_frame
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(Params), C.POINTER(OrcPic), C.c_void_p, C.c_int, C.c_void_p]
            r = fn(C.byref(prm), C.byref(o), cus.ctypes.data, len(cus), tbl.ctypes.data)
        if r < 0:
            raise RuntimeError(f"{self.prefix}deblock_frame failed: {r}")
        return pic

    def alf_frame(self, prm: Params, pic: HostPicture, alf, ctb_flag_luma: np.ndarray | None = None):
        """adaptive loop filter in place; alf = xevd_b200.abi.AlfParams"""
        o = orc_pic(pic)
        fn = getattr(self.lib, self.prefix + "alf_frame")
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(Params), C.POINTER(OrcPic), C.c_void_p, C.c_void_p]
        f = np.ascontiguousarray(ctb_flag_luma, np.uint8) if ctb_flag_luma is not None else None
        r = fn(C.byref(prm), C.byref(o), C.addressof(alf), f.ctypes.data if f is not None else None)
        if r < 0:
            raise RuntimeError(f"{self.prefix}alf_frame failed: {r}")
        return pic

    def dra_apply(self, pic: HostPicture, dra):
        """dynamic range adjustment in place (what xevd_pull applies to a copy of the picture when tool_dra is on)"""
        o = orc_pic(pic)
        fn = getattr(self.lib, self.prefix + "dra_apply")
        fn.restype = None
        fn.argtypes = [C.POINTER(OrcPic), C.c_void_p]
        fn(C.byref(o), C.addressof(dra))
        return pic

    def pad(self, pic: HostPicture):
        o = orc_pic(pic)
        getattr(self.lib, self.prefix + "pad")(C.byref(o))
        return pic


class Oracle(_Backend):
    def __init__(self):
        _ensure_built()
        super().__init__(ORACLE_SO, "orc_")
        L = self.lib
        L.orc_dct2_matrix.restype = C.POINTER(C.c_int8)
        L.orc_dct2_matrix.argtypes = [C.c_int]
        L.orc_ats_matrix.restype = C.POINTER(C.c_int16)
        L.orc_ats_matrix.argtypes = [C.c_int, C.c_int]
        L.orc_mc_luma_taps.restype = C.POINTER(C.c_int16)
        L.orc_mc_luma_taps.argtypes = [C.c_int]
        L.orc_mc_chroma_taps.restype = C.POINTER(C.c_int16)
        L.orc_mc_chroma_taps.argtypes = [C.c_int]

    def dct2_matrix(self, log2n):
        n = 1 << log2n
        return np.ctypeslib.as_array(self.lib.orc_dct2_matrix(log2n), (n, n)).copy()

    def ats_matrix(self, dst7, log2n):
        n = 1 << log2n
        return np.ctypeslib.as_array(self.lib.orc_ats_matrix(int(dst7), log2n), (n, n)).copy()

    def mc_taps(self, main_tables):
        return (np.ctypeslib.as_array(self.lib.orc_mc_luma_taps(int(main_tables)), (16, 8)).copy(),
                np.ctypeslib.as_array(self.lib.orc_mc_chroma_taps(int(main_tables)), (32, 4)).copy())


def oracle_output(oracle, pic: HostPicture, out_bits=16, crop=(0, 0, 0, 0)):
    """orc_output: cropped 16- or 8-bit planes"""
    cl, cr, ct, cb = crop
    w, h = pic.w - cl - cr, pic.h - ct - cb
    dt = np.uint8 if out_bits == 8 else np.int16
    y, u, v = np.zeros((h, w), dt), np.zeros((h // 2, w // 2), dt), np.zeros((h // 2, w // 2), dt)
    o = orc_pic(pic)
    fn = oracle.lib.orc_output
    fn.restype = None
    fn.argtypes = [C.POINTER(OrcPic)] + [C.c_int] * 5 + [C.c_void_p, C.c_int] * 3
    fn(C.byref(o), out_bits, cl, cr, ct, cb, y.ctypes.data, w, u.ctypes.data, w // 2, v.ctypes.data, w // 2)
    return y, u, v


class Reference(_Backend):
    """the unmodified reference library; impl 0 = plain C, 1 = SSE4.1, 2 = AVX2 (dispatched on x86)"""

    def __init__(self, impl: int = 2):
        _ensure_built()
        super().__init__(REF_SO, "ref_")
        self.lib.ref_set_impl.argtypes = [C.c_int]
        self.lib.ref_set_impl(impl)
        self.impl = impl

    def set_impl(self, impl):
        self.lib.ref_set_impl(impl)
        self.impl = impl

    def dct2_matrix(self, log2n):
        n = 1 << log2n
        out = np.zeros((n, n), np.int8)
        assert self.lib.ref_get_dct2(log2n, C.c_void_p(out.ctypes.data)) == 0
        return out

    def ats_matrix(self, dst7, log2n):
        n = 1 << log2n
        out = np.zeros((n, n), np.int16)
        assert self.lib.ref_get_inv_ats(int(dst7), log2n, C.c_void_p(out.ctypes.data)) == 0
        return out

    def mc_taps(self, main_tables):
        l, c = np.zeros((16, 8), np.int16), np.zeros((32, 4), np.int16)
        self.lib.ref_get_mc_taps(int(main_tables), C.c_void_p(l.ctypes.data), C.c_void_p(c.ctypes.data))
        return l, c


def have_reference() -> bool:
    return REF_SO.exists()
