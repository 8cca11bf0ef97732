"""ctypes binding of the reference's public decoder API (inc/xevd.h:369-374) for ANY library that exports it:

    glue/_build/libxevd_gpu.so     the drop-in: the reference decoder with its reconstruction half on the GPU (glue/xevd_b200_glue.c)
    oracle/_ref/libxevd_ref.so     the unmodified reference (CPU)

Same six calls, same structs (XEVD_CDSC, XEVD_BITB, XEVD_STAT, XEVD_IMGB), same error convention: >= 0 success, < 0 failure,
xevd_pull returning XEVD_ERR_UNEXPECTED when the DPB has nothing to hand out.  Nothing here computes pixels.
An elementary stream is a list of NAL units; on disk every NAL unit carries a 4-byte big-endian length prefix, the container
app/xevd_app.c reads (:52-107)."""
from __future__ import annotations

import ctypes as C
import struct
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
GPU_SO = ROOT / "glue" / "_build" / "libxevd_gpu.so"
REF_SO = ROOT / "oracle" / "_ref" / "libxevd_ref.so"

XEVD_OK = 0
XEVD_OK_FRM_DELAYED = 202
XEVD_ERR_UNEXPECTED = -105
NUT_NONIDR, NUT_IDR, NUT_SPS, NUT_PPS, NUT_APS, NUT_SEI = 0, 1, 24, 25, 26, 28
ST_B, ST_P, ST_I = 0, 1, 2


class Cdsc(C.Structure):
    _fields_ = [("threads", C.c_int)]


class Bitb(C.Structure):
    _fields_ = [("addr", C.c_void_p), ("pddr", C.c_void_p), ("bsize", C.c_int), ("ssize", C.c_int), ("err", C.c_int),
                ("ndata", C.c_int * 4), ("pdata", C.c_void_p * 4), ("ts", C.c_longlong * 2)]


class Stat(C.Structure):
    _fields_ = [("read", C.c_int), ("nalu_type", C.c_int), ("stype", C.c_int), ("fnum", C.c_int), ("poc", C.c_int), ("tid", C.c_int),
                ("refpic_num", C.c_ubyte * 2), ("refpic", (C.c_int * 16) * 2)]


class Imgb(C.Structure):
    pass


_IMGB_FN = C.CFUNCTYPE(C.c_int, C.POINTER(Imgb))
Imgb._fields_ = [("cs", C.c_int), ("np", C.c_int), ("w", C.c_int * 4), ("h", C.c_int * 4), ("x", C.c_int * 4), ("y", C.c_int * 4),
                 ("s", C.c_int * 4), ("e", C.c_int * 4), ("a", C.c_void_p * 4), ("ts", C.c_longlong * 2), ("ndata", C.c_int * 4),
                 ("pdata", C.c_void_p * 4), ("aw", C.c_int * 4), ("ah", C.c_int * 4), ("padl", C.c_int * 4), ("padr", C.c_int * 4),
                 ("padu", C.c_int * 4), ("padb", C.c_int * 4), ("baddr", C.c_void_p * 4), ("bsize", C.c_int * 4), ("refcnt", C.c_int),
                 ("addref", _IMGB_FN), ("getref", _IMGB_FN), ("release", _IMGB_FN), ("crop_idx", C.c_int), ("crop_l", C.c_int),
                 ("crop_r", C.c_int), ("crop_t", C.c_int), ("crop_b", C.c_int), ("imgb_active_pps_id", C.c_int), ("imgb_active_aps_id", C.c_int)]


class XevdLibrary:
    """one libxevd-compatible shared library"""

    def __init__(self, path):
        path = Path(path)
        if not path.exists():
            raise RuntimeError(f"{path} is missing (build it: `make -C glue` for the GPU drop-in, `make -C oracle` for the reference)")
        self.path = path
        self.lib = L = C.CDLL(str(path))
        L.xevd_create.restype = C.c_void_p
        L.xevd_create.argtypes = [C.POINTER(Cdsc), C.POINTER(C.c_int)]
        L.xevd_delete.restype = None
        L.xevd_delete.argtypes = [C.c_void_p]
        L.xevd_decode.restype = C.c_int
        L.xevd_decode.argtypes = [C.c_void_p, C.POINTER(Bitb), C.POINTER(Stat)]
        L.xevd_pull.restype = C.c_int
        L.xevd_pull.argtypes = [C.c_void_p, C.POINTER(C.POINTER(Imgb))]
        L.xevd_config.restype = C.c_int
        L.xevd_config.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]


class Decoder:
    """xevd_create ... xevd_delete around one stream, the way app/xevd_app.c drives the library (:449-627)"""

    def __init__(self, library: XevdLibrary, threads: int = 1):
        self.L = library.lib
        err = C.c_int(0)
        cdsc = Cdsc(threads)
        self.id = self.L.xevd_create(C.byref(cdsc), C.byref(err))
        if not self.id:
            raise RuntimeError(f"xevd_create failed: {err.value} ({library.path.name})")

    def close(self):
        if self.id:
            self.L.xevd_delete(self.id)
            self.id = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def decode(self, nal: bytes):
        """one NAL unit (no length prefix).  Returns (return code, XEVD_STAT)."""
        buf = (C.c_ubyte * len(nal)).from_buffer_copy(nal)
        bitb = Bitb()
        bitb.addr = C.addressof(buf)
        bitb.bsize = bitb.ssize = len(nal)
        stat = Stat()
        ret = self.L.xevd_decode(self.id, C.byref(bitb), C.byref(stat))
        return ret, stat

    def pull(self):
        """next picture in output order as (y, u, v) int16 arrays with the cropping window applied, or None"""
        p = C.POINTER(Imgb)()
        ret = self.L.xevd_pull(self.id, C.byref(p))
        if ret < 0 or not p:
            return None
        im = p.contents
        planes = []
        for i in range(3):
            w, h, s = im.w[i], im.h[i], im.s[i]
            base = im.a[i]
            arr = np.ctypeslib.as_array(C.cast(base, C.POINTER(C.c_int16)), (im.y[i] + h, s // 2))
            planes.append(arr[im.y[i]:im.y[i] + h, im.x[i]:im.x[i] + w].copy())
        im.release(p)
        return tuple(planes)


def decode_stream(library: XevdLibrary, nals, threads: int = 1):
    """decode a list of NAL units and return the pictures in output order (bumping at the end as app/xevd_app.c does, :524-531)"""
    pics = []
    with Decoder(library, threads) as d:
        for n in nals:
            ret, stat = d.decode(n)
            if ret < 0:
                raise RuntimeError(f"xevd_decode failed: {ret} ({library.path.name}, NAL type {n[0] >> 1 & 63})")
            if stat.fnum >= 0 or ret == XEVD_OK_FRM_DELAYED:
                while True:
                    p = d.pull()
                    if p is None:
                        break
                    pics.append(p)
        while True:
            p = d.pull()
            if p is None:
                break
            pics.append(p)
    return pics


def write_stream(path, nals):
    with open(path, "wb") as f:
        for n in nals:
            f.write(struct.pack(">I", len(n)))
            f.write(n)


def read_stream(path):
    data = Path(path).read_bytes()
    nals, o = [], 0
    while o + 4 <= len(data):
        (n,) = struct.unpack_from(">I", data, o)
        nals.append(data[o + 4:o + 4 + n])
        o += 4 + n
    return nals
