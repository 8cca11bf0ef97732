"""ctypes / numpy mirror of include/xevd_b200.h (the C ABI of the CUDA path).

Nothing here computes pixels: it only describes memory layouts and loads the shared library.
The library is required -- there is no CPU fallback (load_library raises if it is missing).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
LIB_PATH = ROOT / "libxevd_b200.so"

XB200_OK = 0
XB200_ERR_INVALID_ARGUMENT = -101
XB200_ERR_UNSUPPORTED = -104
XB200_ERR_NO_DEVICE = -401
XB200_ERR_CUDA = -402

MODE_INTRA, MODE_INTER, MODE_IBC, MODE_AFFINE = 0, 1, 4, 5
CUF_LUMA, CUF_CHROMA, CUF_SKIP, CUF_DMVR, CUF_ATS_INTRA, CUF_AFF6 = 1, 2, 4, 8, 16, 32
EDGE_LEFT, EDGE_TOP, EDGE_ATS, EDGE_LEFT_NOC, EDGE_TOP_NOC = 1, 2, 4, 8, 16
SPARSE_CHUNK = 4096                     # XB200_SPARSE_CHUNK
HAS_INTRA, HAS_DUAL_TREE, HAS_DENSE_WAVEFRONT = 1, 2, 4          # has_intra bits of xb200_recon_frame_dev

# struct XB200_CU (32 bytes)
CU_DTYPE = np.dtype(
    [
        ("x", "<u2"), ("y", "<u2"),
        ("log2w", "u1"), ("log2h", "u1"), ("mode", "u1"), ("flags", "u1"),
        ("qp_y", "u1"), ("qp_u", "u1"), ("qp_v", "u1"), ("qp_map", "u1"),
        ("refi", "i1", (2,)),
        ("cbf", "<u2"),
        ("mv", "<i2", (2, 2)),
        ("ats", "u1"), ("avail", "u1"), ("avail_cu", "<u2"),
        ("coef_off", "<u4"),
    ],
    align=False,
)
assert CU_DTYPE.itemsize == 32

# struct XB200_CU_EXT (32 bytes), viewed as 4 x u64 (intra) or 16 x i16 (affine)
EXT_DTYPE = np.dtype([("q", "<u8", (4,))])
assert EXT_DTYPE.itemsize == 32


class Params(C.Structure):
    """struct XB200_PARAMS"""

    _fields_ = [
        ("w", C.c_int32), ("h", C.c_int32),
        ("bit_depth_luma", C.c_int32), ("bit_depth_chroma", C.c_int32),
        ("chroma_format_idc", C.c_int32), ("log2_ctu", C.c_int32),
        ("tool_admvp", C.c_int32), ("tool_iqt", C.c_int32), ("tool_ats", C.c_int32),
        ("tool_addb", C.c_int32), ("tool_alf", C.c_int32), ("tool_htdf", C.c_int32),
        ("tool_dmvr", C.c_int32), ("tool_eipd", C.c_int32), ("tool_affine", C.c_int32),
        ("tool_ibc", C.c_int32),
        ("slice_qp", C.c_int32), ("qp_u_offset", C.c_int32), ("qp_v_offset", C.c_int32),
        ("deblock_alpha_offset", C.c_int32), ("deblock_beta_offset", C.c_int32),
        ("poc", C.c_int32),
        ("ctu_row0", C.c_int32), ("ctu_rows", C.c_int32),
        ("constrained_intra_pred", C.c_int32),
        ("tool_suco", C.c_int32),
        ("reserved", C.c_int32 * 6),
    ]


assert C.sizeof(Params) == 128


class DraParams(C.Structure):
    """struct XB200_DRA"""

    _fields_ = [("luma_inv_scale_lut", C.c_int32 * 1024), ("chroma_inv_scale_lut", (C.c_int32 * 1024) * 2)]


class AlfParams(C.Structure):
    """struct XB200_ALF"""

    _fields_ = [("coef_luma", (C.c_int16 * 13) * 25), ("coef_chroma", C.c_int16 * 7), ("enable", C.c_uint8 * 3), ("reserved", C.c_uint8)]


class PicInfo(C.Structure):
    """struct XB200_PIC_INFO"""

    _fields_ = [
        ("w_l", C.c_int32), ("h_l", C.c_int32), ("w_c", C.c_int32), ("h_c", C.c_int32),
        ("s_l", C.c_int32), ("s_c", C.c_int32), ("pad_l", C.c_int32), ("pad_c", C.c_int32),
        ("dev_y", C.c_void_p), ("dev_u", C.c_void_p), ("dev_v", C.c_void_p),
        ("dev_map_mv", C.c_void_p), ("dev_map_refi", C.c_void_p), ("dev_map_scu", C.c_void_p),
        ("w_scu", C.c_int32), ("h_scu", C.c_int32), ("poc", C.c_int32),
        ("dev_map_edge", C.c_void_p), ("dev_map_unrefined_mv", C.c_void_p),
    ]


def make_params(w, h, bit_depth=10, log2_ctu=6, poc=0, **tools) -> Params:
    p = Params()
    p.w, p.h = w, h
    p.bit_depth_luma = p.bit_depth_chroma = bit_depth
    p.chroma_format_idc = 1
    p.log2_ctu = log2_ctu
    p.poc = poc
    for k, v in tools.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


_SIGS = {
    # name: (restype, argtypes)
    "xb200_abi_version": (C.c_int, []),
    "xb200_device_count": (C.c_int, []),
    "xb200_create": (C.c_void_p, [C.c_int, C.POINTER(C.c_int)]),
    "xb200_destroy": (None, [C.c_void_p]),
    "xb200_last_error": (C.c_char_p, [C.c_void_p]),
    "xb200_sync": (C.c_int, [C.c_void_p]),
    "xb200_stream": (C.c_void_p, [C.c_void_p]),
    "xb200_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "xb200_launch_count": (C.c_longlong, [C.c_void_p]),
    "xb200_host_alloc": (C.c_void_p, [C.c_size_t]),
    "xb200_host_free": (None, [C.c_void_p]),
    "xb200_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "xb200_host_unregister": (C.c_int, [C.c_void_p]),
    "xb200_pic_alloc": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "xb200_pic_free": (None, [C.c_void_p, C.c_void_p]),
    "xb200_pic_info": (C.c_int, [C.c_void_p, C.POINTER(PicInfo)]),
    "xb200_pic_set_poc": (C.c_int, [C.c_void_p, C.c_int]),
    "xb200_pic_upload": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_void_p, C.c_int] * 3),
    "xb200_pic_download": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_void_p, C.c_int] * 3),
    "xb200_pic_download_padded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xb200_pic_download_maps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xb200_pic_download_edge_map": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "xb200_pic_pull": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "xb200_pic_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "xb200_pic_open_peer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "xb200_band_bytes": (C.c_size_t, [C.c_void_p, C.c_int]),
    "xb200_band_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "xb200_band_unpack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "xb200_pic_download_unrefined_mv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "xb200_recon_frame": (
        C.c_int,
        [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
         C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t],
    ),
    "xb200_recon_frame_sparse": (
        C.c_int,
        [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
         C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t],
    ),
    "xb200_recon_frame_dev": (
        C.c_int,
        [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
         C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int],
    ),
    "xb200_deblock": (C.c_int, [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "xb200_set_chroma_qp_table": (C.c_int, [C.c_void_p, C.c_void_p]),
    "xb200_set_tiles": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "xb200_pic_upload_maps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xb200_pad": (C.c_int, [C.c_void_p, C.c_void_p]),
    "xb200_alf": (C.c_int, [C.c_void_p, C.POINTER(Params), C.c_void_p, C.POINTER(AlfParams), C.c_void_p]),
    "xb200_itdq_blocks_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "xb200_mc_blocks_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    """Load libxevd_b200.so and attach prototypes.  Raises if the extension is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise RuntimeError(
            f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the CUDA extension is mandatory; there is no CPU fallback)"
        )
    lib = C.CDLL(str(p))
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # raises AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib
