"""CPU-side checks of the boundary: the shared library loads, exports every symbol that include/xevd_b200.h
declares, struct layouts match, and (without a GPU) compute entry points fail loudly instead of falling back."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from xevd_b200 import abi

ROOT = Path(__file__).resolve().parents[1]


def _declared_symbols():
    text = (ROOT / "include" / "xevd_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/xevd_b200.h but not exported"
    assert set(names) == set(abi.EXPORTED_SYMBOLS), set(names) ^ set(abi.EXPORTED_SYMBOLS)
    assert lib.xb200_abi_version() == 1


def test_struct_layouts_match_header():
    src = r'''
    #include "xevd_b200.h"
    #include <stdio.h>
    #include <stddef.h>
    int main(void){
      printf("%zu %zu %zu %zu ", sizeof(XB200_CU), sizeof(XB200_CU_EXT), sizeof(XB200_PARAMS), sizeof(XB200_PIC_INFO));
      printf("%zu %zu %zu %zu %zu\n", offsetof(XB200_CU, refi), offsetof(XB200_CU, cbf), offsetof(XB200_CU, mv), offsetof(XB200_CU, ats), offsetof(XB200_CU, coef_off));
      return 0; }
    '''
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "t.c").write_text(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), "-o", f"{d}/t", f"{d}/t.c"], check=True)
        out = subprocess.run([f"{d}/t"], capture_output=True, text=True, check=True).stdout.split()
    vals = list(map(int, out))
    assert vals[:4] == [abi.CU_DTYPE.itemsize, abi.EXT_DTYPE.itemsize, C.sizeof(abi.Params), C.sizeof(abi.PicInfo)]
    f = abi.CU_DTYPE.fields
    assert vals[4:] == [f["refi"][1], f["cbf"][1], f["mv"][1], f["ats"][1], f["coef_off"][1]]


def test_constants_match_header():
    """the Python mirror of the header's #defines (modes, CU flags, edge flags, has_intra bits, error codes)"""
    text = (ROOT / "include" / "xevd_b200.h").read_text()
    d = {k: int(v, 0) for k, v in re.findall(r"#define\s+(XB200_[A-Z0-9_]+)\s+\(?(-?(?:0x[0-9a-fA-F]+|\d+))\)?", text)}
    want = {"XB200_MODE_INTRA": abi.MODE_INTRA, "XB200_MODE_INTER": abi.MODE_INTER, "XB200_MODE_IBC": abi.MODE_IBC, "XB200_MODE_AFFINE": abi.MODE_AFFINE,
            "XB200_CUF_LUMA": abi.CUF_LUMA, "XB200_CUF_CHROMA": abi.CUF_CHROMA, "XB200_CUF_SKIP": abi.CUF_SKIP, "XB200_CUF_DMVR": abi.CUF_DMVR,
            "XB200_CUF_ATS_INTRA": abi.CUF_ATS_INTRA, "XB200_CUF_AFF6": abi.CUF_AFF6,
            "XB200_EDGE_LEFT": abi.EDGE_LEFT, "XB200_EDGE_TOP": abi.EDGE_TOP, "XB200_EDGE_ATS": abi.EDGE_ATS,
            "XB200_EDGE_LEFT_NOC": abi.EDGE_LEFT_NOC, "XB200_EDGE_TOP_NOC": abi.EDGE_TOP_NOC,
            "XB200_HAS_INTRA": abi.HAS_INTRA, "XB200_HAS_DUAL_TREE": abi.HAS_DUAL_TREE,
            "XB200_HAS_DENSE_WAVEFRONT": abi.HAS_DENSE_WAVEFRONT}
    for k, v in want.items():
        assert d.get(k) == v, (k, d.get(k), v)


def test_no_cpu_fallback_without_device():
    lib = abi.load_library()
    if lib.xb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    err = C.c_int(0)
    h = lib.xb200_create(0, C.byref(err))
    assert not h and err.value == abi.XB200_ERR_NO_DEVICE
    from xevd_b200.device import Context, XevdB200Error
    with pytest.raises(XevdB200Error):
        Context(0)


def test_product_package_never_imports_oracle():
    for p in (ROOT / "xevd_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".c", ".h", ".cpp"):
            t = p.read_text()
            assert "oracle" not in t.lower().replace("oracle/", "oracle/") or "import oracle" not in t and "from oracle" not in t and "orc_" not in t, p
