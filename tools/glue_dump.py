#!/usr/bin/env python3
"""Reads the work lists libxevd_gpu.so dumps with XEVD_B200_DUMP=<dir> (glue/xevd_b200_glue.c) and replays them through the CPU oracle:
tells a wrong work item (glue) from a wrong kernel when a stream decodes differently on the device (debugging aid, test infrastructure).
    python tools/glue_dump.py <dump dir> <stream.evc>"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle.pyoracle import Oracle  # noqa: E402
from xevd_b200 import synth  # noqa: E402
from xevd_b200 import xevd_api as X  # noqa: E402
from xevd_b200.abi import CU_DTYPE, EXT_DTYPE, Params  # noqa: E402
from xevd_b200.frame import CuList, HostPicture  # noqa: E402


def read_slice(path):
    b = Path(path).read_bytes()
    hdr = np.frombuffer(b, np.int32, 8)
    n_cu, n_ctu, n_ext, n_coef, n0, n1, stype, dbk = [int(v) for v in hdr[:8]]
    o = 32
    prm = Params.from_buffer_copy(b[o:o + C.sizeof(Params)]); o += C.sizeof(Params)
    pocs = np.frombuffer(b, np.int32, 42, o); o += 42 * 4
    cus = np.frombuffer(b, CU_DTYPE, n_cu, o).copy(); o += 32 * n_cu
    first = np.frombuffer(b, np.uint32, n_ctu + 1, o).copy(); o += 4 * (n_ctu + 1)
    ext = np.frombuffer(b, EXT_DTYPE, n_ext, o).copy(); o += 32 * n_ext
    coef = np.frombuffer(b, np.int16, n_coef, o).copy()
    cl = CuList(w=prm.w, h=prm.h, log2_ctu=prm.log2_ctu, cus=cus, ctu_first=first, coef=coef, ext=ext if n_ext else np.zeros(1, EXT_DTYPE))
    return prm, cl, pocs[:n0].tolist(), pocs[21:21 + n1].tolist(), stype, dbk


def stage_cmp(slice_path, stage, pic, cl):
    """the device picture the glue dumped after `stage` against the oracle's picture at the same point"""
    f = slice_path.with_name(slice_path.name.replace("slice_", "pic_").replace(".bin", f"_{stage}.bin"))
    if not f.exists():
        return
    b = f.read_bytes()
    w, h = np.frombuffer(b, np.int32, 2)
    a = np.frombuffer(b, np.int16, offset=8)
    planes = [a[:w * h].reshape(h, w), a[w * h:w * h * 5 // 4].reshape(h // 2, w // 2), a[w * h * 5 // 4:].reshape(h // 2, w // 2)]
    for pl, (g, want) in enumerate(zip(planes, pic.planes())):
        d = g != want
        if d.any():
            ys, xs = np.nonzero(d)
            s = 1 if pl == 0 else 2
            y, x = int(ys[0]), int(xs[0])
            own = [i for i, c in enumerate(cl.cus) if c["x"] <= x * s < c["x"] + (1 << c["log2w"]) and c["y"] <= y * s < c["y"] + (1 << c["log2h"])]
            print(f"\n   [{stage}] device != oracle, plane {pl}: {int(d.sum())} samples, first y={y} x={x} dev {int(g[y, x])} oracle {int(want[y, x])}, "
                  f"rows {ys.min()}..{ys.max()} cols {xs.min()}..{xs.max()}; CUs {[(i, cl.cus[i]) for i in own]}", end="")


def main():
    dump, stream = Path(sys.argv[1]), sys.argv[2]
    ref_pics = X.decode_stream(X.XevdLibrary(X.REF_SO), X.read_stream(stream))
    o = Oracle()
    done = {}
    for k, f in enumerate(sorted(dump.glob("slice_*.bin"))):
        prm, cl, p0, p1, stype, dbk = read_slice(f)
        cl.validate()
        refs0 = [done[p] for p in p0]
        refs1 = [done[p] for p in p1]
        pic = o.recon_frame(prm, HostPicture(prm.w, prm.h, prm.poc), refs0, refs1, cl)
        rec = pic.copy()
        stage_cmp(f, "recon", rec, cl)
        ids = {}
        rid = lambda lst: tuple(ids.setdefault(p, len(ids)) for p in lst)
        if dbk and not __import__('os').environ.get('NO_DBK'):
            o.deblock_frame(prm, pic, cl, synth.chroma_qp_table(bool(prm.tool_iqt)), bool(prm.tool_addb), (rid(p0) or (0,), rid(p1) or (0,)))
        stage_cmp(f, "dbk", pic, cl)
        o.pad(pic)
        done[prm.poc] = pic
        want = ref_pics[k] if k < len(ref_pics) else None
        print(f"{f.name}: poc {prm.poc} type {stype} {cl.n_cu} CUs, refs {p0} {p1}", end="")
        if want is not None:
            for pl, (a, b) in enumerate(zip(pic.planes(), want)):
                d = a != b
                if d.any():
                    ys, xs = np.nonzero(d)
                    s = 1 if pl == 0 else 2
                    y, x = int(ys[0]) * s, int(xs[0]) * s
                    own = [i for i, c in enumerate(cl.cus) if c["x"] <= x < c["x"] + (1 << c["log2w"]) and c["y"] <= y < c["y"] + (1 << c["log2h"])]
                    print(f"\n   plane {pl}: oracle(work list) != reference in {int(d.sum())} samples, first at {ys[0]},{xs[0]}; CUs there: {[ (i, cl.cus[i]) for i in own]}", end="")
        print()


if __name__ == "__main__":
    main()
