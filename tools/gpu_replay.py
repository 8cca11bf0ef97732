#!/usr/bin/env python3
"""GPU box: replays work lists dumped by libxevd_gpu.so (XEVD_B200_DUMP, see tools/glue_dump.py) through the CUDA library AND the CPU
oracle, stage by stage (recon planes + maps, deblock, pad), every picture starting from the ORACLE's reference pictures so that one
difference does not hide the next.  Prints the first differing sample of every stage together with the CU that owns it.
    python tools/gpu_replay.py gpurun_out/dump/<stream> [...]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import Oracle  # noqa: E402
from tools.glue_dump import read_slice  # noqa: E402
from xevd_b200 import synth  # noqa: E402
from xevd_b200.device import Context  # noqa: E402
from xevd_b200.frame import HostPicture  # noqa: E402


def owners(cl, x, y, plane):
    out = []
    for i, c in enumerate(cl.cus):
        if c["x"] <= x < c["x"] + (1 << c["log2w"]) and c["y"] <= y < c["y"] + (1 << c["log2h"]):
            if (plane == 0 and c["flags"] & 1) or (plane > 0 and c["flags"] & 2) or not (c["flags"] & 3):
                out.append((i, {k: (c[k].tolist() if hasattr(c[k], "tolist") else c[k]) for k in c.dtype.names}))
    return out


def compare(tag, got, want, cl, verbose=True):
    bad = False
    for pl, (a, b) in enumerate(zip(got.planes(), want.planes())):
        d = a != b
        if d.any():
            bad = True
            ys, xs = np.nonzero(d)
            s = 1 if pl == 0 else 2
            y, x = int(ys[0]), int(xs[0])
            print(f"   {tag}: plane {pl}: {int(d.sum())} samples differ, first at y={y} x={x} (gpu {int(a[y, x])} oracle {int(b[y, x])})"
                  f" rows {int(ys.min())}..{int(ys.max())} cols {int(xs.min())}..{int(xs.max())}")
            if verbose:
                for i, c in owners(cl, x * s, y * s, pl):
                    print(f"      CU {i}: {c}")
    return bad


def compare_maps(tag, got, want):
    bad = False
    for name in ("map_scu", "map_refi", "map_mv", "map_unrefined_mv"):
        a, b = getattr(got, name), getattr(want, name)
        if not np.array_equal(a, b):
            bad = True
            idx = np.argwhere(a != b)[0].tolist()
            print(f"   {tag}: {name} differs in {int((a != b).sum())} entries, first at {idx}: gpu {a[tuple(idx)]} oracle {b[tuple(idx)]}")
    return bad


def main():
    o = Oracle()
    n_bad = 0
    for dump in sys.argv[1:]:
        dump = Path(dump)
        done = {}
        with Context(0) as ctx:
            for f in sorted(dump.glob("slice_*.bin")):
                prm, cl, p0, p1, stype, dbk = read_slice(f)
                print(f"{dump.name}/{f.name}: poc {prm.poc} type {stype} {cl.n_cu} CUs refs {p0} {p1} ctu {1 << prm.log2_ctu}", flush=True)
                refs0 = [done[p] for p in p0]
                refs1 = [done[p] for p in p1]
                want = o.recon_frame(prm, HostPicture(prm.w, prm.h, prm.poc), refs0, refs1, cl)
                dev = {p: ctx.pic_alloc(prm.w, prm.h).upload(done[p]) for p in set(p0 + p1)}
                for p, d in dev.items():
                    d.upload_maps(done[p])
                cur = ctx.pic_alloc(prm.w, prm.h)
                tbl = synth.chroma_qp_table(bool(prm.tool_iqt))
                ctx.set_chroma_qp_table(tbl)
                ctx.recon_frame(prm, cur, [dev[p] for p in p0], [dev[p] for p in p1], cl)
                got = cur.download(maps=True)
                bad = compare("recon", got, want, cl)
                bad |= compare_maps("recon", got, want)
                mf = f.with_name(f.name.replace("slice_", "pic_").replace(".bin", "_maps.bin"))
                if mf.exists():     # the maps the glue's device picture held before deblocking
                    b = mf.read_bytes()
                    n = int(np.frombuffer(b, np.int32, 1)[0])
                    o8 = 8
                    gm = dict(map_mv=np.frombuffer(b, np.int16, n * 4, o8), map_refi=np.frombuffer(b, np.int8, n * 2, o8 + n * 8),
                              map_scu=np.frombuffer(b, np.uint32, n, o8 + n * 10), map_unrefined_mv=np.frombuffer(b, np.int16, n * 4, o8 + n * 14),
                              edge=np.frombuffer(b, np.uint8, n, o8 + n * 22))
                    mine = dict(map_mv=got.map_mv, map_refi=got.map_refi, map_scu=got.map_scu, map_unrefined_mv=got.map_unrefined_mv,
                                edge=cur.download_edge_map())
                    for k in gm:
                        a, bb = gm[k].ravel(), np.asarray(mine[k]).ravel()
                        if not np.array_equal(a, bb):
                            i = int(np.argwhere(a != bb)[0][0])
                            print(f"   glue-vs-replay: {k} differs in {int((a != bb).sum())} entries, first {i}: glue {a[i]} replay {bb[i]}")
                pf = f.with_name(f.name.replace("slice_", "pic_").replace(".bin", "_dbkprm.bin"))
                if pf.exists():
                    import ctypes as C
                    from xevd_b200.abi import Params
                    p2 = Params.from_buffer_copy(pf.read_bytes()[:C.sizeof(Params)])
                    for fld, _ in Params._fields_:
                        if getattr(p2, fld) != getattr(prm, fld):
                            print(f"   deblock-time params: {fld} = {getattr(p2, fld)} (recon-time {getattr(prm, fld)})")
                if dbk:
                    ids = {}
                    rid = lambda lst: tuple(ids.setdefault(p, len(ids)) for p in lst)  # noqa: E731
                    if bad:         # continue from the oracle's reconstruction so that deblocking is judged on its own
                        cur.upload(want, padded=False)
                        cur.upload_maps(want)
                    o.deblock_frame(prm, want, cl, tbl, bool(prm.tool_addb), (rid(p0) or (0,), rid(p1) or (0,)))
                    ctx.deblock(prm, cur, [dev[p] for p in p0], [dev[p] for p in p1])
                    got = cur.download()
                    bad |= compare("deblock", got, want, cl)
                o.pad(want)
                done[prm.poc] = want
                n_bad += bool(bad)
                for d in dev.values():
                    d.free()
                cur.free()
    print(f"{n_bad} pictures differ")
    return 1 if n_bad else 0


if __name__ == "__main__":
    sys.exit(main())
