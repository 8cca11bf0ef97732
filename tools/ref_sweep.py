#!/usr/bin/env python3
"""CPU-only: the oracle against the UNMODIFIED reference (oracle/_ref/libxevd_ref.so) over the randomised cases of tests/test_gpu_fuzz.py -
recon (planes), then deblocking and ALF on the oracle's maps, then padding.  The GPU sweep (tools/fuzz_sweep.py) compares the CUDA path with
the oracle on the same cases; this one closes the chain to the reference.  Needs /root/reference to have been compiled (oracle/Makefile).
    python tools/ref_sweep.py --first 0 --count 200 [--tree]"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def sweep(first, count, tree):
    """[(seed, stage)] of the cases where the oracle and the reference differ"""
    from oracle.pyoracle import Oracle, Reference
    from tests.test_gpu_fuzz import draw_case
    from xevd_b200 import synth
    from xevd_b200.frame import HostPicture
    o, r, bad = Oracle(), Reference(2), []
    for seed in range(first, first + count):
        k = draw_case(seed, tree=tree)
        prm, cl, refs, w, h = k["prm"], k["cl"], k["refs"], k["w"], k["h"]
        a = o.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
        b = r.recon_frame(prm, HostPicture(w, h, prm.poc), refs, refs[::-1], cl)
        stage = None
        if not all(np.array_equal(x, y) for x, y in zip(a.planes(), b.planes())):
            stage = "recon"
        else:
            for kk in ("map_mv", "map_refi", "map_scu", "map_unrefined_mv"):        # the harness has no slice-level state: maps are the oracle's
                getattr(b, kk)[...] = getattr(a, kk)
            tbl = synth.chroma_qp_table(k["main"])
            o.deblock_frame(prm, a, cl, tbl, bool(prm.tool_addb), ((0, 1), (1, 0)))
            r.deblock_frame(prm, b, cl, tbl, bool(prm.tool_addb), ((0, 1), (1, 0)))
            if not all(np.array_equal(x, y) for x, y in zip(a.planes(), b.planes())):
                stage = "deblock"
            elif k["alf"] is not None:
                o.alf_frame(prm, a, k["alf"], k["flags"])
                r.alf_frame(prm, b, k["alf"], k["flags"])
                if not all(np.array_equal(x, y) for x, y in zip(a.planes(), b.planes())):
                    stage = "alf"
            if stage is None:
                o.pad(a)
                r.pad(b)
                if not (np.array_equal(a.buf_y, b.buf_y) and np.array_equal(a.buf_u, b.buf_u) and np.array_equal(a.buf_v, b.buf_v)):
                    stage = "pad"
        if stage:
            bad.append((seed, stage))
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--count", type=int, default=100)
    ap.add_argument("--tree", action="store_true")
    args = ap.parse_args()
    bad = sweep(args.first, args.count, args.tree)
    print(f"oracle vs reference{' (tree)' if args.tree else ''} seeds {args.first}..{args.first + args.count - 1}: "
          f"{args.count - len(bad)} identical, mismatches: {bad}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
