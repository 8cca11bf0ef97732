#!/usr/bin/env python3
"""Long randomised parity sweep on one GPU: the cases of tests/test_gpu_fuzz.py for seeds [--first, --first + --count) through
recon -> deblock -> ALF -> pad, final padded pictures against the oracle.  Prints the mismatching seeds (none expected).
    python tools/fuzz_sweep.py --first 64 --count 400 [--tree]"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    from oracle.pyoracle import Oracle
    from tests.test_gpu_fuzz import draw_case
    from xevd_b200 import synth
    from xevd_b200.device import Context
    from xevd_b200.frame import HostPicture
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=64)
    ap.add_argument("--count", type=int, default=200)
    ap.add_argument("--tree", action="store_true", help="Main pictures with local dual tree nodes / constrained intra mixed in (draw_case(seed, tree=True))")
    args = ap.parse_args()
    o, c, bad, skipped = Oracle(), Context(0), [], 0
    for seed in range(args.first, args.first + args.count):
        k = draw_case(seed, tree=args.tree)
        if args.tree and not k["main"]:
            skipped += 1
            continue
        prm, w, h = k["prm"], k["w"], k["h"]
        want = o.recon_frame(prm, HostPicture(w, h, prm.poc), k["refs"], k["refs"][::-1], k["cl"])
        dr = [c.pic_alloc(w, h).upload(r) for r in k["refs"]]
        cur = c.pic_alloc(w, h)
        tbl = synth.chroma_qp_table(k["main"])
        c.set_chroma_qp_table(tbl)
        c.recon_frame(prm, cur, dr, dr[::-1], k["cl"])
        o.deblock_frame(prm, want, k["cl"], tbl, bool(prm.tool_addb), ((0, 1), (1, 0)))
        c.deblock(prm, cur, dr, dr[::-1])
        if k["alf"] is not None:
            o.alf_frame(prm, want, k["alf"], k["flags"])
            c.alf(prm, cur, k["alf"], k["flags"])
        o.pad(want)
        c.pad(cur)
        out = cur.download_padded()
        if not (np.array_equal(out.buf_y, want.buf_y) and np.array_equal(out.buf_u, want.buf_u) and np.array_equal(out.buf_v, want.buf_v)):
            bad.append(seed)
        for p in dr + [cur]:
            p.free()
    print(f"fuzz sweep{' (tree)' if args.tree else ''} seeds {args.first}..{args.first + args.count - 1}: {args.count - skipped - len(bad)} bit-exact, "
          f"{skipped} Baseline draws skipped, mismatching seeds: {bad}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
