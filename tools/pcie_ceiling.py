#!/usr/bin/env python3
"""Ceiling of the end-to-end arm: pinned host<->device copies of the sizes one 4K picture moves (coefficients + descriptors in, 16-bit
picture out), both directions at once on separate streams, no kernels.  bench.py's e2e figure is compared against this in DESIGN.md.
    python tools/pcie_ceiling.py [--h2d-mb 25.9 --d2h-mb 24.9 --pictures 64]"""
import argparse
import json

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h2d-mb", type=float, default=25.93)
    ap.add_argument("--d2h-mb", type=float, default=24.88)
    ap.add_argument("--pictures", type=int, default=64)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n_in, n_out = int(args.h2d_mb * 1e6), int(args.d2h_mb * 1e6)
    slots = 3
    h_in = [torch.empty(n_in, dtype=torch.uint8).pin_memory() for _ in range(slots)]
    h_out = [torch.empty(n_out, dtype=torch.uint8).pin_memory() for _ in range(slots)]
    d_in = [torch.empty(n_in, dtype=torch.uint8, device=dev) for _ in range(slots)]
    d_out = [torch.zeros(n_out, dtype=torch.uint8, device=dev) for _ in range(slots)]
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(do_in, do_out, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s_in.wait_stream(torch.cuda.current_stream()); s_out.wait_stream(torch.cuda.current_stream())
        for i in range(n):
            if do_in:
                with torch.cuda.stream(s_in):
                    d_in[i % slots].copy_(h_in[i % slots], non_blocking=True)
            if do_out:
                with torch.cuda.stream(s_out):
                    h_out[i % slots].copy_(d_out[i % slots], non_blocking=True)
        torch.cuda.current_stream().wait_stream(s_in); torch.cuda.current_stream().wait_stream(s_out)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3

    run(True, True, 8)
    t_in, t_out, t_both = run(True, False, args.pictures), run(False, True, args.pictures), run(True, True, args.pictures)
    print(json.dumps({"h2d_alone_gbs": round(n_in * args.pictures / t_in / 1e9, 1), "d2h_alone_gbs": round(n_out * args.pictures / t_out / 1e9, 1),
                      "both_h2d_gbs": round(n_in * args.pictures / t_both / 1e9, 1), "both_d2h_gbs": round(n_out * args.pictures / t_both / 1e9, 1),
                      "pictures_per_s_ceiling": round(args.pictures / t_both, 1), "h2d_bytes_per_picture": n_in, "d2h_bytes_per_picture": n_out}))


if __name__ == "__main__":
    main()
