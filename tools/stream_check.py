#!/usr/bin/env python3
"""GPU box: decode streams with libxevd_gpu.so, dumping the work lists (XEVD_B200_DUMP) under gpurun_out/dump/<stream>/, and print where
the pictures differ from the recorded MD5s / the reference.  Bring gpurun_out/dump back and replay with tools/glue_dump.py."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from xevd_b200 import xevd_api as X  # noqa: E402

for path in sys.argv[1:]:
    path = Path(path)
    d = ROOT / "gpurun_out" / "dump" / path.stem
    d.mkdir(parents=True, exist_ok=True)
    for f in d.glob("slice_*.bin"):
        f.unlink()
    os.environ["XEVD_B200_DUMP"] = str(d)
    nals = X.read_stream(path)
    # one library instance per stream: the dump serial number is per process, so fork
    pid = os.fork()
    if pid == 0:
        gpu = X.decode_stream(X.XevdLibrary(X.GPU_SO), nals)
        ref = X.decode_stream(X.XevdLibrary(X.REF_SO), nals)
        for i, (a, b) in enumerate(zip(gpu, ref)):
            diffs = [int((x != y).sum()) for x, y in zip(a, b)]
            if any(diffs):
                print(f"{path.stem}: picture {i} differs {diffs}", flush=True)
        print(f"{path.stem}: {len(gpu)} pictures decoded", flush=True)
        os._exit(0)
    os.waitpid(pid, 0)
