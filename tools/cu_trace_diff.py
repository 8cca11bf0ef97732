#!/usr/bin/env python3
"""GPU box: decodes streams with libxevd_gpu.so and with libxevd_reftrace.so (the unmodified reference + one logging wrapper), both
writing one line per coding unit (glue/cu_trace.h), and prints the first lines that differ: the first CU whose host-side state (mode,
reference indices, control points, published vectors ...) differs between the drop-in library and the reference (debugging aid).
    python tools/cu_trace_diff.py tests/golden/streams/main_*.evc"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from xevd_b200 import xevd_api as X  # noqa: E402

TRACE_SO = ROOT / "glue" / "_build" / "libxevd_reftrace.so"


def run(so, stream, log):
    pid = os.fork()
    if pid == 0:
        os.environ["XEVD_CU_TRACE"] = str(log)
        try:
            X.decode_stream(X.XevdLibrary(so), X.read_stream(stream))
        finally:
            os._exit(0)
    os.waitpid(pid, 0)
    return Path(log).read_text().splitlines() if Path(log).exists() else []


for path in sys.argv[1:]:
    path = Path(path)
    out = ROOT / "gpurun_out" / "trace"
    out.mkdir(parents=True, exist_ok=True)
    a = run(X.GPU_SO, path, out / f"{path.stem}.gpu.txt")
    b = run(TRACE_SO, path, out / f"{path.stem}.ref.txt")
    n = 0
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y:
            print(f"{path.stem}: line {i}\n   gpu {x}\n   ref {y}")
            n += 1
            if n >= 4:
                break
    print(f"{path.stem}: {len(a)} / {len(b)} CUs, {'identical' if a == b else 'DIFFERENT'}", flush=True)
