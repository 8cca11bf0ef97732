import sys, time, os
sys.path.insert(0, '/root/repo')
os.environ["XEVD_B200_STATS"] = "1"
from xevd_b200 import xevd_api as X
for name in sys.argv[1:]:
    nals = X.read_stream(f"/root/repo/tests/golden/streams/{name}.evc")
    for so in (X.GPU_SO, X.REF_SO):
        lib = X.XevdLibrary(so)
        for rep in range(2):
            t0 = time.perf_counter()
            d = X.Decoder(lib)
            t1 = time.perf_counter()
            n = 0
            ts = []
            for nal in nals:
                ta = time.perf_counter()
                ret, stat = d.decode(nal)
                while True:
                    p = d.pull()
                    if p is None: break
                    n += 1
                ts.append(time.perf_counter() - ta)
            t2 = time.perf_counter()
            d.close()
            t3 = time.perf_counter()
            print(f"{name} {so.name} rep {rep}: create {t1-t0:.3f}s decode {t2-t1:.3f}s ({n} pics) delete {t3-t2:.3f}s per-NAL ms {[round(1e3*t,1) for t in ts]}", flush=True)
