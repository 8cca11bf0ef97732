#!/usr/bin/env python3
"""Generates tests/golden/streams/*.evc (+ .md5 of every decoded picture as the UNMODIFIED reference produces it).

The reference ships no streams and the box has no encoder: the streams come from tools/evcgen (the reference's own
syntax parser run as a generator).  Every stream is kept only if (1) the unmodified reference decodes it to the generator's own
pictures, and (2) the reference's application built with AddressSanitizer (tools/evcgen `make asan`) decodes it without touching
memory out of bounds - i.e. it is a stream the reference decoder is well defined on.
Run (needs /root/reference):  make -C tools/evcgen all asan && python tests/golden/make_streams.py"""
import hashlib
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools" / "evcgen"))
import evcgen as G  # noqa: E402
from xevd_b200 import xevd_api as X  # noqa: E402

OUT = Path(__file__).resolve().parent / "streams"
ASAN = ROOT / "tools" / "evcgen" / "_build" / "xevd_app_asan"

# name: (tool set, tool overrides, generator arguments)
STREAMS = {
    "base_i_64x64_8b": ("baseline", {}, dict(w=64, h=64, bd=8, frames=4, seed=1, types="I")),
    "base_ipb_128x64_8b": ("baseline", {}, dict(w=128, h=64, bd=8, frames=8, seed=3, types="IPB")),
    "base_ipbb_320x192_10b": ("baseline", {}, dict(w=320, h=192, bd=10, frames=8, seed=5, types="IPBB")),
    "base_ibp_256x128_8b_qp38": ("baseline", {}, dict(w=256, h=128, bd=8, frames=6, seed=9, types="IBP", qp=38)),
    "base_ipp_200x120_10b_nodbk": ("baseline", {}, dict(w=200, h=120, bd=10, frames=5, seed=11, types="IPP", deblock=0)),
    "base_ipb_416x240_8b_cip": ("baseline", {}, dict(w=416, h=240, bd=8, frames=6, seed=13, types="IPB", constrained_intra=1)),
    # BASELINE.json config 1: a Baseline-profile 1080p 8-bit stream (also the drop-in timing case of bench.py)
    "base_1080p_8b": ("baseline", {}, dict(w=1920, h=1080, bd=8, frames=8, seed=31, types="IPBB")),
    "main_1080p_10b": ("main", {}, dict(w=1920, h=1080, bd=10, frames=6, seed=32, types="IBB", lps_scale=350)),
    # Main profile: BTT + SUCO partitions, ADMVP (1/16-pel), affine, AMVR, DMVR, MMVD, HMVP, EIPD, CM_INIT + ADCC, IQT, ATS, ADDB, HTDF
    "main_all_256x128_10b": ("main", {}, dict(w=256, h=128, bd=10, frames=6, seed=21, types="IPP")),
    "main_all_320x192_8b_lps": ("main", {}, dict(w=320, h=192, bd=8, frames=6, seed=22, types="IBB", lps_scale=420)),
    "main_ctu128_384x256_10b": ("main", {}, dict(w=384, h=256, bd=10, frames=5, seed=23, types="IBB", log2_ctu=7, lps_scale=350)),
    "main_ctu32_200x120_8b": ("main", {}, dict(w=200, h=120, bd=8, frames=5, seed=24, types="IBB", log2_ctu=5, lps_scale=350)),
    "main_noaddb_nohtdf_256x144_10b": ("main", dict(addb=0, htdf=0), dict(w=256, h=144, bd=10, frames=5, seed=25, types="IBB", lps_scale=350)),
    "main_intra_eipd_ats_192x128_10b": ("main", {}, dict(w=192, h=128, bd=10, frames=4, seed=26, types="I", lps_scale=400)),
    # ALF: one random adaptation parameter set per picture (5x5 / 7x7 luma, 1..25 filters, fixed-filter patterns, chroma), CTB flags
    "main_alf_256x128_10b": ("main", dict(alf=1), dict(w=256, h=128, bd=10, frames=6, seed=41, types="IPB", lps_scale=350)),
    "main_alf_416x240_8b": ("main", dict(alf=1), dict(w=416, h=240, bd=8, frames=5, seed=42, types="IBB", lps_scale=350)),
    "main_alf_ctu128_384x256_10b": ("main", dict(alf=1), dict(w=384, h=256, bd=10, frames=4, seed=43, types="IPP", log2_ctu=7, lps_scale=350)),
    # IBC: block vectors chosen by the generator among the decoded, in-picture positions at or left of / above the CU's CTU
    "main_ibc_256x128_10b": ("main", dict(ibc=1), dict(w=256, h=128, bd=10, frames=5, seed=52, types="IPB", lps_scale=350)),
    "main_ibc_alf_320x192_8b": ("main", dict(ibc=1, alf=1), dict(w=320, h=192, bd=8, frames=5, seed=53, types="IBB", lps_scale=350)),
    "main_ibc_i_ctu128_256x256_10b": ("main", dict(ibc=1), dict(w=256, h=256, bd=10, frames=3, seed=54, types="I", log2_ctu=7, lps_scale=400)),
    # tiles: every tile is its own arithmetic code word, entry points in the slice header; one slice per picture.  across = the PPS's
    # loop_filter_across_tiles_enabled_flag (deblocking of tile-boundary edges, ALF margins at tile borders)
    "base_tiles2x2_320x192_8b": ("baseline", {}, dict(w=320, h=192, bd=8, frames=6, seed=61, types="IPB", tiles=dict(cols=2, rows=2, across=1))),
    "base_tiles3x1_noacross_416x240_10b": ("baseline", {}, dict(w=416, h=240, bd=10, frames=6, seed=62, types="IPBB", tiles=dict(cols=3, rows=1, across=0))),
    "main_tiles2x2_alf_noacross_384x256_10b": ("main", dict(alf=1), dict(w=384, h=256, bd=10, frames=5, seed=63, types="IPB", lps_scale=350,
                                                                     tiles=dict(cols=2, rows=2, across=0))),
    "main_tiles2x2_alf_across_320x192_8b": ("main", dict(alf=1), dict(w=320, h=192, bd=8, frames=5, seed=64, types="IBB", lps_scale=350,
                                                                   tiles=dict(cols=2, rows=2, across=1))),
    "main_tiles_explicit_ctu32_alf_256x160_10b": ("main", dict(alf=1), dict(w=256, h=160, bd=10, frames=5, seed=65, types="IPP", log2_ctu=5, lps_scale=350,
                                                                         tiles=dict(cols=3, rows=2, col_w=[2, 5, 1], row_h=[4, 1], across=0))),
    # several slices per picture, each a rectangle of tiles.  All pictures are IDR: without sps_pocs the reference derives a new POC for
    # every non-IDR slice NAL (src_main/xevdm.c:3037-3041), so it cannot decode a multi-slice P / B picture of such a sequence itself
    "base_slices2_tiles2x2_idr_320x192_8b": ("baseline", {}, dict(w=320, h=192, bd=8, frames=4, seed=66, types="I", gop=1,
                                                                   tiles=dict(cols=2, rows=2, across=0), slices=[(0, 1), (2, 3)])),
    "main_slices2_tiles3x2_alf_idr_384x256_10b": ("main", dict(alf=1), dict(w=384, h=256, bd=10, frames=4, seed=67, types="I", gop=1, lps_scale=350,
                                                                         tiles=dict(cols=3, rows=2, across=1), slices=[(0, 3), (1, 5)])),
    "main_slices3_tiles2x2_alf_idr_256x256_8b": ("main", dict(alf=1), dict(w=256, h=256, bd=8, frames=3, seed=68, types="I", gop=1, lps_scale=350,
                                                                        tiles=dict(cols=2, rows=2, across=0), slices=[(0, 0), (1, 1), (2, 3)])),
    "main_nodmvr_noaffine_256x128_8b_cip": ("main", dict(dmvr=0, affine=0), dict(w=256, h=128, bd=8, frames=5, seed=27, types="IPP", lps_scale=350, constrained_intra=1)),
}


def pic_md5(p):
    return hashlib.md5(p[0].tobytes() + p[1].tobytes() + p[2].tobytes()).hexdigest()


def main():
    only = set(sys.argv[1:])
    g = G.Generator()
    ref = X.XevdLibrary(X.REF_SO)
    OUT.mkdir(exist_ok=True)
    for name, (profile, over, kw) in STREAMS.items():
        if only and name not in only:
            continue
        tools = dict(G.BASELINE if profile == "baseline" else G.MAIN)
        tools.update(over)
        kw = dict(kw)
        for attempt in range(20):
            nals, own = g.make(tools, **kw)
            path = OUT / f"{name}.evc"
            X.write_stream(path, nals)
            pics = X.decode_stream(ref, nals)
            ok = G.same_pictures(own, pics)
            why = "" if ok else "reference != generator"
            if ok and ASAN.exists():
                # the ASan build is also the reference with its PLAIN-C kernels (X86 undefined): its pictures must be the dispatched AVX2
                # build's pictures, or the reference is not one decoder on this stream (e.g. IQT coefficients beyond 32 of a 64-point block)
                yuv = Path("/tmp") / f"{name}.asan.yuv"
                r = subprocess.run([str(ASAN), "-i", str(path), "-o", str(yuv), "--output-bit-depth", "10"], capture_output=True, text=True)      # 10: the 16-bit samples as they are
                if "ERROR: AddressSanitizer" in r.stderr or r.returncode != 0:
                    ok, why = False, "AddressSanitizer: " + next((l for l in r.stderr.splitlines() if "SUMMARY" in l), f"rc {r.returncode}")
                else:
                    want = b"".join(p.astype("<i2").tobytes() for pic in pics for p in pic)
                    if yuv.read_bytes() != want:
                        ok, why = False, "plain-C reference != dispatched (AVX2) reference"
                yuv.unlink(missing_ok=True)
            if ok:
                (OUT / f"{name}.md5").write_text("".join(pic_md5(p) + "\n" for p in pics))
                print(f"{name}: {len(nals)} NAL units, {path.stat().st_size} bytes, {len(pics)} pictures (seed {kw['seed']})")
                break
            print(f"{name}: seed {kw['seed']} rejected ({why})")
            kw["seed"] += 1000
        else:
            path.unlink(missing_ok=True)
            print(f"{name}: no conforming stream found")


if __name__ == "__main__":
    main()
