#!/usr/bin/env python3
"""evcgen -- random but syntactically valid MPEG-5 EVC elementary streams for decoder parity tests (test tooling).

There is no EVC encoder and no conformance stream on the build box.  This tool writes the parameter sets and
slice headers bit by bit (field order: xevdm_eco_sps / xevdm_eco_pps / xevdm_eco_sh, src_main/xevdm_eco.c:1847,2006,2510) and lets the
reference's OWN slice-data parser generate the slice data: tools/evcgen/_build/libxevd_gen.so is the unmodified decoder whose CABAC
decoding primitives choose-and-encode instead of decode (gen_engine.c).  Every stream is then decoded by the unmodified reference
(oracle/_ref/libxevd_ref.so) as a self-check: it must parse without error and reproduce the generator's own pictures.

    python tools/evcgen/evcgen.py --out tests/golden/streams/base_64x64.evc --profile baseline --w 64 --h 64 --frames 6
"""
from __future__ import annotations

import argparse
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from xevd_b200 import xevd_api as X  # noqa: E402

GEN_SO = Path(__file__).resolve().parent / "_build" / "libxevd_gen.so"


class BitWriter:
    def __init__(self):
        self.bits = []

    def u(self, v, n):
        for i in range(n - 1, -1, -1):
            self.bits.append((v >> i) & 1)

    def ue(self, v):
        v += 1
        n = v.bit_length()
        self.u(0, n - 1)
        self.u(v, n)

    def se(self, v):
        self.ue(2 * v - 1 if v > 0 else -2 * v)

    def align(self):
        while len(self.bits) % 8:
            self.bits.append(0)

    def bytes(self):
        self.align()
        return bytes(int("".join(map(str, self.bits[i:i + 8])), 2) for i in range(0, len(self.bits), 8))


def nal_header(nut, tid=0):
    b = BitWriter()
    b.u(0, 1); b.u(nut + 1, 6); b.u(tid, 3); b.u(0, 5); b.u(0, 1)
    return b


# tool sets -------------------------------------------------------------------------------------------------------------------
BASELINE = dict(btt=0, suco=0, admvp=0, affine=0, amvr=0, dmvr=0, mmvd=0, hmvp=0, eipd=0, ibc=0, cm_init=0, adcc=0, iqt=0, ats=0,
                addb=0, alf=0, htdf=0, rpl=0, pocs=0, dquant=0, dra=0)
MAIN = dict(btt=1, suco=1, admvp=1, affine=1, amvr=1, dmvr=1, mmvd=1, hmvp=1, eipd=1, ibc=0, cm_init=1, adcc=1, iqt=1, ats=1,
            addb=1, alf=0, htdf=1, rpl=0, pocs=0, dquant=0, dra=0)


def write_sps(t, w, h, bd, max_refs=2, log2_ctu=6):
    b = nal_header(X.NUT_SPS)
    b.ue(0)                                # sps_seq_parameter_set_id
    b.u(1 if t["btt"] or t["admvp"] or t["eipd"] or t["iqt"] else 0, 8)          # profile_idc (0 baseline, 1 main)
    b.u(51, 8)                             # level_idc
    b.u(0, 32); b.u(0, 32)                 # toolset_idc_h / _l (not interpreted by the decoder)
    b.ue(1)                                # chroma_format_idc 4:2:0
    b.ue(w); b.ue(h)
    b.ue(bd - 8); b.ue(bd - 8)
    b.u(t["btt"], 1)
    if t["btt"]:
        b.ue(log2_ctu - 5)                 # log2_ctu_size_minus5
        b.ue(0)                            # log2_min_cb_size_minus2
        b.ue(0)                            # log2_diff_ctu_max_14_cb_size
        b.ue(0)                            # log2_diff_ctu_max_tt_cb_size
        b.ue(0)                            # log2_diff_min_cb_min_tt_cb_size_minus2
    b.u(t["suco"], 1)
    if t["suco"]:
        b.ue(0); b.ue(2 if log2_ctu > 5 else 1)      # log2_diff_ctu_size_max_suco_cb_size, log2_diff_max_suco_min_suco_cb_size
    b.u(t["admvp"], 1)
    if t["admvp"]:
        for k in ("affine", "amvr", "dmvr", "mmvd", "hmvp"):
            b.u(t[k], 1)
    b.u(t["eipd"], 1)
    if t["eipd"]:
        b.u(t["ibc"], 1)
        if t["ibc"]:
            b.ue(2)
    b.u(t["cm_init"], 1)
    if t["cm_init"]:
        b.u(t["adcc"], 1)
    b.u(t["iqt"], 1)
    if t["iqt"]:
        b.u(t["ats"], 1)
    b.u(t["addb"], 1); b.u(t["alf"], 1); b.u(t["htdf"], 1); b.u(t["rpl"], 1); b.u(t["pocs"], 1); b.u(t["dquant"], 1); b.u(t["dra"], 1)
    if t["pocs"]:
        b.ue(4)
    if not t["rpl"] or not t["pocs"]:
        b.ue(0)                            # log2_sub_gop_length = 0: low delay
        b.ue(0)                            # log2_ref_pic_gap_length
    assert not t["rpl"]
    b.ue(max_refs)                         # max_num_ref_pics
    b.u(0, 1)                              # picture_cropping_flag
    b.u(0, 1)                              # chroma_qp_table_present_flag
    b.u(0, 1)                              # vui_parameters_present_flag
    return b.bytes()


def write_pps(constrained_intra=0, cu_qp_delta=0, tiles=None):
    """tiles: None (one tile) or dict(cols, rows, col_w=[...] / row_h=[...] in CTUs for explicit spacing, across=0/1)
    (xevdm_eco_pps, src_main/xevdm_eco.c:2019-2052)"""
    b = nal_header(X.NUT_PPS)
    b.ue(0); b.ue(0)                       # pps id, sps id
    b.ue(0); b.ue(0)                       # num_ref_idx_default_active_minus1[0..1]
    b.ue(0)                                # additional_lt_poc_lsb_len
    b.u(0, 1)                              # rpl1_idx_present_flag
    if tiles is None:
        b.u(1, 1)                          # single_tile_in_pic_flag
        b.ue(0)                            # tile_id_len_minus1
    else:
        b.u(0, 1)
        b.ue(tiles["cols"] - 1); b.ue(tiles["rows"] - 1)
        uniform = "col_w" not in tiles
        b.u(1 if uniform else 0, 1)        # uniform_tile_spacing_flag
        if not uniform:
            for v in tiles["col_w"][:-1]:
                b.ue(v - 1)                # tile_column_width_minus1
            for v in tiles["row_h"][:-1]:
                b.ue(v - 1)                # tile_row_height_minus1
        b.u(tiles.get("across", 0), 1)     # loop_filter_across_tiles_enabled_flag
        b.ue(TILE_OFFSET_BITS - 1)         # tile_offset_lens_minus1
        b.ue(tile_id_bits(tiles) - 1)      # tile_id_len_minus1
    b.u(0, 1)                              # explicit_tile_id_flag
    b.u(0, 1)                              # pic_dra_enabled_flag
    b.u(0, 1)                              # arbitrary_slice_present_flag
    b.u(constrained_intra, 1)
    b.u(cu_qp_delta, 1)
    if cu_qp_delta:
        b.ue(0)                            # cu_qp_delta_area - 6
    return b.bytes()


TILE_OFFSET_BITS = 24                      # entry points are written with a fixed width, so they can be filled in after the tiles are generated


def tile_id_bits(tiles):
    return max(1, int(tiles["cols"] * tiles["rows"] - 1).bit_length())


GOLOMB_IDX = {5: [0, 0, 1, 0, 0, 1], 7: [0, 0, 1, 0, 0, 1, 2, 1, 0, 0, 1, 2]}      # golombIdx5 / golombIdx7 (src_main/xevdm_alf.h:165-178)


def alf_golomb(b, v, k, signed):
    """the code xevdm_alfGolombDecode reads (src_main/xevdm_eco.c:2154-2187): unary prefix of zeros, a one, prefix + k suffix bits, sign"""
    a = abs(v)
    n = 0
    while a >= ((1 << (n + 1)) - 1) << k:
        n += 1
    b.u(0, n); b.u(1, 1)
    if n + k > 0:
        b.u(a - (((1 << n) - 1) << k), n + k)
    if signed and a:
        b.u(1 if v > 0 else 0, 1)


def write_alf_filter(b, rng, chroma, n_filters, size):
    """xevdm_eco_alf_filter (src_main/xevdm_eco.c:2224-2320)"""
    delta_flag = 0
    flags = [1] * n_filters
    if not chroma:
        delta_flag = int(rng.integers(0, 2))
        b.u(delta_flag, 1)                                     # alf_coefficients_delta_flag
        if not delta_flag and n_filters > 1:
            b.u(int(rng.integers(0, 2)), 1)                    # coeff_delta_pred_mode_flag
    kmin = int(rng.integers(0, 3))
    b.ue(kmin)                                                 # alf_luma_min_eg_order_minus1
    k, ktab = kmin + 1, []
    for _ in range(2 if size == 5 else 3):
        inc = int(rng.integers(0, 2))
        b.u(inc, 1)                                            # alf_eg_order_increase_flag
        k += inc
        ktab.append(k)
    if not chroma and delta_flag:
        flags = [int(rng.integers(0, 2)) for _ in range(n_filters)]
        for fl in flags:
            b.u(fl, 1)                                         # filter_coefficient_flag
    n_coef = size * size // 4                                  # numCoeff - 1
    for i in range(1 if chroma else n_filters):
        if not flags[i]:
            continue
        for c in range(n_coef):
            alf_golomb(b, int(rng.integers(-24, 25)), ktab[GOLOMB_IDX[size][c]], True)


def write_aps_alf(aps_id, rng, tid=0):
    """an ALF adaptation parameter set with random content (xevdm_eco_aps_gen / xevdm_eco_alf_aps_param, src_main/xevdm_eco.c:2081-2477)"""
    b = nal_header(X.NUT_APS, tid)
    b.u(aps_id, 5); b.u(0, 3)                                  # aps id, aps_type_id 0 = ALF
    luma, chroma = 1, int(rng.integers(0, 2))
    b.u(luma, 1); b.u(chroma, 1)                               # alf_luma_filter_signal_flag, alf_chroma_filter_signal_flag
    n_filters = int(rng.choice([1, 2, 3, 8, 25]))
    b.ue(n_filters - 1)                                        # alf_luma_num_filters_signalled_minus1
    size = int(rng.choice([5, 7]))
    b.u(1 if size == 7 else 0, 1)                              # alf_luma_type_flag
    if n_filters > 1:
        bits = int(n_filters - 1).bit_length()                 # xevd_tbl_log2[n - 1] + 1
        for _ in range(25):
            b.u(int(rng.integers(0, n_filters)), bits)         # alf_luma_coeff_delta_idx
    pattern = int(rng.integers(0, 3))
    alf_golomb(b, pattern, 0, False)                           # alf_luma_fixed_filter_usage_pattern
    use = [1] * 25 if pattern == 1 else [0] * 25
    if pattern == 2:
        use = [int(rng.integers(0, 2)) for _ in range(25)]
        for u_ in use:
            b.u(u_, 1)
    if pattern > 0:
        for u_ in use:
            if u_:
                b.u(int(rng.integers(0, 16)), 4)               # alf_luma_fixed_filter_set_idx
    write_alf_filter(b, rng, False, n_filters, size)
    if chroma:
        write_alf_filter(b, rng, True, 1, 5)
    b.u(0, 1)                                                  # aps_extension_flag
    return b.bytes(), chroma


def write_sh(t, nut, slice_type, qp, deblock=1, alpha=0, beta=0, qp_u_off=0, qp_v_off=0, tid=0, alf=None, tiles=None, tile_range=None, entry=None):
    """tiles: the PPS tile description; tile_range = (first_tile_id, last_tile_id) of this slice; entry: byte sizes of the slice's tiles
    but the last (None while they are not known yet: the generator only needs a header of the right length)"""
    b = nal_header(nut, tid)
    b.ue(0)                                # slice_pic_parameter_set_id
    n_tiles = 1
    if tiles is not None:
        first, last = tile_range
        n_tiles = tiles_in_slice(tiles, first, last)
        # single_tile_in_slice_flag: always 0 - with 1 the parser leaves last_tile_id at the previous slice's value (xevdm_eco.c:2530-2539)
        # and the rectangle it derives (:2553-2590) is wrong for every single-tile slice but the first
        b.u(0, 1)
        b.u(first, tile_id_bits(tiles))    # first_tile_id
        b.u(last, tile_id_bits(tiles))     # last_tile_id (arbitrary_slice_present_flag = 0)
    b.ue(slice_type)
    if nut == X.NUT_IDR:
        b.u(0, 1)                          # no_output_of_prior_pics_flag
    if t["mmvd"] and slice_type in (X.ST_B, X.ST_P):
        b.u(1, 1)                          # mmvd_group_enable_flag
    if t["alf"]:
        b.u(1 if alf else 0, 1)            # alf_on
        if alf:
            aps_id, ctb_on, chroma_idc = alf
            b.u(aps_id, 5)                 # aps_id_y
            b.u(ctb_on, 1)                 # is_ctb_alf_on: per-CTB enable flags follow in the slice data (CABAC)
            b.u(chroma_idc, 2)             # alf_chroma_idc: bit 0 Cb, bit 1 Cr
            if chroma_idc:
                b.u(aps_id, 5)             # aps_id_ch
    if slice_type != X.ST_I:
        b.u(0, 1)                          # num_ref_idx_active_override_flag
        if t["admvp"]:
            b.u(0, 1)                      # temporal_mvp_asigned_flag
    b.u(deblock, 1)
    if deblock and t["addb"]:
        b.se(alpha); b.se(beta)
    b.u(qp, 6)
    b.se(qp_u_off); b.se(qp_v_off)
    for i in range(n_tiles - 1):           # entry_point_offset_minus1 (xevdm_eco.c:2789-2795)
        b.u((entry[i] - 1) if entry else 255, TILE_OFFSET_BITS)
    return b.bytes()


def tiles_in_slice(tiles, first, last):
    """rectangle of tiles from first to last (xevdm_eco_sh, src_main/xevdm_eco.c:2553-2590)"""
    wt = tiles["cols"]
    return ((last % wt) - (first % wt) + 1) * ((last // wt) - (first // wt) + 1)


class NonConforming(Exception):
    """the random syntax described something no encoder emits (see gen_engine.c, conformance watch): discard the stream"""


class Generator:
    def __init__(self):
        if not GEN_SO.exists():
            raise RuntimeError(f"{GEN_SO} missing: make -C tools/evcgen (needs /root/reference)")
        self.lib = X.XevdLibrary(GEN_SO)
        L = self.lib.lib
        L.gen_reset.argtypes = [C.c_uint64, C.c_int, C.c_int]
        L.gen_reset.restype = None
        L.gen_take.argtypes = [C.c_void_p, C.c_size_t]
        L.gen_take.restype = C.c_size_t
        L.gen_selfcheck.restype = C.c_longlong
        L.gen_invalid.argtypes = [C.c_char_p, C.c_int]
        L.gen_replay.argtypes = [C.c_void_p, C.c_size_t]
        L.gen_replay.restype = None
        L.gen_log.argtypes = [C.c_void_p, C.c_size_t]
        L.gen_log.restype = C.c_size_t
        L.gen_tile_sizes.argtypes = [C.c_void_p, C.c_int]
        L.gen_tile_sizes.restype = C.c_int

    def make(self, tools, w, h, bd, frames, seed, types="IPB", qp=30, lps_scale=256, ep_one=128, log2_ctu=6, deblock=1, gop=0, tries=40, tiles=None,
             slices=None, **pps_kw):
        """returns (list of NAL units, pictures the generator's own decode produced in output order).
        Slice by slice: draw a random slice; if the conformance watch of gen_engine.c objects, a FRESH decoder instance is brought to
        the same state by replaying the recorded bins of the slices accepted so far, and the slice is drawn again with another seed.
        tiles: the PPS tile grid (write_pps); slices: the (first_tile_id, last_tile_id) rectangles a picture is cut into (default: one slice)"""
        rng = np.random.default_rng(seed)
        L = self.lib.lib
        tail = bytes(max(1 << 17, w * h * 8))
        params = [write_sps(tools, w, h, bd, log2_ctu=log2_ctu), write_pps(tiles=tiles, **pps_kw)]
        n_tiles = tiles["cols"] * tiles["rows"] if tiles else 1
        slices = slices or [(0, n_tiles - 1)]
        units = []                         # one per slice NAL: dict(f, first (of its picture), aps, sh (write_sh arguments), hdr, n_tiles)
        for f in range(frames):
            idr = f == 0 or (gop and f % gop == 0)
            st = X.ST_I if (idr or len(types) == 1) else {"I": X.ST_I, "P": X.ST_P, "B": X.ST_B}[types[1 + (f - 1) % (len(types) - 1)]]
            alf, aps = None, None
            if tools["alf"] and rng.random() < 0.85:             # most pictures filtered, each with its own parameter set
                aps, has_chroma = write_aps_alf(f % 32, rng)
                alf = (f % 32, int(rng.integers(0, 2)), int(rng.integers(0, 4)) if has_chroma else 0)
            common = dict(nut=X.NUT_IDR if idr else X.NUT_NONIDR, slice_type=st, deblock=deblock, alpha=int(rng.integers(-3, 4)), beta=int(rng.integers(-3, 4)),
                          qp_u_off=int(rng.integers(-3, 4)), qp_v_off=int(rng.integers(-3, 4)), alf=alf, tiles=tiles)
            for k, tr in enumerate(slices):
                sh = dict(common, qp=int(np.clip(qp + rng.integers(-4, 5), 0, 51)), tile_range=tr)
                units.append(dict(f=f, first=k == 0, aps=aps if k == 0 else None, sh=sh, hdr=write_sh(tools, **sh),
                                  n_tiles=tiles_in_slice(tiles, *tr) if tiles else 1))
        accepted = []                      # per slice: (slice data bytes, bins, tile sizes)

        def run(upto, attempt):
            """fresh decoder: replay slices [0, upto), then draw slice `upto` (None: replay only).  Returns (pictures, result)"""
            pics, result = [], None
            with X.Decoder(self.lib) as d:
                for n in params:
                    ret, _ = d.decode(n)
                    assert ret >= 0, ("parameter set rejected", ret)
                for u in range(upto + (0 if attempt is None else 1)):
                    unit = units[u]
                    replay = u < len(accepted) and (attempt is None or u < upto)
                    L.gen_reset(int(seed) * 1000003 + u * 101 + (attempt or 0), lps_scale, ep_one)
                    if replay:
                        bins = accepted[u][1]
                        L.gen_replay(bins.ctypes.data, bins.size)
                    if unit["aps"] is not None:
                        ret, _ = d.decode(unit["aps"])
                        assert ret >= 0, ("adaptation parameter set rejected", u, ret)
                    ret, stat = d.decode(unit["hdr"] + tail)
                    assert ret >= 0, ("generator decode failed", u, ret)
                    why = C.create_string_buffer(200)
                    invalid = L.gen_invalid(why, 200)
                    if replay:
                        assert L.gen_replay_done() and not invalid, "replay diverged"
                    else:
                        buf = (C.c_ubyte * len(tail))()
                        n = L.gen_take(buf, len(tail))
                        assert n > 0, "slice data not terminated"
                        bad = L.gen_selfcheck()
                        assert bad < 0, f"arithmetic encoder self-check: bin {bad} of slice {u} does not decode as chosen"
                        lb = (C.c_int32 * (2 * 4000000))()
                        k = L.gen_log(lb, 2 * 4000000)
                        assert k <= 4000000
                        bins = np.frombuffer(lb, np.int32, 2 * k)[1::2].astype(np.uint8)
                        sizes = (C.c_int32 * 512)()
                        nt = L.gen_tile_sizes(sizes, 512)
                        assert nt == unit["n_tiles"] and sum(sizes[:nt]) == n, ("tiles generated", nt, list(sizes[:nt]), n)
                        result = (bytes(buf[:n]), bins, why.value.decode() if invalid else None, list(sizes[:nt]))
                    while True:
                        p = d.pull()
                        if p is None:
                            break
                        pics.append(p)
                while True:
                    p = d.pull()
                    if p is None:
                        break
                    pics.append(p)
            return pics, result

        for u in range(len(units)):
            for attempt in range(tries):
                _, (data, bins, why, sizes) = run(u, attempt)
                if why is None:
                    accepted.append((data, bins, sizes))
                    break
            else:
                raise NonConforming(f"slice {u} (picture {units[u]['f']}): no conforming slice in {tries} draws (last: {why})")
        pics, _ = run(len(units), None)
        out = list(params)
        for u, unit in enumerate(units):
            if unit["aps"] is not None:
                out.append(unit["aps"])
            hdr = write_sh(tools, entry=accepted[u][2], **unit["sh"]) if unit["n_tiles"] > 1 else unit["hdr"]     # the entry points are known now
            assert len(hdr) == len(unit["hdr"])
            out.append(hdr + accepted[u][0])
        return out, pics


def same_pictures(a, b):
    return len(a) == len(b) and all(np.array_equal(x, y) for pa, pb in zip(a, b) for x, y in zip(pa, pb))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--profile", default="baseline", choices=["baseline", "main"])
    ap.add_argument("--w", type=int, default=64)
    ap.add_argument("--h", type=int, default=64)
    ap.add_argument("--bd", type=int, default=8)
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--types", default="IPB")
    ap.add_argument("--qp", type=int, default=30)
    ap.add_argument("--lps-scale", type=int, default=256)
    ap.add_argument("--set", action="append", default=[], help="tool=0/1 overrides, e.g. --set dmvr=0")
    args = ap.parse_args()
    tools = dict(BASELINE if args.profile == "baseline" else MAIN)
    for kv in args.set:
        k, v = kv.split("=")
        tools[k] = int(v)
    g = Generator()
    nals, own = g.make(tools, args.w, args.h, args.bd, args.frames, args.seed, args.types, args.qp, args.lps_scale)
    ref = X.decode_stream(X.XevdLibrary(X.REF_SO), nals)
    ok = same_pictures(own, ref)
    X.write_stream(args.out, nals)
    print(f"{args.out}: {len(nals)} NAL units, {sum(map(len, nals))} bytes, {len(ref)} pictures; unmodified reference reproduces the generator's pictures: {ok}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
