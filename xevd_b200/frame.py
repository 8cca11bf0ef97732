"""Host-side data formats either side of the reconstruction path.

* HostPicture  -- a picture in the reference's padded layout (xevd_imgb_create / xevd_picbuf_lc_alloc,
                  src_base/xevd_util.c:153-230,250-363): pad 144 (luma) / 72 (chroma) on every side,
                  16-bit samples, per-SCU maps map_mv / map_refi / map_scu.
* CuList       -- the flat CU work-item array (XB200_CU, include/xevd_b200.h) in decoding order plus the
                  per-CTU first-CU index and the packed coefficient stream, i.e. what xevd_recon_unit
                  would read from XEVD_CU_DATA (xevd_def.h:1145-1190) after motion derivation.

Pure layout code (numpy); no pixel arithmetic happens here.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .abi import CU_DTYPE, EXT_DTYPE

PIC_PAD_SIZE_L = 144  # MAX_CU_SIZE + 16 (xevd_def.h:211)
PIC_PAD_SIZE_C = 72
MIN_CU_LOG2 = 2


class HostPicture:
    """16-bit 4:2:0 picture with replicated-border storage, reference layout."""

    def __init__(self, w: int, h: int, poc: int = 0, pad_l: int = PIC_PAD_SIZE_L, pad_c: int = PIC_PAD_SIZE_C):
        if w % 8 or h % 8:
            raise ValueError("picture dimensions must be multiples of 8 (EVC minimum CU size)")
        self.w, self.h, self.poc = w, h, poc
        self.pad_l, self.pad_c = pad_l, pad_c
        self.w_c, self.h_c = w // 2, h // 2
        self.buf_y = np.zeros((h + 2 * pad_l, w + 2 * pad_l), np.int16)
        self.buf_u = np.zeros((self.h_c + 2 * pad_c, self.w_c + 2 * pad_c), np.int16)
        self.buf_v = np.zeros_like(self.buf_u)
        self.w_scu, self.h_scu = (w + 3) >> 2, (h + 3) >> 2
        n = self.w_scu * self.h_scu
        self.map_mv = np.zeros((n, 2, 2), np.int16)
        self.map_unrefined_mv = np.zeros((n, 2, 2), np.int16)     # Main: vectors before DMVR refinement (map_mv holds the refined ones)
        self.map_refi = np.full((n, 2), -1, np.int8)
        self.map_scu = np.zeros(n, np.uint32)

    # interior views (sample (0,0) first)
    @property
    def y(self):
        p = self.pad_l
        return self.buf_y[p:p + self.h, p:p + self.w]

    @property
    def u(self):
        p = self.pad_c
        return self.buf_u[p:p + self.h_c, p:p + self.w_c]

    @property
    def v(self):
        p = self.pad_c
        return self.buf_v[p:p + self.h_c, p:p + self.w_c]

    @property
    def s_l(self):
        return self.buf_y.shape[1]

    @property
    def s_c(self):
        return self.buf_u.shape[1]

    def planes(self):
        return self.y, self.u, self.v

    def copy(self):
        o = HostPicture(self.w, self.h, self.poc, self.pad_l, self.pad_c)
        o.buf_y[...] = self.buf_y
        o.buf_u[...] = self.buf_u
        o.buf_v[...] = self.buf_v
        o.map_mv[...] = self.map_mv
        o.map_unrefined_mv[...] = self.map_unrefined_mv
        o.map_refi[...] = self.map_refi
        o.map_scu[...] = self.map_scu
        return o

    @staticmethod
    def random(w, h, bit_depth, rng, poc=0):
        """uniform random u<bit_depth> content, borders NOT yet replicated"""
        p = HostPicture(w, h, poc)
        hi = 1 << bit_depth
        p.y[...] = rng.integers(0, hi, (h, w), dtype=np.int16)
        p.u[...] = rng.integers(0, hi, (h // 2, w // 2), dtype=np.int16)
        p.v[...] = rng.integers(0, hi, (h // 2, w // 2), dtype=np.int16)
        return p

    def pad_borders(self):
        """host-side layout helper equal in effect to picbuf_expand (used only to prepare INPUT
        reference pictures of synthetic workloads; the decoded-picture padding is a device kernel)"""
        for buf, pad, w, h in ((self.buf_y, self.pad_l, self.w, self.h),
                               (self.buf_u, self.pad_c, self.w_c, self.h_c),
                               (self.buf_v, self.pad_c, self.w_c, self.h_c)):
            buf[pad:pad + h, :pad] = buf[pad:pad + h, pad:pad + 1]
            buf[pad:pad + h, pad + w:] = buf[pad:pad + h, pad + w - 1:pad + w]
            buf[:pad, :] = buf[pad:pad + 1, :]
            buf[pad + h:, :] = buf[pad + h - 1:pad + h, :]
        return self


def ats_inter_tu(ats: int, log2w: int, log2h: int):
    """(log2 tu_w, log2 tu_h, x offset, y offset) of the sub-block transform unit of an ats_inter CU
    (xevdm_get_tu_size / get_tu_pos_offset, src_main/xevdm_util.c:3585-3634); the whole CU when ats_inter_idx == 0"""
    idx, pos = (ats >> 2) & 7, (ats >> 5) & 1
    if idx == 0:
        return log2w, log2h, 0, 0
    sh = 2 if idx in (3, 4) else 1
    if idx in (2, 4):
        return log2w, log2h - sh, 0, ((1 << log2h) - (1 << (log2h - sh))) if pos else 0
    return log2w - sh, log2h, ((1 << log2w) - (1 << (log2w - sh))) if pos else 0, 0


def sparse_coef(coef: np.ndarray, chunk: int = 4096):
    """the sparse form of a dense coefficient stream that xb200_recon_frame_sparse takes: (entries, chunk_first) with
    entries[i] = position inside its chunk | level << 16 and chunk k owning entries[chunk_first[k] : chunk_first[k + 1]]"""
    coef = np.ascontiguousarray(coef, np.int16)
    idx = np.flatnonzero(coef)
    entries = ((idx & (chunk - 1)).astype(np.uint32) | (coef[idx].view(np.uint16).astype(np.uint32) << 16)).astype(np.uint32)
    n_chunks = (coef.size + chunk - 1) // chunk
    chunk_first = np.searchsorted(idx >> 12 if chunk == 4096 else idx // chunk, np.arange(n_chunks + 1), side="left").astype(np.uint32)
    return entries, chunk_first


def cu_coef_count(cu) -> int:
    """number of int16 coefficients a CU contributes to the stream (planes with cbf == 0 are absent)"""
    lw, lh = ats_inter_tu(int(cu["ats"]), int(cu["log2w"]), int(cu["log2h"]))[:2] if int(cu["mode"]) != 0 else (int(cu["log2w"]), int(cu["log2h"]))
    n = 1 << (lw + lh)
    cbf = int(cu["cbf"])
    a8 = lambda v: (v + 7) & ~7   # plane blocks are padded to multiples of 8 int16
    return (a8(n) if cbf & 0x00F else 0) + (a8(n // 4) if cbf & 0x0F0 else 0) + (a8(n // 4) if cbf & 0xF00 else 0)


@dataclass
class CuList:
    """Flat per-picture work description handed to xb200_recon_frame."""

    w: int
    h: int
    log2_ctu: int
    cus: np.ndarray                      # CU_DTYPE, decoding order
    ctu_first: np.ndarray                # uint32 [n_ctu + 1]
    coef: np.ndarray                     # int16 stream
    ext: np.ndarray = field(default_factory=lambda: np.zeros(1, EXT_DTYPE))

    @property
    def n_cu(self):
        return len(self.cus)

    @property
    def n_ctu(self):
        return len(self.ctu_first) - 1

    def validate(self):
        c = self.cus
        assert c.dtype == CU_DTYPE and self.coef.dtype == np.int16 and self.ctu_first.dtype == np.uint32
        assert self.ctu_first[0] == 0 and self.ctu_first[-1] == len(c)
        assert np.all(np.diff(self.ctu_first.astype(np.int64)) >= 0)
        x1 = c["x"].astype(np.int64) + (1 << c["log2w"].astype(np.int64))
        y1 = c["y"].astype(np.int64) + (1 << c["log2h"].astype(np.int64))
        assert np.all(x1 <= self.w) and np.all(y1 <= self.h), "CU outside the picture"
        ctu = 1 << self.log2_ctu
        wc = (self.w + ctu - 1) // ctu
        idx = (c["y"] // ctu).astype(np.int64) * wc + c["x"] // ctu
        owner = np.repeat(np.arange(self.n_ctu), np.diff(self.ctu_first.astype(np.int64)))
        assert np.array_equal(idx, owner), "CUs not grouped by CTU in raster order"
        sizes = np.array([cu_coef_count(cu) for cu in c], np.int64)
        run = np.concatenate(([0], np.cumsum(sizes)))
        assert np.array_equal(c["coef_off"].astype(np.int64), run[:-1]), "coefficient stream not in decoding order"
        assert run[-1] == self.coef.size
        return self

    def band(self, ctu_row0: int, ctu_rows: int) -> "CuList":
        """the work of CTU rows [ctu_row0, ctu_row0 + ctu_rows) as a self-contained list (band mode of xb200_recon_frame): CUs are in
        CTU-raster order, so the band is one contiguous slice of the CU array and of the coefficient stream; offsets are rebased"""
        ctu = 1 << self.log2_ctu
        w_ctu = (self.w + ctu - 1) // ctu
        c0, c1 = int(self.ctu_first[ctu_row0 * w_ctu]), int(self.ctu_first[(ctu_row0 + ctu_rows) * w_ctu])
        cus = self.cus[c0:c1].copy()
        k0 = int(self.cus["coef_off"][c0]) if c0 < len(self.cus) else self.coef.size
        k1 = int(self.cus["coef_off"][c1]) if c1 < len(self.cus) else self.coef.size
        cus["coef_off"] -= k0
        first = (self.ctu_first[ctu_row0 * w_ctu:(ctu_row0 + ctu_rows) * w_ctu + 1].astype(np.int64) - c0).astype(np.uint32)
        return CuList(w=self.w, h=self.h, log2_ctu=self.log2_ctu, cus=cus, ctu_first=first, coef=self.coef[k0:k1].copy(), ext=self.ext)

    def edge_flags(self) -> np.ndarray:
        """one byte per SCU: bit0 = CU/TU boundary on the left side, bit1 = on the top side, bit2 = ats_inter CU, bits 3/4 = the
        left / top boundary is a luma edge only (inner leaf boundaries of a local dual tree node)
        (what deblock_tree derives from map_split, src_base/xevd.c:1057-1114: CU boundaries plus the
        64-sample transform split of larger CUs)"""
        w_scu, h_scu = (self.w + 3) >> 2, (self.h + 3) >> 2
        f = np.zeros((h_scu, w_scu), np.uint8)
        for cu in self.cus:
            x0, y0 = int(cu["x"]) >> 2, int(cu["y"]) >> 2
            nw, nh = 1 << (int(cu["log2w"]) - 2), 1 << (int(cu["log2h"]) - 2)
            fl = int(cu["flags"]) & 3
            if fl == 2:
                # chroma-only CU of a local dual tree node (visited after its luma leaves): its outline carries chroma again
                f[y0:y0 + nh, x0] &= ~np.uint8(0x08)
                f[y0, x0:x0 + nw] &= ~np.uint8(0x10)
                continue
            # luma-only leaves: XB200_EDGE_LEFT_NOC / XB200_EDGE_TOP_NOC = luma edge only
            for xs in range(x0, x0 + nw, 16):
                f[y0:y0 + nh, xs] |= 1 | (0x08 if fl == 1 else 0)
            for ys in range(y0, y0 + nh, 16):
                f[ys, x0:x0 + nw] |= 2 | (0x10 if fl == 1 else 0)
            if int(cu["mode"]) not in (0, 4) and (int(cu["ats"]) >> 2) & 7:
                f[y0:y0 + nh, x0:x0 + nw] |= 4          # XB200_EDGE_ATS: SCUs of ats_inter CUs (map_ats_inter != 0)
        return f.reshape(-1)
