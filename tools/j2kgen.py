"""Synthetic HTJ2K workload generator (test + benchmark infrastructure, CPU only).

`synth_image()` makes the seeded content model of SURVEY.md section 8(d) (band-limited texture +
gradients + filtered noise); `encode()` turns planes into an HTJ2K codestream with tools/htj2k_enc.c
(built on demand with gcc into tools/libj2kgen.so).  `CONFIGS` names the BASELINE.json workloads.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


class _Params(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int32), ("height", ctypes.c_int32), ("ncomp", ctypes.c_int32),
        ("bitdepth", ctypes.c_int32), ("tile_w", ctypes.c_int32), ("tile_h", ctypes.c_int32),
        ("levels", ctypes.c_int32), ("cbw_log2", ctypes.c_int32), ("cbh_log2", ctypes.c_int32),
        ("reversible", ctypes.c_int32), ("mct", ctypes.c_int32), ("prog", ctypes.c_int32),
        ("use_precincts", ctypes.c_int32), ("prec_exps", ctypes.c_int32 * 33),
        ("guard_bits", ctypes.c_int32), ("npasses", ctypes.c_int32), ("sop", ctypes.c_int32),
        ("eph", ctypes.c_int32), ("vsc", ctypes.c_int32), ("base_step", ctypes.c_float),
        ("x_off", ctypes.c_int32), ("y_off", ctypes.c_int32), ("stats_blocks", ctypes.c_int32),
        ("stats_max_lcup", ctypes.c_int32), ("error", ctypes.c_int32),
        ("layers", ctypes.c_int32), ("sub_x", ctypes.c_int32 * 4), ("sub_y", ctypes.c_int32 * 4),
        ("npoc", ctypes.c_int32), ("poc", (ctypes.c_int32 * 6) * 8), ("poc_split", ctypes.c_int32),
        ("packed_headers", ctypes.c_int32), ("ppx_chunk", ctypes.c_int32),
        ("tp_packets", ctypes.c_int32), ("tp_interleave", ctypes.c_int32), ("tile_var", ctypes.c_int32),
        ("ht_set", ctypes.c_int32),
    ]


def build(force: bool = False) -> Path:
    so = _HERE / "libj2kgen.so"
    src = _HERE / "htj2k_enc.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["gcc", "-std=c11", "-O2", "-fPIC", "-shared", f"-I{_HERE.parent / 'ffmpeg-ht_b200' / 'csrc'}",
                               "-o", str(so), str(src), "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build()))
        _LIB.j2kgen_encode.restype = ctypes.c_int64
        _LIB.j2kgen_encode.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(_Params),
                                       ctypes.c_void_p, ctypes.c_int64]
        _LIB.j2kgen_encode_block.restype = ctypes.c_int
        _LIB.j2kgen_encode_block.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                             ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    return _LIB


@dataclass
class EncParams:
    levels: int = 5
    cblk: tuple = (64, 64)
    reversible: bool = True
    mct: bool = True
    prog: int = 0                 # 0 LRCP 1 RLCP 2 RPCL 3 PCRL 4 CPRL
    precincts: list | None = None  # [(PPx, PPy)] for r = 0..levels
    guard_bits: int = 2
    npasses: int = 1
    sop: bool = False
    eph: bool = False
    vsc: bool = False
    base_step: float = 1.0
    tile: tuple = (0, 0)
    offset: tuple = (0, 0)
    layers: int = 1               # 2: blocks spread over two quality layers (split HT sets, late inclusion, placeholders)
    subsampling: list | None = None   # [(XRsiz, YRsiz)] per component; plane c is ceil(W/XRsiz) x ceil(H/YRsiz)
    poc: list | None = None       # [(RSpoc, CSpoc, LYEpoc, REpoc, CEpoc, Ppoc)]: POC marker; packets follow the entries
    poc_split: int = 0            # > 0: the first poc_split entries go into a POC segment of their own
    packed_headers: int = 0       # 1: packet headers in PPT marker segments (tile-part header), 2: in PPM (main header)
    ppx_chunk: int = 0            # payload bytes per PPT / PPM marker segment (0 = 65532)
    tp_packets: int = 0           # > 0: a new tile-part after this many packets (at most 32 parts per tile)
    tp_interleave: bool = False   # write part k of every tile before part k + 1 of any
    tile_var: int = 0             # odd tiles coded differently (COD / QCD / POC in their tile-part header): bit 1 progression,
    #                               2 one level less, 4 narrower code-blocks, 8 guard bit / step, 16 tile POC
    ht_set: int = 0               # 0 HTONLY, 1 HTDECLARED, 3 MIXED (all blocks read as HT), 4 MIXED with HTONLY-style lengths
    #                               (some blocks read as Part-1 blocks: parity material), 5 like 4 under an HTONLY Ccap15
    stats: dict = field(default_factory=dict)


def encode(planes, bitdepth: int, p: EncParams) -> bytes:
    """planes: list of 2-D integer arrays (same shape), unsigned samples of `bitdepth` bits."""
    L = lib()
    planes = [np.ascontiguousarray(a, dtype=np.int32) for a in planes]
    h, w = planes[0].shape
    if p.subsampling:
        for a, (sx, sy) in zip(planes, p.subsampling):
            assert a.shape == (-(-h // sy), -(-w // sx)), "plane shape does not match its subsampling"
    guard = p.guard_bits
    while True:
        P = _Params()
        P.width, P.height, P.ncomp, P.bitdepth = w, h, len(planes), bitdepth
        P.tile_w, P.tile_h = p.tile
        P.levels = p.levels
        P.cbw_log2, P.cbh_log2 = int(np.log2(p.cblk[0])), int(np.log2(p.cblk[1]))
        P.reversible, P.mct, P.prog = int(p.reversible), int(p.mct and len(planes) >= 3), p.prog
        if p.precincts:
            P.use_precincts = 1
            for r in range(p.levels + 1):
                px, py = p.precincts[min(r, len(p.precincts) - 1)]
                P.prec_exps[r] = px | (py << 4)
        P.guard_bits, P.npasses, P.sop, P.eph, P.vsc = guard, p.npasses, int(p.sop), int(p.eph), int(p.vsc)
        P.base_step = p.base_step
        P.x_off, P.y_off = p.offset
        P.layers = p.layers
        if p.subsampling:
            for c, (sx, sy) in enumerate(p.subsampling):
                P.sub_x[c], P.sub_y[c] = sx, sy
        if p.poc:
            P.npoc = len(p.poc)
            P.poc_split = p.poc_split
            for i, e in enumerate(p.poc):
                for k in range(6):
                    P.poc[i][k] = e[k]
        P.packed_headers, P.ppx_chunk = p.packed_headers, p.ppx_chunk
        P.tp_packets, P.tp_interleave = p.tp_packets, int(p.tp_interleave)
        P.tile_var = p.tile_var
        P.ht_set = p.ht_set
        cap = int(w * h * len(planes) * 4 + (1 << 20))
        out = np.empty(cap, np.uint8)
        ptrs = (ctypes.c_void_p * len(planes))(*[a.ctypes.data for a in planes])
        n = L.j2kgen_encode(ptrs, ctypes.byref(P), out.ctypes.data, cap)
        if n == -1 and guard < 7:      # exponent bound exceeded: one more guard bit
            guard += 1
            continue
        if n < 0:
            raise RuntimeError(f"j2kgen_encode failed: {n}")
        p.stats = {"blocks": P.stats_blocks, "max_lcup": P.stats_max_lcup, "guard_bits": guard, "bytes": int(n)}
        return out[:n].tobytes()


def encode_block(coef: np.ndarray, npasses: int = 1, vsc: bool = False, maxbp: int = 31):
    """One HT code-block from signed integer coefficients -> (bytes, Lcup, Lref)."""
    L = lib()
    coef = np.ascontiguousarray(coef, dtype=np.int32)
    h, w = coef.shape
    out = np.empty(w * h * 8 + 1024, np.uint8)
    lc, lr = ctypes.c_int(), ctypes.c_int()
    n = L.j2kgen_encode_block(coef.ctypes.data, w, h, npasses, int(vsc), maxbp, out.ctypes.data, out.size,
                              ctypes.byref(lc), ctypes.byref(lr))
    if n < 0:
        raise RuntimeError(f"encode_block failed: {n}")
    return out[:n].tobytes(), lc.value, lr.value


def synth_image(w: int, h: int, ncomp: int, bitdepth: int, seed: int, phase: float = 0.0):
    """Seeded content model: sinusoid texture + gradients + 1/f-ish noise, sigma_noise ~ 2^(bitdepth-9)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    full = float((1 << bitdepth) - 1)
    planes = []
    for c in range(ncomp):
        img = np.full((h, w), 0.5, np.float32)
        for _ in range(6):
            fx, fy = rng.uniform(-0.08, 0.08, 2)
            amp = rng.uniform(0.02, 0.10)
            ph = rng.uniform(0, 2 * np.pi) + phase * rng.uniform(0.5, 2.0)
            img += amp * np.sin(2 * np.pi * (fx * xx + fy * yy) + ph).astype(np.float32)
        img += 0.15 * (xx / max(w - 1, 1) - 0.5) * rng.uniform(-1, 1) + 0.15 * (yy / max(h - 1, 1) - 0.5) * rng.uniform(-1, 1)
        # cheap 1/f-like noise: white noise at three scales, box-upsampled
        noise = np.zeros((h, w), np.float32)
        for s, a in ((1, 1.0), (4, 1.5), (16, 2.0)):
            hh, ww = -(-h // s), -(-w // s)
            n = rng.standard_normal((hh, ww)).astype(np.float32)
            noise += a * np.kron(n, np.ones((s, s), np.float32))[:h, :w]
        img += noise * (2.0 ** (bitdepth - 9)) / full / 2.0
        planes.append(np.clip(np.rint(img * full), 0, full).astype(np.int32))
    return planes


# BASELINE.json configs (index = position in `configs`, 1-based in SURVEY.md section 8d)
CONFIGS = {
    "cfg1": dict(w=2048, h=1080, ncomp=3, bitdepth=12, seed=1, enc=dict(levels=5, cblk=(64, 64), reversible=True, mct=True)),
    "cfg2": dict(w=3840, h=2160, ncomp=3, bitdepth=10, seed=2, enc=dict(levels=5, cblk=(64, 64), reversible=False, mct=True, base_step=1.0)),
    "cfg3": dict(w=7680, h=4320, ncomp=1, bitdepth=16, seed=3, enc=dict(levels=5, cblk=(64, 64), reversible=True, mct=False, tile=(3840, 2160))),
    "cfg4": dict(w=3840, h=2160, ncomp=3, bitdepth=12, seed=4, enc=dict(levels=5, cblk=(64, 64), reversible=True, mct=True)),
    "cfg5": dict(w=2048, h=1080, ncomp=3, bitdepth=12, seed=5, enc=dict(levels=5, cblk=(32, 32), reversible=True, mct=True, prog=2, npasses=3,
                                                                     precincts=[(7, 7)] + [(8, 8)] * 5)),
    "target8k": dict(w=7680, h=4320, ncomp=1, bitdepth=16, seed=6, enc=dict(levels=5, cblk=(64, 64), reversible=True, mct=False)),
}


def make_config(name: str, scale: int = 1, phase: float = 0.0, seed_offset: int = 0):
    """Returns (codestream bytes, planes, EncParams).  scale > 1 shrinks width/height for quick tests."""
    c = CONFIGS[name]
    w, h = c["w"] // scale, c["h"] // scale
    planes = synth_image(w, h, c["ncomp"], c["bitdepth"], c["seed"] + seed_offset, phase)
    enc = dict(c["enc"])
    if "tile" in enc and scale > 1:
        enc["tile"] = (enc["tile"][0] // scale, enc["tile"][1] // scale)
    p = EncParams(**enc)
    return encode(planes, c["bitdepth"], p), planes, p


def _box(kind: bytes, payload: bytes) -> bytes:
    return (8 + len(payload)).to_bytes(4, "big") + kind + payload


def wrap_jp2(cs: bytes, w: int, h: int, ncomp: int, bitdepth: int, colour_space: int | None = 16,
             cdef: list | None = None, palette: tuple | None = None, res: tuple | None = None,
             brand: bytes = b"jph ", extra_before: int = 0) -> bytes:
    """Wrap a codestream into a JP2 / JPH file (ISO/IEC 15444-1 Annex I boxes), as much of it as the
    reference's jp2_find_codestream reads: signature, ftyp, jp2h{ihdr, colr, [pclr], [cdef], [res ]}, jp2c.
    colour_space: enumerated colour space of the colr box (16 sRGB, 17 greyscale, 18 sYCC; None = no colr box).
    cdef: [(channel, type, association)].  palette: (entries, depth) -> a 3-channel pclr box.
    extra_before: number of unknown boxes placed before jp2h (the reference gives up after 10)."""
    out = _box(b"jP  ", bytes([0x0D, 0x0A, 0x87, 0x0A]))
    out += _box(b"ftyp", brand + (0).to_bytes(4, "big") + brand + b"jp2 ")
    for _ in range(extra_before):
        out += _box(b"free", b"\0" * 5)
    ihdr = h.to_bytes(4, "big") + w.to_bytes(4, "big") + ncomp.to_bytes(2, "big") + bytes([bitdepth - 1, 7, 0, 0])
    hdr = _box(b"ihdr", ihdr)
    if colour_space is not None:
        hdr += _box(b"colr", bytes([1, 0, 0]) + colour_space.to_bytes(4, "big"))
    if palette is not None:
        entries, depth = palette
        nbytes = (depth + 7) // 8
        body = entries.to_bytes(2, "big") + bytes([3, depth - 1, depth - 1, depth - 1])
        for i in range(entries):
            for c in range(3):
                body += ((i * (c + 1)) % (1 << depth)).to_bytes(nbytes, "big")
        hdr += _box(b"pclr", body)
    if cdef is not None:
        body = len(cdef).to_bytes(2, "big")
        for cn, typ, asoc in cdef:
            body += cn.to_bytes(2, "big") + typ.to_bytes(2, "big") + asoc.to_bytes(2, "big")
        hdr += _box(b"cdef", body)
    if res is not None:
        vnum, vden, hnum, hden, vexp, hexp = res
        resc = vnum.to_bytes(2, "big") + vden.to_bytes(2, "big") + hnum.to_bytes(2, "big") + hden.to_bytes(2, "big") + \
            bytes([vexp, hexp])
        hdr += _box(b"res ", _box(b"resc", resc))
    out += _box(b"jp2h", hdr)
    out += _box(b"jp2c", cs)
    return out
