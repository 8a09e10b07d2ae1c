#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libj2kref.so, built by
`make -C oracle ref` where /root/reference is mounted).  The fixtures travel to boxes that do not have
the reference; tests/test_golden.py checks the oracle (CPU) and the CUDA path (GPU) against them.

  python tests/golden/make_golden.py
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tools"))
import j2kgen  # noqa: E402
from helpers import ht_block, ref, ref_decode  # noqa: E402

OUT = Path(__file__).resolve().parent

# hand-built known-answer codestreams (hex), confirmed against the reference
PREFIX = ("ff4fff510029400000000040000000400000000000000000000000400000004000000000000000000001070101"
          "ff500008000200000000ff52000c00000001000004044001ff5c00044040")
KATS = {
    "K1": PREFIX + "ff90000a00000000001a0001ff93c154017fff7fff7fff006900ffd9",
    "K2": PREFIX + "ff90000a00000000001d0001ff93c1ca84017fff7fff7fff0069000001ffd9",
    "K3": PREFIX + "ff90000a00000000001d0001ff93c1ca84017fff7fff7fff0069002101ffd9",
    "K4": PREFIX + "ff90000a00000000001a0001ff93c154017fff7fff7fff477900ffd9",
    "K5": "ff4fff510029400000000040000000400000000000000000000000400000004000000000000000000001070101ff500008000200000020ff52000c00000001000104044000ff5c000b424000406440c8412cff90000a00000000001b0001ff93c154017fff7fff7fff00690000ffd9",
    "K6": "ff4fff510029400000000040000000400000000000000000000000400000004000000000000000000001070101ff500008000200000020ff52000c00000001000104044000ff5c000b424000406440c8412cff90000a00000000001c0001ff9300905500017fff7fff7fff006900ffd9",
    "K7": "ff4fff51002f4000000000400000004000000000000000000000004000000040000000000000000000030b01010b01010b0101ff500008000200000000ff52000c00000001010104044001ff5c00074068686868ff90000a00000000001f0001ff9300c154017fff7fff7fff00690000000000ffd9",
    "K8": "ff4fff510029400000000080000000400000000000000000000000400000004000000000000000000001070101ff500008000200000000ff52000c00000001000004044001ff5c00044040ff90000a00000000000f0001ff9300ff90000a00010000001a0001ff93c154017fff7fff7fff006900ffd9",
}

STREAMS = [
    ("gray8_lv0", dict(w=64, h=64, nc=1, bd=8, levels=0)),
    ("rgb12_53_tiles_off", dict(w=120, h=75, nc=3, bd=12, levels=4, offset=(13, 7), tile=(64, 48))),
    ("rgb10_97", dict(w=128, h=96, nc=3, bd=10, levels=5, reversible=False, base_step=2.0)),
    ("gray16_53", dict(w=150, h=100, nc=1, bd=16, levels=5)),
    ("rgb12_rpcl_prec_3pass", dict(w=128, h=128, nc=3, bd=12, levels=3, cblk=(32, 32), precincts=[(6, 6), (7, 7)],
                                   prog=2, npasses=3)),
    ("rgb8_97_cprl_sop", dict(w=96, h=64, nc=3, bd=8, levels=3, cblk=(16, 16), precincts=[(5, 5), (6, 6)], prog=4,
                              sop=True, eph=True, reversible=False, base_step=0.5)),
    ("ya8_vsc_2pass", dict(w=70, h=50, nc=2, bd=8, levels=2, cblk=(16, 16), mct=False, npasses=2, vsc=True)),
    ("rgba12", dict(w=64, h=40, nc=4, bd=12, levels=2, cblk=(16, 16))),
]


def main():
    # ---- whole codestreams -----------------------------------------------------------------------
    frames = {}
    for name, hexs in KATS.items():
        cs = bytes.fromhex(hexs)
        for be in (0, 1):
            r, info, out, pix = ref_decode(cs, be)
            assert r == 0, name
            frames[f"{name}|{be}"] = (cs, info[0], info[1], pix, out)
    for name, cfg in STREAMS:
        cfg = dict(cfg)
        w, h, nc, bd = cfg.pop("w"), cfg.pop("h"), cfg.pop("nc"), cfg.pop("bd")
        planes = j2kgen.synth_image(w, h, nc, bd, seed=21)
        p = j2kgen.EncParams(**cfg)
        cs = j2kgen.encode(planes, bd, p)
        for be in ((0,) if p.reversible else (0, 1)):
            r, info, out, pix = ref_decode(cs, be)
            assert r == 0, name
            frames[f"{name}|{be}"] = (cs, info[0], info[1], pix, out)
    pack = {}
    for k, (cs, w, h, pix, out) in frames.items():
        pack[k + "|cs"] = np.frombuffer(cs, np.uint8)
        pack[k + "|meta"] = np.array([w, h], np.int32)
        pack[k + "|pix"] = np.frombuffer(pix.encode(), np.uint8)
        pack[k + "|out"] = out
    np.savez_compressed(OUT / "frames.npz", **pack)

    # ---- the widened surface: lowres, JP2 / JPH wrapper, chroma subsampling (planar output) ----------------
    # key = name|flags with flags = bitexact | lowres << 8 (what ref_decode / oracle_decode take)
    from test_oracle_vs_ref import SUBSAMPLED, subsampled_stream
    wide = {}
    planes = j2kgen.synth_image(200, 120, 3, 12, seed=31)
    cs = j2kgen.encode(planes, 12, j2kgen.EncParams(levels=4, tile=(128, 64)))
    cs97 = j2kgen.encode(j2kgen.synth_image(200, 120, 3, 10, seed=32), 10,
                         j2kgen.EncParams(levels=3, reversible=False, base_step=1.0))
    for lowres in (1, 2, 4):
        wide[f"lowres{lowres}_rgb12_tiles|{lowres << 8}"] = cs
    for lowres in (1, 3):
        for be in (0, 1):
            wide[f"lowres{lowres}_rgb10_97|{be | lowres << 8}"] = cs97
    for name, subs, bd in SUBSAMPLED:
        c, _, _ = subsampled_stream(subs, bd, 131, 77, dict(levels=3, tile=(64, 48)))
        wide[f"sub_{name}|0"] = c
        wide[f"sub_{name}_jp2_cdef|0"] = j2kgen.wrap_jp2(c, 131, 77, len(subs), bd, colour_space=18,
                                                         cdef=[(0, 0, 1), (1, 0, 3), (2, 0, 2), (3, 1, 0)][:len(subs)])
    for nc, bd in [(3, 8), (3, 12), (1, 16), (4, 8)]:
        c = j2kgen.encode(j2kgen.synth_image(160, 90, nc, bd, seed=8), bd, j2kgen.EncParams(levels=3, mct=nc >= 3))
        for csp in (16, 17, 18):
            wide[f"jp2_nc{nc}_bd{bd}_colr{csp}|0"] = j2kgen.wrap_jp2(c, 160, 90, nc, bd, colour_space=csp)
    from test_oracle_vs_ref import LAYERED
    for i in (1, 3, 7):                                   # two quality layers: split HT sets, placeholders, late inclusion
        cfg = dict(LAYERED[i])
        w_, h_, nc_, bd_ = cfg.pop("w"), cfg.pop("h"), cfg.pop("nc"), cfg.pop("bd")
        c = j2kgen.encode(j2kgen.synth_image(w_, h_, nc_, bd_, seed=4), bd_, j2kgen.EncParams(layers=2, **cfg))
        for be in ((0,) if cfg.get("reversible", True) else (0, 1)):
            wide[f"layers2_{i}|{be}"] = c
    pack = {}
    for k, data in wide.items():
        r, info, out, pix = ref_decode(data, int(k.split("|")[1]))
        pack[k + "|cs"] = np.frombuffer(data, np.uint8)
        pack[k + "|meta"] = np.array([info[0], info[1], r], np.int64)
        pack[k + "|pix"] = np.frombuffer((pix or "none").encode(), np.uint8)
        pack[k + "|out"] = out if r == 0 else np.zeros(0, np.uint8)
    np.savez_compressed(OUT / "frames_wide.npz", **pack)

    # ---- HT code-blocks ----------------------------------------------------------------------------
    rng = np.random.default_rng(2024)
    blocks = {}
    i = 0
    while i < 120:
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        npasses = int(rng.choice([1, 1, 2, 3]))
        bits = int(rng.integers(1, 13))
        coef = (rng.standard_normal((h, w)) * rng.uniform(0.3, 2 ** bits)).astype(np.int32)
        refine = npasses > 1
        if not np.any(np.abs(coef) >> refine):
            continue
        M_b = bits + 6
        zbp = M_b - 1 - refine
        vsc = int(rng.integers(0, 2))
        data, Lcup, Lref = j2kgen.encode_block(coef, npasses, bool(vsc), zbp + 1)
        if i % 10 == 9:   # a few damaged ones: the reference's error => zeros behaviour
            d = bytearray(data)
            d[int(rng.integers(0, len(d)))] ^= 0x5A
            data = bytes(d)
        ret, out = ht_block(ref().ref_ht_block, data, npasses, Lcup, Lref, zbp, w, h, M_b, 0, vsc)
        blocks[f"{i}|data"] = np.frombuffer(data, np.uint8)
        blocks[f"{i}|par"] = np.array([npasses, Lcup, Lref, zbp, w, h, M_b, 0, vsc, ret], np.int64)
        blocks[f"{i}|out"] = out
        i += 1
    np.savez_compressed(OUT / "blocks.npz", **blocks)

    # ---- the reference's own DWT self-test output (tests/ref/fate/j2k-dwt) is reproduced by
    # libavcodec/tests/jpeg2000dwt.c; here: IDWT vectors on fixed seeds -----------------------------
    dwt = {}
    for t in (0, 1, 2):
        for j in range(12):
            x0, y0 = int(rng.integers(0, 9)), int(rng.integers(0, 9))
            w, h = int(rng.integers(1, 40)), int(rng.integers(1, 40))
            levels = int(rng.integers(0, 6))
            border = np.array([x0, x0 + w, y0, y0 + h], np.int32)
            a = (rng.standard_normal(w * h) * 500).astype(np.float32) if t == 0 else \
                rng.integers(-(1 << 18), 1 << 18, w * h).astype(np.int32)
            b = a.copy()
            assert ref().ref_idwt(t, border.ctypes.data, levels, b.ctypes.data) == 0
            dwt[f"{t}|{j}|in"] = a
            dwt[f"{t}|{j}|par"] = np.array([x0, x0 + w, y0, y0 + h, levels], np.int32)
            dwt[f"{t}|{j}|out"] = b
    np.savez_compressed(OUT / "idwt.npz", **dwt)
    for f in ("frames.npz", "frames_wide.npz", "blocks.npz", "idwt.npz"):
        b = (OUT / f).read_bytes()
        print(f, len(b), hashlib.md5(b).hexdigest())


if __name__ == "__main__":
    main()
