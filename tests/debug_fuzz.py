#!/usr/bin/env python3
"""Debug helper: re-run tests/test_gpu_parity.py::test_fuzzed_blocks_match_oracle and print every mismatch."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tools"), str(ROOT / "tests")]
import ffmpeg_ht_b200 as fb
import j2kgen
from helpers import ht_block, oracle

npasses = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rng = np.random.default_rng(900 + npasses)
blocks = []
while len(blocks) < 400:
    w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
    if rng.random() < 0.4:
        w, h = [(64, 64), (32, 32), (64, 32), (128, 16)][int(rng.integers(0, 4))]
    bits = int(rng.integers(2, 13))
    coef = (rng.standard_normal((h, w)) * rng.uniform(0.5, 2 ** bits)).astype(np.int32)
    if rng.random() < 0.4:
        coef[rng.random((h, w)) < rng.uniform(0.5, 0.97)] = 0
    refine = npasses > 1
    if not np.any(np.abs(coef) >> refine):
        continue
    M_b = bits + 7
    zbp = M_b - 1 - refine
    vsc = int(rng.integers(0, 2))
    data, Lcup, Lref = j2kgen.encode_block(coef, npasses, bool(vsc), zbp + 1)
    d = bytearray(data)
    for _ in range(int(rng.integers(1, 6))):
        kind = rng.random()
        if kind < 0.5:
            d[int(rng.integers(0, len(d)))] ^= 1 << int(rng.integers(0, 8))
        elif kind < 0.8:
            pos = int(rng.integers(0, len(d)))
            d[pos] = int(rng.choice([0x00, 0xFF, 0x7F, 0x8F, 0x90]))
        else:
            pos = int(rng.integers(0, max(1, len(d) - 4)))
            d[pos:pos + 4] = rng.integers(0, 256, len(d[pos:pos + 4]), dtype=np.uint8).tobytes()
    if rng.random() < 0.15:
        zbp = max(0, zbp - int(rng.integers(1, 4)))
    blocks.append(dict(data=bytes(d), npasses=npasses, Lcup=Lcup, Lref=Lref, zbp=zbp, w=w, h=h, M_b=M_b, roi=0, vsc=vsc))
got, status = fb.ht_blocks(blocks)
bad = 0
for i, b in enumerate(blocks):
    r, want = ht_block(oracle().oracle_ht_block, b["data"], b["npasses"], b["Lcup"], b["Lref"], b["zbp"], b["w"],
                       b["h"], b["M_b"], 0, b["vsc"])
    if (r > 0) != (status[i] > 0) or not np.array_equal(got[i], want):
        bad += 1
        diff = np.argwhere(got[i] != want)
        print(f"block {i}: w={b['w']} h={b['h']} np={b['npasses']} Lcup={b['Lcup']} Lref={b['Lref']} zbp={b['zbp']} M_b={b['M_b']} "
              f"vsc={b['vsc']} r={r} status={status[i]} ndiff={len(diff)} first={diff[:6].tolist()}")
        if len(diff):
            y, x = diff[0]
            for y, x in diff[:12]:
                print(f"   ({y},{x}) got {int(got[i][y, x]) & 0xFFFFFFFF:08x} want {int(want[y, x]) & 0xFFFFFFFF:08x}")
            nzg, nzw = np.argwhere(got[i] != 0), np.argwhere(want != 0)
            print("   nonzero got", len(nzg), "want", len(nzw), "first rows got", sorted(set(nzg[:, 0].tolist()))[:10], "want", sorted(set(nzw[:, 0].tolist()))[:10])
        print("   data", b["data"].hex())
print("mismatches", bad, "of", len(blocks))
