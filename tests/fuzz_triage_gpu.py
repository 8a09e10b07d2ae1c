#!/usr/bin/env python3
"""Triage of tests/fuzz_gpu.py findings on the GPU box: for every saved case, the HT code-blocks of the product's
plan go through the stage-level entry (b200j2k_ht_blocks: same kernels as a frame) and through oracle_ht_block; the
blocks that differ are printed and saved (gpurun_out/fuzz_blocks.npz) for a closer look on the CPU.
  python tests/fuzz_triage_gpu.py gpurun_out/fuzzgpu_*.j2k"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tools"), str(ROOT / "tests")]
import numpy as np  # noqa: E402

import ffmpeg_ht_b200 as fb  # noqa: E402
from helpers import ht_block, oracle  # noqa: E402

saved = {}
classes = {}
for f in sys.argv[1:]:
    cs = Path(f).read_bytes()
    try:
        P = fb.plan(cs, 0)
    except fb.DecodeError as e:
        print(f, "plan fails", e.code)
        continue
    data = np.concatenate([np.frombuffer(cs, np.uint8), P["extra"], np.zeros(64, np.uint8)])
    blocks, idx = [], []
    n_mq = 0
    for i, d in enumerate(P["cblk"]):
        fl = int(d["flags"])
        if fl & 0x80:
            continue
        if fl & 0x08:
            n_mq += 1
            continue
        seg = data[int(d["data_off"]): int(d["data_off"]) + int(d["length"])].tobytes()
        blocks.append(dict(data=seg, npasses=int(d["npasses"]), Lcup=int(d["Lcup"]), Lref=int(d["Lref"]), zbp=int(d["zbp"]),
                           w=int(d["w"]), h=int(d["h"]), M_b=int(d["M_b"]), roi=int(d["roi_shift"]), vsc=fl & 1))
        idx.append(i)
    got, status = fb.ht_blocks(blocks)
    nbad = 0
    for b, g, s, i in zip(blocks, got, status, idx):
        r, want = ht_block(oracle().oracle_ht_block, b["data"], b["npasses"], b["Lcup"], b["Lref"], b["zbp"], b["w"], b["h"],
                           b["M_b"], b["roi"], b["vsc"])
        if r < 0:
            want = np.zeros_like(want)
        same_verdict = (r > 0) == (s > 0) if r >= 0 else s < 0
        if not np.array_equal(g, want) or not same_verdict:
            nbad += 1
            key = f"{Path(f).stem}_{i}"
            saved[key + "|data"] = np.frombuffer(b["data"], np.uint8)
            saved[key + "|par"] = np.array([b["npasses"], b["Lcup"], b["Lref"], b["zbp"], b["w"], b["h"], b["M_b"], b["roi"],
                                            b["vsc"], r, int(s)], np.int32)
            saved[key + "|gpu"] = g
            kind = (f"oracle {'err' if r < 0 else r} gpu {'err' if s < 0 else int(s)} np{b['npasses'] % 3 or 3}"
                    f"{' placeholders' if b['npasses'] > 3 else ''}")
            classes[kind] = classes.get(kind, 0) + 1
            if nbad <= 2:
                print(f"{Path(f).name} block {i}: {b['w']}x{b['h']} np {b['npasses']} Lcup {b['Lcup']} Lref {b['Lref']} len "
                      f"{len(b['data'])} zbp {b['zbp']} M_b {b['M_b']}: oracle ret {r}, gpu status {int(s)}, "
                      f"{int((g != want).sum())} samples differ", flush=True)
    if not nbad:
        print(f"{Path(f).name}: all {len(blocks)} HT blocks agree ({n_mq} Part-1 blocks not looked at)", flush=True)
print(classes)
np.savez_compressed(ROOT / "gpurun_out" / "fuzz_blocks.npz", **saved)
