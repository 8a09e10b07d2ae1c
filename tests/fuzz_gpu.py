#!/usr/bin/env python3
"""Differential fuzzing on the GPU box: the damaged streams of tests/fuzz_host.py decoded by the product through the
C ABI (kernels and all) and by the oracle; verdict and every byte of the picture must agree.  Reversible streams as
they are, irreversible ones with the bit-exact flag (integer 9/7) so that the comparison is exact.  Test
infrastructure; cases that differ are written to gpurun_out/fuzzgpu_<seed>_<n>.j2k.

  python tests/fuzz_gpu.py [--seconds 120] [--seed 1] [--families ...]"""
import argparse
import importlib.util
import random
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tools"), str(ROOT / "tests")]
import numpy as np  # noqa: E402

import ffmpeg_ht_b200 as fb  # noqa: E402
from helpers import oracle_decode  # noqa: E402

spec = importlib.util.spec_from_file_location("fuzz_host", ROOT / "tests" / "fuzz_host.py")
fuzz_host = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fuzz_host)


def irreversible(cs: bytes) -> bool:
    """any COD / COC in the stream with the 9/7 wavelet (a coarse scan: good enough to pick the flag)"""
    i = cs.find(b"\xff\x52")
    return i >= 0 and i + 13 < len(cs) and cs[i + 13] == 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cases", type=int, default=0, help="stop after this many cases instead of after --seconds")
    ap.add_argument("--families", default="ht_sets,layered,plain,packed,tile_parts,poc,tile_styles,jp2")
    a = ap.parse_args()
    rng = random.Random(a.seed)
    srcs = []
    for f in a.families.split(","):
        srcs += [(f, i, s) for i, s in enumerate(fuzz_host.family(f))]
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    # one decoder context per mode: (bit-exact flag, AVCodecContext.lowres)
    decs = {(be, lr): fb.Jpeg2000Decoder(device=0, slots=1, bitexact=bool(be), flags=fb.flag_lowres(lr))
            for be in (0, 1) for lr in (0, 1, 2)}
    t0 = time.time()
    n = pictures = errors = bad = failed_blocks = 0
    while (n < a.cases) if a.cases else (time.time() - t0 < a.seconds):
        fam, idx, src = rng.choice(srcs)
        cs, where = fuzz_host.damage(rng, src)
        be = 1 if irreversible(src) else int(rng.random() < 0.1)
        lr = rng.choice((0, 0, 0, 0, 1, 2))
        n += 1
        ro, oinfo, opic = oracle_decode(cs, be | (lr << 8))
        try:
            info, pic = decs[(be, lr)].decode(cs)
            rp = 0
            failed_blocks += info.nblocks_failed
        except fb.DecodeError as e:
            rp, pic = e.code, None
        if ro == 0 and rp == 0 and (info.width, info.height) != (oinfo[0], oinfo[1]):
            pic = None                                  # counts as a difference below
        if (ro == 0) != (rp == 0) or (ro == 0 and not np.array_equal(opic, pic)):
            bad += 1
            name = out / f"fuzzgpu_{a.seed}_{n}.j2k"
            name.write_bytes(cs)
            print(f"DIFFER case {n} ({fam}[{idx}], bitexact {be} lowres {lr}, damage at {where}): oracle {ro} / product {rp} -> {name.name}",
                  flush=True)
        elif ro == 0:
            pictures += 1
        else:
            errors += 1
    for d in decs.values():
        d.close()
    print(f"{n} cases in {time.time() - t0:.0f} s: {pictures} pictures equal ({failed_blocks} code-blocks the kernels "
          f"reported as failed among them), {errors} rejected by both, {bad} differ")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
