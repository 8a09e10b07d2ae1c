#!/usr/bin/env python3
"""Differential fuzzing of the host side (no GPU): byte damage in generated codestreams, decoded three ways --
  the unmodified reference decoder (oracle/_ref/libj2kref.so, when it was built here),
  the oracle (oracle/liboracle.so),
  the product's parser + planner (libb200j2k.so: b200j2k_plan) with the block coders / transforms of the oracle
  applied to ITS tables (tests/test_host_plan.py: picture_from_plan).
They must agree on the verdict (picture or error) and, for a picture, on every byte.  Test infrastructure:
prints the cases that differ and writes them to gpurun_out/fuzz_<seed>_<n>.j2k.

  python tests/fuzz_host.py [--seconds 300] [--seed 1] [--families ht_sets,layered,...]"""
import argparse
import hashlib
import random
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tools"), str(ROOT / "tests")]
import numpy as np  # noqa: E402

import ffmpeg_ht_b200 as fb  # noqa: E402
import j2kgen  # noqa: E402
from helpers import have_ref, oracle_decode, ref_decode  # noqa: E402
from test_host_plan import picture_from_plan  # noqa: E402
import test_oracle_vs_ref as T  # noqa: E402


def family(name):
    if name == "ht_sets":
        return [bytes(T.ht_set_stream(t)[0]) for t in T.HT_SETS]
    if name == "layered":
        out = []
        for cfg in T.LAYERED:
            cfg = dict(cfg)
            w, h, nc, bd = cfg.pop("w"), cfg.pop("h"), cfg.pop("nc"), cfg.pop("bd")
            out.append(bytes(j2kgen.encode(j2kgen.synth_image(w, h, nc, bd, seed=4), bd, j2kgen.EncParams(layers=2, **cfg))))
        return out
    if name == "plain":
        out = []
        for cfg in T.STREAMS:
            cfg = dict(cfg)
            w, h, nc, bd = cfg.pop("w"), cfg.pop("h"), cfg.pop("nc"), cfg.pop("bd")
            if w * h > 70000:
                continue
            out.append(bytes(j2kgen.encode(j2kgen.synth_image(w, h, nc, bd, seed=2), bd, j2kgen.EncParams(**cfg))))
        return out
    if name == "packed":
        return [bytes(T.packed_stream(t)[0]) for t in T.PACKED]
    if name == "tile_parts":
        return [bytes(T.tile_part_stream(t)[0]) for t in T.TILE_PARTS]
    if name == "poc":
        return [bytes(T.poc_stream(t)[0]) for t in T.POC_STREAMS]
    if name == "tile_styles":
        return [bytes(T.tile_style_stream(t)[0]) for t in T.TILE_STYLES]
    if name == "part1":                         # EBCOT / MQ streams from OpenJPEG (Pillow), raw and in a JP2 file
        return [bytes(T.part1_stream(t, jp2)) for t in T.PART1[:12] for jp2 in (False, True)]
    if name == "subsampled":
        out = []
        for i, (_, subs, bd) in enumerate(T.SUBSAMPLED):
            w, h, kw = T.SUBSAMPLED_GEOMS[i % len(T.SUBSAMPLED_GEOMS)]
            out.append(bytes(T.subsampled_stream(subs, bd, w, h, kw)[0]))
        return out
    if name == "jp2":                           # JP2 wrappers: palettes (PAL8), colour specifications, D-Cinema profiles,
        out = []                                # resolution boxes
        for bd, csp, pal in T.PAL8_CASES:
            out.append(bytes(T.pal8_file(bd, csp, pal)[0]))
            out.append(bytes(T.pal8_file(bd, csp, pal, tile=(64, 32))[0]))
        out += [bytes(c) for _, c in T.dcinema_cases()]
        out += [bytes(T.res_box_file(r)) for r in T.RES_BOXES]
        return out
    raise SystemExit("unknown family " + name)


def fresh_reference_sha1(cs: bytes, mode: int = 0) -> str:
    """SHA-1 of the reference's picture decoded in a process of its own ('' when it fails there)"""
    code = ("import sys,hashlib; sys.path[:0]=[%r,%r]; from helpers import ref_decode; "
            "r=ref_decode(sys.stdin.buffer.read(), %d); print(hashlib.sha1(r[2].tobytes()).hexdigest() if r[0]==0 else '')"
            % (str(ROOT / "tests"), str(ROOT), mode))
    out = subprocess.run([sys.executable, "-c", code], input=cs, capture_output=True, timeout=120)
    return out.stdout.decode().strip()


def damage(rng, src):
    cs = bytearray(src)
    where = []
    for _ in range(rng.randint(1, 3)):
        # mostly behind the main header; now and then anywhere
        lo = 0 if rng.random() < 0.1 else min(100, len(cs) // 2)
        p = rng.randrange(lo, len(cs))
        kind = rng.random()
        if kind < 0.5:
            cs[p] = rng.randrange(256)
        elif kind < 0.85:
            cs[p] ^= 1 << rng.randrange(8)
        elif kind < 0.95:
            cs[p] = 0xFF
        elif kind < 0.97:
            del cs[p: p + rng.randint(1, 40)]          # bytes lost
        elif kind < 0.985:
            del cs[p:]                                 # the rest of the frame never arrived
        else:
            n = rng.randint(2, 200)                    # a run of bytes arrives twice
            cs[p:p] = cs[p: p + n]
        where.append(p)
    return bytes(cs), where


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=300)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cases", type=int, default=0, help="stop after this many cases instead of after --seconds")
    ap.add_argument("--families", default="ht_sets,layered,plain,packed,tile_parts,poc,tile_styles")
    a = ap.parse_args()
    rng = random.Random(a.seed)
    srcs = []
    for f in a.families.split(","):
        ss = family(f)
        srcs += [(f, i, s) for i, s in enumerate(ss)]
        print(f"{f}: {len(ss)} streams", flush=True)
    use_ref = have_ref()
    print("reference decoder:", "yes" if use_ref else "not built here (oracle and product only)", flush=True)
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    t0 = time.time()
    n = pictures = errors = bad = stricter = unstable = 0
    PATCHWELCOME = -1163346256
    while (n < a.cases) if a.cases else (time.time() - t0 < a.seconds):
        fam, idx, src = rng.choice(srcs)
        cs, where = damage(rng, src)
        n += 1
        # now and then with AV_CODEC_FLAG_BITEXACT (fixed-point 9/7) and / or AVCodecContext.lowres = 1, 2
        bitexact = int(rng.random() < 0.2)
        lowres = rng.choice((0, 0, 0, 0, 1, 2))
        mode = bitexact | (lowres << 8)
        ro, _, opic = oracle_decode(cs, mode)
        pinfo = None
        try:
            pinfo, ppic = picture_from_plan(cs, (fb.FLAG_BITEXACT if bitexact else 0) | fb.flag_lowres(lowres))
            rp = 0
        except fb.DecodeError as e:
            rp, ppic = e.code, None
        what = []
        if use_ref and ro == PATCHWELCOME and rp == PATCHWELCOME:
            # the documented class (quirks): Part-1 blocks with bypass / termination on each pass, and
            # errors inside a tile job (M_b >= 31 ...) -- the reference drops the job's return value and hands out a
            # frame whose tile was never written; oracle and product refuse the frame
            if ref_decode(cs, mode)[0] == 0:
                stricter += 1
                continue
        if (ro == 0) != (rp == 0) or (ro == 0 and not np.array_equal(opic, ppic)):
            what.append(f"oracle {ro} / product {rp}")
        if use_ref:
            rr, rinfo, rpic, rname = ref_decode(cs, mode)
            if rr == 0 and ro == 0 and not np.array_equal(rpic, opic) and fresh_reference_sha1(cs, mode) != hashlib.sha1(rpic.tobytes()).hexdigest():
                # damaged lengths make the reference read heap bytes behind a block's data: what it returns then
                # depends on what this process decoded before (a fresh process is the yardstick)
                unstable += 1
                continue
            if (rr == 0) != (ro == 0) or (rr == 0 and not np.array_equal(rpic, opic)):
                what.append(f"reference {rr} / oracle {ro}")
            if (rr == 0) != (rp == 0) or (rr == 0 and rp == 0 and not np.array_equal(rpic, ppic)):
                what.append(f"reference {rr} / product {rp}")
            elif rr == 0:
                # what the caller sees besides the picture: size, format, how far the decoder read (the return value
                # of FFCodec.cb.decode)
                got = (pinfo.width, pinfo.height, pinfo.pix_name, pinfo.bytes_consumed)
                want = (rinfo[0], rinfo[1], rname, rinfo[4])
                if got != want:
                    what.append(f"frame info: reference {want} / product {got}")
        if what:
            bad += 1
            name = out / f"fuzz_{a.seed}_{n}.j2k"
            name.write_bytes(cs)
            print(f"DIFFER case {n} ({fam}[{idx}], bitexact {bitexact} lowres {lowres}, damage at {where}): {'; '.join(what)} -> {name.name}", flush=True)
        elif ro == 0:
            pictures += 1
        else:
            errors += 1
    print(f"{n} cases in {time.time() - t0:.0f} s: {pictures} pictures equal, {errors} rejected by all, "
          f"{stricter} refused where the reference hands out an unfinished frame, {unstable} where the reference "
          f"itself depends on heap contents, {bad} differ")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
