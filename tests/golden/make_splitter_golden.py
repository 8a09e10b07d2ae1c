#!/usr/bin/env python3
"""Generates tests/golden/splitter.npz: call traces of the REFERENCE parser (oracle/_ref/libj2kref.so, built from
/root/reference by oracle/Makefile) on chunked synthetic streams.  Run where the reference is mounted:
    python tests/golden/make_splitter_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "tools"), str(ROOT / "tests")]
import test_splitter as ts  # noqa: E402

L = ts.ref_lib()
out = {}
cases = [(0, 0), (1, 1), (2, 2), (3, 3), (4, 3), (5, 0)]
for k, (seed, c) in enumerate(cases):
    blob, sizes, _ = ts.make_case(seed, c)
    h, ref = ts.ref_parse_fn(L)
    tr = ts.drive(ref, blob, sizes)
    L.ref_parser_free(h)
    out[f"blob{k}"] = np.frombuffer(blob, np.uint8)
    out[f"sizes{k}"] = np.array(sizes, np.int32)
    out[f"ret{k}"] = np.array([t[0] for t in tr], np.int32)
    out[f"len{k}"] = np.array([-1 if t[1] is None else t[1][0] for t in tr], np.int64)
    out[f"sha{k}"] = np.array(["" if t[1] is None else t[1][1] for t in tr])
out["ncases"] = np.array(len(cases))
np.savez_compressed(ROOT / "tests" / "golden" / "splitter.npz", **out)
print("wrote", ROOT / "tests" / "golden" / "splitter.npz")
