"""Shared test helpers: ctypes views of the oracle (oracle/liboracle.so), of the compiled reference
(oracle/_ref/libj2kref.so, when present) and small utilities.  Test infrastructure only."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
GOLDEN = ROOT / "tests" / "golden"

_HT_SIG = [ctypes.c_char_p] + [ctypes.c_int] * 10 + [ctypes.c_void_p]
_DEC_SIG = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_int64), ctypes.c_void_p,
            ctypes.c_int64]

ORACLE_PIX = {1: "gray", 2: "gray16le", 3: "ya8", 4: "ya16le", 5: "rgb24", 6: "rgb48le", 7: "rgba",
              8: "rgba64le", 9: "planar8", 10: "planar16"}

_oracle = None
_ref = None


def oracle():
    """The CPU restatement.  Built on demand (gcc, a second or two)."""
    global _oracle
    if _oracle is None:
        subprocess.check_call(["make", "-s", "-C", str(ORACLE_DIR), "liboracle.so"])
        lib = ctypes.CDLL(str(ORACLE_DIR / "liboracle.so"))
        lib.oracle_ht_block.argtypes = _HT_SIG
        lib.oracle_decode.argtypes = _DEC_SIG
        lib.oracle_idwt.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.oracle_mct.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.oracle_last_pix_fmt_name.restype = ctypes.c_char_p
        lib.oracle_decode_bench.restype = ctypes.c_double
        lib.oracle_decode_bench.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int]
        _oracle = lib
    return _oracle


def have_ref() -> bool:
    return (ORACLE_DIR / "_ref" / "libj2kref.so").exists()


def ref():
    """The unmodified reference, compiled by oracle/Makefile (only where it was built)."""
    global _ref
    if _ref is None:
        lib = ctypes.CDLL(str(ORACLE_DIR / "_ref" / "libj2kref.so"))
        lib.ref_ht_block.argtypes = _HT_SIG
        lib.ref_decode.argtypes = _DEC_SIG
        lib.ref_pix_fmt_name.restype = ctypes.c_char_p
        lib.ref_idwt.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.ref_mct.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.ref_decode_bench.restype = ctypes.c_double
        lib.ref_decode_bench.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int]
        _ref = lib
    return _ref


def ht_block(lib_fn, data: bytes, npasses, Lcup, Lref, zbp, w, h, M_b, roi=0, vsc=0):
    out = np.full(w * h, 0x55, np.int32)
    ret = lib_fn(data, len(data), npasses, Lcup, Lref, zbp, w, h, M_b, roi, vsc, out.ctypes.data)
    return ret, out.reshape(h, w)


def mq_block(data: bytes, npasses, nonzerobits, w, h, bandpos, M_b, roi=0, style=0):
    """oracle_mq_block_style: a Part-1 code-block (style bits 0x02 reset, 0x08 causal, 0x20 segmentation symbols) ->
    (ret, sign-magnitude words)"""
    fn = oracle().oracle_mq_block_style
    fn.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                   ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    out = np.zeros(w * h, np.int32)
    ret = fn(data, len(data), npasses, nonzerobits, w, h, bandpos, M_b, roi, style, out.ctypes.data)
    return ret, out.reshape(h, w)


def decode_with(lib_fn, cs: bytes, bitexact=0, cap=None):
    """-> (ret, info list, packed bytes or None)"""
    info = (ctypes.c_int64 * 8)()
    out = np.zeros(cap or (1 << 26), np.uint8)
    ret = lib_fn(cs, len(cs), bitexact, info, out.ctypes.data, out.size)
    return ret, list(info), (out[: info[3]].copy() if ret == 0 else None)


def oracle_decode(cs: bytes, bitexact=0, cap=None):
    return decode_with(oracle().oracle_decode, cs, bitexact, cap)


def oracle_pix_name() -> str:
    """libavutil name of the AVPixelFormat the last oracle_decode() chose"""
    return oracle().oracle_last_pix_fmt_name().decode()


def ref_decode(cs: bytes, bitexact=0, cap=None):
    ret, info, out = decode_with(ref().ref_decode, cs, bitexact, cap)
    name = ref().ref_pix_fmt_name(info[2]).decode() if ret == 0 else None
    return ret, info, out, name


def sm_to_int(block: np.ndarray, M_b: int) -> np.ndarray:
    """sign-magnitude code-block output -> integer coefficient (reversible dequantisation)."""
    mag = (block & 0x7FFFFFFF) >> (31 - M_b)
    return np.where(block < 0, -mag, mag)
