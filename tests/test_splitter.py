"""Frame splitter (AVCodecParser "jpeg2000"): the product's C-ABI splitter against the
compiled reference parser call by call, and against golden vectors where the reference is absent.  CPU only."""
import ctypes
import hashlib
from pathlib import Path

import numpy as np
import pytest

import ffmpeg_ht_b200 as fb
import j2kgen
from helpers import ROOT

REF = ROOT / "oracle" / "_ref" / "libj2kref.so"
GOLD = ROOT / "tests" / "golden" / "splitter.npz"


SHIM = ROOT / "oracle" / "_ref" / "libj2kb200shim.so"


def ref_lib(path=REF, prefix="ref"):
    """The harness of oracle/ref_shim.c around an AVCodecParser: the reference's (libj2kref.so, ref_*) or the
    product's AVCodecParser shim (libj2kb200shim.so, b200shim_*)."""
    L = ctypes.CDLL(str(path))
    new, parse, free = (getattr(L, f"{prefix}_parser_{n}") for n in ("new", "parse", "free"))
    new.restype = ctypes.c_void_p
    new.argtypes = [ctypes.c_int]
    parse.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
                      ctypes.POINTER(ctypes.c_int), ctypes.c_void_p, ctypes.c_int64]
    free.argtypes = [ctypes.c_void_p]
    L.ref_parser_new, L.ref_parser_parse, L.ref_parser_free = new, parse, free
    return L


def streams(seed):
    """A few small codestreams / JP2 files of different layouts (tiles => several SOT, SOP/EPH, 9/7, JP2 boxes)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(5):
        w, h = int(rng.integers(17, 90)), int(rng.integers(9, 70))
        nc = int(rng.choice([1, 3]))
        planes = j2kgen.synth_image(w, h, nc, 8, seed=seed * 10 + i)
        kw = dict(levels=int(rng.integers(0, 4)), cblk=(32, 32))
        if rng.random() < 0.5:
            kw["tile"] = (int(rng.integers(16, 64)), int(rng.integers(16, 64)))
        if rng.random() < 0.3:
            kw.update(reversible=False, base_step=1.0)
        if rng.random() < 0.3:
            kw.update(sop=True, eph=True)
        cs = j2kgen.encode(planes, 8, j2kgen.EncParams(**kw))
        if rng.random() < 0.4:
            cs = j2kgen.wrap_jp2(cs, w, h, nc, 8, colour_space=16 if nc == 3 else 17)
        out.append(cs)
    return out


def chunkings(total, rng):
    """Piece sizes covering `total` bytes: single bytes around, small and large random pieces."""
    mode = int(rng.integers(0, 4))
    sizes, left = [], total
    while left > 0:
        n = {0: 1, 1: int(rng.integers(1, 8)), 2: int(rng.integers(1, 700)), 3: int(rng.integers(1, 9000))}[mode]
        n = min(n, left)
        sizes.append(n)
        left -= n
    return sizes


def drive(parse, blob, sizes):
    """av_parser_parse2's loop over the pieces; parse(bytes) -> (ret, frame or None).  Returns the call trace."""
    trace, pos = [], 0
    for n in sizes:
        piece = blob[pos:pos + n]
        pos += n
        off = 0
        guard = 0
        while off < len(piece):
            r, f = parse(piece[off:])
            trace.append((r, None if f is None else (len(f), hashlib.sha1(f).hexdigest())))
            off += max(r, 0)
            guard += 1
            if (r <= 0 and f is None) or guard > 10000:
                break
    r, f = parse(b"")
    trace.append((r, None if f is None else (len(f), hashlib.sha1(f).hexdigest())))
    return trace


def ours_parse_fn():
    sp = fb.FrameSplitter()
    return sp, sp.parse


def ref_parse_fn(L):
    h = L.ref_parser_new(0)
    copy = ctypes.create_string_buffer(1 << 22)

    def parse(data):
        # the reference reads up to 64 bytes of padding past the piece (libavcodec pads its packets)
        buf = ctypes.create_string_buffer(bytes(data) + b"\0" * 64, len(data) + 64)
        off, n = ctypes.c_int64(0), ctypes.c_int(0)
        r = L.ref_parser_parse(h, buf, len(data), ctypes.byref(off), ctypes.byref(n), copy, len(copy))
        return r, (copy.raw[:n.value] if n.value > 0 else None)
    return h, parse


CASES = [(s, c) for s in range(6) for c in range(4)]


def make_case(seed, c):
    rng = np.random.default_rng(1000 * seed + c)
    ss = streams(seed)
    if c == 3:                      # damage: random bytes, a truncated stream, garbage between frames
        ss = [bytearray(x) for x in ss]
        for x in ss[:2]:
            for _ in range(6):
                x[int(rng.integers(0, len(x)))] = int(rng.integers(0, 256))
        ss[2] = ss[2][:len(ss[2]) // 2]
        ss.insert(1, bytes(rng.integers(0, 256, 37, dtype=np.uint8)))
        ss = [bytes(x) for x in ss]
    blob = b"".join(ss)
    return blob, chunkings(len(blob), rng), ss


@pytest.mark.skipif(not REF.exists(), reason="reference not built here (oracle/_ref)")
@pytest.mark.parametrize("seed,c", CASES)
def test_splitter_matches_reference_call_by_call(seed, c):
    L = ref_lib()
    blob, sizes, _ = make_case(seed, c)
    sp, ours = ours_parse_fn()
    h, ref = ref_parse_fn(L)
    try:
        assert drive(ours, blob, sizes) == drive(ref, blob, sizes)
    finally:
        L.ref_parser_free(h)
        sp.close()


@pytest.mark.skipif(not (REF.exists() and SHIM.exists()), reason="reference / shim not built here (oracle/_ref)")
@pytest.mark.parametrize("seed,c", [(0, 1), (1, 3), (2, 2), (3, 0)])
def test_avcodecparser_shim_matches_reference(seed, c):
    """ffcodec/jpeg2000_parser_b200.c compiled against the reference's headers, driven through the same
    harness as ff_jpeg2000_parser."""
    L, S = ref_lib(), ref_lib(SHIM, "b200shim")
    blob, sizes, _ = make_case(seed, c)
    h, ref = ref_parse_fn(L)
    g, shim = ref_parse_fn(S)
    try:
        assert drive(shim, blob, sizes) == drive(ref, blob, sizes)
    finally:
        L.ref_parser_free(h)
        S.ref_parser_free(g)


def test_splitter_recovers_the_frames():
    for seed in range(3):
        ss = streams(seed)
        blob = b"".join(ss)
        for step in (1, 5, 4096, len(blob)):
            sp = fb.FrameSplitter()
            got = []
            for i in range(0, len(blob), step):
                got += sp.feed(blob[i:i + step])
            got += sp.feed(b"", flush=True)
            assert got == ss, (seed, step)


def test_complete_frames_flag_passes_packets_through():
    sp = fb.FrameSplitter(complete_frames=True)
    for cs in streams(1):
        r, f = sp.parse(cs)
        assert r == len(cs) and f == cs


def test_find_frame_end_alone():
    cs = streams(2)[0]
    L = fb.lib()
    h = ctypes.c_void_p()
    assert L.b200j2k_splitter_create(ctypes.byref(h), 0) == 0
    two = cs + cs
    a = np.frombuffer(two, np.uint8)
    raw = cs[:2] == b"\xff\x4f"
    r = L.b200j2k_splitter_find_frame_end(h, a.ctypes.data, len(two))
    # a raw codestream ends behind its EOC; a JP2 file only when the next signature box shows up
    assert r == (len(cs) if raw else len(cs)), r
    assert L.b200j2k_splitter_find_frame_end(h, a.ctypes.data, 0) == 0
    L.b200j2k_splitter_destroy(h)


def test_golden_traces():
    """Reference-generated call traces (tests/golden/make_splitter_golden.py), for boxes without the reference."""
    g = np.load(GOLD, allow_pickle=False)
    for k in range(int(g["ncases"])):
        blob = g[f"blob{k}"].tobytes()
        sizes = [int(x) for x in g[f"sizes{k}"]]
        want = [(int(r), None if n < 0 else (int(n), h)) for r, n, h in
                zip(g[f"ret{k}"], g[f"len{k}"], [str(x) for x in g[f"sha{k}"]])]
        sp, ours = ours_parse_fn()
        assert drive(ours, blob, sizes) == want, k
        sp.close()
