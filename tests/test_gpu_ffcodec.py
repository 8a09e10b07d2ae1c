"""The drop-in boundary itself, on the GPU box: the product's FFCodec shim
(ffmpeg-ht_b200/ffcodec/jpeg2000_b200.c), compiled against the reference's own headers and driven by the
same harness that drives the reference's ff_jpeg2000_decoder (oracle/ref_shim.c), must give the same
AVFrame bytes, dimensions and pixel format as the reference decoder."""
import ctypes

import numpy as np
import pytest

import j2kgen
from helpers import ORACLE_DIR, have_ref, ref_decode

SHIM = ORACLE_DIR / "_ref" / "libj2kb200shim.so"
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (have_ref() and SHIM.exists()), reason="oracle/_ref not built (needs /root/reference)")]


def shim_decode(cs: bytes, bitexact=0):
    lib = ctypes.CDLL(str(SHIM))
    lib.b200shim_decode.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
                                    ctypes.c_void_p, ctypes.c_int64]
    info = (ctypes.c_int64 * 8)()
    out = np.zeros(1 << 26, np.uint8)
    ret = lib.b200shim_decode(cs, len(cs), bitexact, info, out.ctypes.data, out.size)
    return ret, list(info), out[: max(info[3], 0)].copy()


@pytest.mark.parametrize("kw", [dict(nc=3, bd=12, levels=4, tile=(128, 96)),
                                dict(nc=1, bd=16, levels=5),
                                dict(nc=3, bd=10, levels=5, reversible=False, base_step=2.0),
                                dict(nc=3, bd=8, levels=3, cblk=(32, 32), npasses=3, prog=2, precincts=[(6, 6), (7, 7)]),
                                dict(nc=4, bd=12, levels=2, cblk=(16, 16))])
def test_ffcodec_shim_equals_reference(kw):
    kw = dict(kw)
    nc, bd = kw.pop("nc"), kw.pop("bd")
    planes = j2kgen.synth_image(300, 200, nc, bd, seed=9)
    p = j2kgen.EncParams(**kw)
    cs = j2kgen.encode(planes, bd, p)
    for bitexact in ((0,) if p.reversible else (0, 1)):
        for lowres in (0, 1, 2):           # AVCodecContext.lowres, passed in bits 8-11 of the harness flags
            flags = bitexact | (lowres << 8)
            r1, i1, o1, _ = ref_decode(cs, flags)
            r2, i2, o2 = shim_decode(cs, flags)
            assert r1 == 0 and r2 == 0
            assert i1[:3] == i2[:3]            # width, height, AVPixelFormat
            assert i1[4] == i2[4]              # bytes consumed (decode() return value)
            assert np.array_equal(o1, o2)


def test_ffcodec_shim_part1_streams():
    """Part-1 (MQ) streams from OpenJPEG through both FFCodecs: same frame bytes, format, bytes consumed; lowres too."""
    from test_oracle_vs_ref import PART1, part1_stream
    for t in (PART1[0], PART1[1], PART1[3], PART1[7], PART1[9], PART1[10]):
        for jp2 in (False, True):
            cs = part1_stream(t, jp2)
            for bitexact in ((0,) if not t["kw"].get("irreversible") else (0, 1)):
                for lowres in (0, 1):
                    flags = bitexact | (lowres << 8)
                    r1, i1, o1, _ = ref_decode(cs, flags)
                    r2, i2, o2 = shim_decode(cs, flags)
                    assert r1 == r2, (t["mode"], jp2, flags, r1, r2)
                    if r1 == 0:
                        assert i1[:3] == i2[:3] and i1[4] == i2[4]
                        assert np.array_equal(o1, o2)


def test_ffcodec_shim_jp2_and_subsampled():
    """JP2 / JPH files and chroma-subsampled streams through the FFCodec shim: same AVPixelFormat (planar YUV
    included), same AVFrame bytes as the reference decoder."""
    from test_oracle_vs_ref import SUBSAMPLED, subsampled_stream
    cases = []
    for name, subs, bd in SUBSAMPLED[:6]:
        cs, planes, p = subsampled_stream(subs, bd, 131, 77, dict(levels=3, tile=(64, 48)))
        cases += [cs, j2kgen.wrap_jp2(cs, 131, 77, len(subs), bd, colour_space=18, cdef=[(0, 0, 1), (1, 0, 3), (2, 0, 2)])]
    for nc, bd in [(3, 8), (3, 12), (1, 16), (4, 8)]:
        planes = j2kgen.synth_image(160, 90, nc, bd, seed=8)
        cs = j2kgen.encode(planes, bd, j2kgen.EncParams(levels=3, mct=nc >= 3))
        cases += [j2kgen.wrap_jp2(cs, 160, 90, nc, bd, colour_space=csp) for csp in (16, 17, 18)]
    from test_oracle_vs_ref import PAL8_CASES, dcinema_cases, pal8_file
    cases += [pal8_file(bd, csp, pal)[0] for bd, csp, pal in PAL8_CASES]       # AV_PIX_FMT_PAL8: data[1] = palette
    cases += [cs for name, cs in dcinema_cases() if "x8" not in name and "8-bit" not in name]     # XYZ12 and friends
    n = 0
    for data in cases:
        r1, i1, o1, _ = ref_decode(data)
        r2, i2, o2 = shim_decode(data)
        if r1 < 0:
            assert r2 == r1
            continue
        assert r2 == 0 and i1[:3] == i2[:3] and i1[4] == i2[4]
        assert np.array_equal(o1, o2)
        n += 1
    assert n >= 20


def test_ffcodec_shim_sample_aspect_ratio():
    """AVCodecContext.sample_aspect_ratio from the 'res ' box of a JP2 file, like the reference decoder sets it."""
    from test_oracle_vs_ref import RES_BOXES, res_box_file
    for res in RES_BOXES:
        data = res_box_file(res)
        r1, i1, o1, _ = ref_decode(data)
        r2, i2, o2 = shim_decode(data)
        assert r1 == 0 and r2 == 0 and np.array_equal(o1, o2)
        assert (i1[5], i1[6]) == (i2[5], i2[6]), res


def test_ffcodec_shim_frame_pool_on_and_off(monkeypatch):
    """The shim's own pool of page-locked frames (private option pinned_frames, default on while the caller uses the
    default get_buffer2) and the ff_thread_get_buffer() path give the same AVFrame bytes as the reference, for packed,
    planar and palettised formats; a change of picture size re-creates the pool."""
    from test_oracle_vs_ref import PAL8_CASES, SUBSAMPLED, pal8_file, subsampled_stream
    cases = [pal8_file(*PAL8_CASES[0])[0], subsampled_stream(SUBSAMPLED[0][1], SUBSAMPLED[0][2], 131, 77, dict(levels=3))[0]]
    for w, h, nc, bd in ((160, 90, 3, 12), (333, 111, 3, 12), (64, 64, 1, 16), (160, 90, 4, 8)):
        planes = j2kgen.synth_image(w, h, nc, bd, seed=8)
        cases.append(j2kgen.encode(planes, bd, j2kgen.EncParams(levels=3, mct=nc >= 3)))
    for pinned in ("1", "0"):
        monkeypatch.setenv("REF_OPT_pinned_frames", pinned)
        for data in cases:
            r1, i1, o1, _ = ref_decode(data)
            r2, i2, o2 = shim_decode(data)
            assert r1 == 0 and r2 == 0 and i1[:3] == i2[:3] and np.array_equal(o1, o2), pinned


def test_ffcodec_shim_error_codes():
    planes = j2kgen.synth_image(64, 64, 1, 8, seed=2)
    cs = j2kgen.encode(planes, 8, j2kgen.EncParams(levels=2))
    r1, *_ = ref_decode(cs[:40])
    r2, *_ = shim_decode(cs[:40])
    assert r1 < 0 and r2 < 0


def shim_decode_cuda(streams, mode, sw_name, w, h):
    lib = ctypes.CDLL(str(SHIM))
    n = len(streams)
    bufs = [np.frombuffer(s, np.uint8) for s in streams]
    ptrs = (ctypes.c_void_p * n)(*[b.ctypes.data for b in bufs])
    lens = (ctypes.c_int64 * n)(*[b.size for b in bufs])
    info = (ctypes.c_int64 * (8 * n))()
    out = np.zeros(1 << 26, np.uint8)
    lib.b200shim_decode_cuda.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p,
                                         ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64), ctypes.c_void_p,
                                         ctypes.c_int64]
    ret = lib.b200shim_decode_cuda(ptrs, lens, n, mode, sw_name.encode(), w, h, info, out.ctypes.data, out.size)
    frames, pos = [], 0
    for i in range(n):
        fi = list(info[8 * i: 8 * i + 8])
        frames.append((fi, out[pos: pos + max(fi[3], 0)].copy()))
        pos += max(fi[3], 0)
    return ret, frames


@pytest.mark.parametrize("kw", [dict(nc=3, bd=12, levels=4), dict(nc=1, bd=16, levels=3, tile=(128, 128)),
                                dict(nc=3, bd=10, levels=3, reversible=False, base_step=2.0),
                                dict(nc=3, bd=8, levels=2, sub=[(1, 1), (2, 2), (2, 2)])])
def test_ffcodec_shim_cuda_frames(kw):
    """SURVEY 8(f) N2 through the reference's own hardware-frame API: an application offers a CUDA frames context
    (mode 0) or only a CUDA device (mode 1); the shim proposes AV_PIX_FMT_CUDA through get_format, takes the picture
    from av_hwframe_get_buffer() and leaves it in device memory (AVFrame.data[] are device pointers, sw_pix_fmt is the
    reference's format).  Downloaded, every frame of a three-frame sequence equals the reference decoder's AVFrame;
    when get_format declines the hardware format (mode 2) the ordinary software frames come out.  libavutil's
    generic hwcontext.c is the reference's own; the CUDA type behind it is the stand-in of oracle/hwcuda_standin.c."""
    kw = dict(kw)
    nc, bd, sub = kw.pop("nc"), kw.pop("bd"), kw.pop("sub", None)
    w, h = 272, 176
    streams = []
    for seed in (1, 2, 3):
        full = j2kgen.synth_image(w, h, nc, bd, seed=seed)
        if sub:
            planes = [np.ascontiguousarray(a[::sy, ::sx]) for a, (sx, sy) in zip(full, sub)]
            streams.append(j2kgen.encode(planes, bd, j2kgen.EncParams(mct=False, subsampling=sub, **kw)))
        else:
            streams.append(j2kgen.encode(full, bd, j2kgen.EncParams(**kw)))
    want = [ref_decode(cs, 0) for cs in streams]
    assert all(r[0] == 0 for r in want)
    sw_name = want[0][3]
    AV_PIX_FMT_CUDA = None
    for mode in (0, 1, 2):
        ret, frames = shim_decode_cuda(streams, mode, sw_name, w, h)
        assert ret == 0, (mode, ret)
        for (fi, got), (r, i1, o1, name) in zip(frames, want):
            assert fi[0] == i1[0] and fi[1] == i1[1]
            assert np.array_equal(got, o1), mode
            if mode < 2:
                assert fi[4] == 1 and fi[5] == i1[2]        # device memory; sw_pix_fmt = the reference's pix_fmt
                AV_PIX_FMT_CUDA = AV_PIX_FMT_CUDA or fi[2]
                assert fi[2] == AV_PIX_FMT_CUDA and fi[2] != i1[2]
            else:
                assert fi[4] == 0 and fi[2] == i1[2]


def decode_seq(lib, prefix, streams, flags=0):
    n = len(streams)
    bufs = [np.frombuffer(s, np.uint8) for s in streams]
    ptrs = (ctypes.c_void_p * n)(*[b.ctypes.data for b in bufs])
    lens = (ctypes.c_int64 * n)(*[b.size for b in bufs])
    info = (ctypes.c_int64 * (8 * n))()
    nout = ctypes.c_int(0)
    out = np.zeros(1 << 26, np.uint8)
    fn = getattr(lib, prefix + "decode_seq")
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
                   ctypes.POINTER(ctypes.c_int), ctypes.c_void_p, ctypes.c_int64]
    ret = fn(ptrs, lens, n, flags, info, ctypes.byref(nout), out.ctypes.data, out.size)
    frames, pos = [], 0
    for i in range(nout.value):
        fi = list(info[8 * i: 8 * i + 8])
        frames.append((fi[:3], out[pos: pos + fi[3]].copy()))
        pos += fi[3]
    return ret, frames


@pytest.mark.parametrize("depth", [1, 2, 4])
def test_ffcodec_shim_pictures_in_flight(depth, monkeypatch):
    """Private option async_depth: up to that many pictures in flight inside one decoder context, the frame of packet
    N comes out while packet N + depth - 1 goes in, the rest when the stream is drained (AV_CODEC_CAP_DELAY).  Seven
    frames of changing content AND size through one context: every frame, in order, equals the reference decoder's."""
    from helpers import ref
    streams = []
    for i, (w, h) in enumerate([(300, 200)] * 3 + [(208, 144)] * 2 + [(300, 200)] * 2):
        planes = j2kgen.synth_image(w, h, 3, 12, seed=40 + i)
        streams.append(j2kgen.encode(planes, 12, j2kgen.EncParams(levels=3, tile=(128, 128) if i % 2 else (0, 0))))
    r1, want = decode_seq(ref(), "ref_", streams)
    assert r1 == 0 and len(want) == len(streams)
    monkeypatch.setenv("REF_OPT_async_depth", str(depth))
    r2, got = decode_seq(ctypes.CDLL(str(SHIM)), "b200shim_", streams)
    assert r2 == 0 and len(got) == len(streams)
    for (i1, o1), (i2, o2) in zip(want, got):
        assert i1 == i2 and np.array_equal(o1, o2)
