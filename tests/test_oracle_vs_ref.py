"""Pins the oracle (oracle/*.c restatement) against the UNMODIFIED reference compiled by oracle/Makefile
(oracle/_ref/libj2kref.so).  Runs wherever that library exists (this container; the GPU box gets the
prebuilt .so with the snapshot); elsewhere the golden fixtures in tests/golden do the same job
(tests/test_golden.py).  CPU only."""
import numpy as np
import pytest

import j2kgen
from helpers import have_ref, ht_block, oracle, oracle_decode, oracle_pix_name, ref, ref_decode, sm_to_int, ORACLE_PIX

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libj2kref.so not built (needs /root/reference)")


def test_ht_block_random_bytes_bit_exact():
    """Damaged / random segments: same return code and same coefficients (zeros on error, quirk Q1)."""
    rng = np.random.default_rng(1234)
    decoded = 0
    for _ in range(1500):
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        Lcup = int(rng.integers(2, 300))
        npasses = int(rng.choice([0, 1, 1, 2, 3, 3, 4, 5, 6, 7]))
        Lref = int(rng.integers(0, 60)) if npasses % 3 != 1 else 0
        data = bytearray(rng.integers(0, 256, Lcup + Lref, dtype=np.uint8).tobytes())
        if rng.random() < 0.9:   # mostly valid suffix lengths, otherwise nearly everything stops at Scup
            scup = int(rng.integers(2, min(Lcup, 200) + 1))
            data[Lcup - 1] = scup >> 4
            data[Lcup - 2] = (data[Lcup - 2] & 0xF0) | (scup & 15)
        M_b = int(rng.integers(2, 28))
        zbp = int(rng.integers(0, M_b))
        vsc = int(rng.integers(0, 2))
        ra, a = ht_block(ref().ref_ht_block, bytes(data), npasses, Lcup, Lref, zbp, w, h, M_b, 0, vsc)
        rb, b = ht_block(oracle().oracle_ht_block, bytes(data), npasses, Lcup, Lref, zbp, w, h, M_b, 0, vsc)
        assert ra == rb, (w, h, npasses, Lcup, Lref, zbp, M_b)
        assert np.array_equal(a, b), (w, h, npasses, Lcup, Lref, zbp, M_b)
        decoded += ra == 1
    assert decoded > 100


@pytest.mark.parametrize("npasses", [1, 2, 3])
def test_ht_block_valid_segments(npasses):
    """Dense valid code-blocks from the test encoder, all sizes incl. odd ones, with ROI shift."""
    rng = np.random.default_rng(99 + npasses)
    n = 0
    for _ in range(250):
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        bits = int(rng.integers(1, 14))
        coef = (rng.standard_normal((h, w)) * rng.uniform(0.3, 2 ** bits)).astype(np.int32)
        if rng.random() < 0.3:
            coef[rng.random((h, w)) < 0.8] = 0
        refine = npasses > 1
        if not np.any(np.abs(coef) >> refine):
            continue
        M_b = bits + 6
        zbp = M_b - 1 - refine
        vsc = int(rng.integers(0, 2))
        roi = int(rng.choice([0, 0, 3]))
        data, Lcup, Lref = j2kgen.encode_block(coef, npasses, bool(vsc), zbp + 1)
        ra, a = ht_block(ref().ref_ht_block, data, npasses, Lcup, Lref, zbp, w, h, M_b, roi, vsc)
        rb, b = ht_block(oracle().oracle_ht_block, data, npasses, Lcup, Lref, zbp, w, h, M_b, roi, vsc)
        assert ra == 1 and rb == 1
        assert np.array_equal(a, b)
        if npasses == 1 and roi == 0:
            assert np.array_equal(sm_to_int(a, M_b), coef)
        n += 1
    assert n > 150


@pytest.mark.parametrize("dwt_type", [0, 1, 2])
def test_idwt_bit_exact(dwt_type):
    """W1-W4 on random borders (odd/even origins, lengths down to 1, 0..6 levels), compared as raw bits."""
    rng = np.random.default_rng(7 + dwt_type)
    for _ in range(200):
        x0, y0 = int(rng.integers(0, 40)), int(rng.integers(0, 40))
        w, h = int(rng.integers(1, 48)), int(rng.integers(1, 48))
        levels = int(rng.integers(0, 7))
        border = np.array([x0, x0 + w, y0, y0 + h], np.int32)
        if dwt_type == 0:
            a = (rng.standard_normal(w * h) * 1000).astype(np.float32)
        else:
            a = rng.integers(-(1 << 20), 1 << 20, w * h).astype(np.int32)
        b = a.copy()
        assert ref().ref_idwt(dwt_type, border.ctypes.data, levels, a.ctypes.data) == 0
        assert oracle().oracle_idwt(dwt_type, border.ctypes.data, levels, b.ctypes.data) == 0
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (x0, y0, w, h, levels)


@pytest.mark.parametrize("dwt_type", [0, 1, 2])
def test_mct_bit_exact(dwt_type):
    rng = np.random.default_rng(17 + dwt_type)
    n = 4099
    if dwt_type == 0:
        p = [(rng.standard_normal(n) * 3000).astype(np.float32) for _ in range(3)]
    else:
        p = [rng.integers(-(1 << 30), 1 << 30, n).astype(np.int32) for _ in range(3)]
    q = [x.copy() for x in p]
    ref().ref_mct(dwt_type, *[x.ctypes.data for x in p], n)
    oracle().oracle_mct(dwt_type, *[x.ctypes.data for x in q], n)
    for a, b in zip(p, q):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


STREAMS = [
    dict(w=64, h=64, nc=1, bd=8, levels=0),
    dict(w=256, h=256, nc=1, bd=8, levels=3, reversible=False, mct=False, base_step=1.0),
    dict(w=256, h=200, nc=3, bd=10, levels=5, reversible=False, base_step=2.0),
    dict(w=250, h=131, nc=3, bd=12, levels=5, tile=(128, 64)),
    dict(w=250, h=131, nc=3, bd=12, levels=5, offset=(3, 5)),
    dict(w=250, h=131, nc=3, bd=12, levels=4, offset=(13, 7), tile=(100, 50)),
    dict(w=250, h=131, nc=3, bd=10, levels=4, offset=(13, 7), tile=(100, 50), reversible=False, base_step=0.5),
    dict(w=256, h=256, nc=3, bd=12, levels=3, cblk=(32, 32), precincts=[(7, 7), (8, 8)], prog=2),
    dict(w=256, h=256, nc=3, bd=12, levels=3, cblk=(32, 32), precincts=[(6, 6), (7, 7)], prog=3),
    dict(w=256, h=256, nc=3, bd=12, levels=3, cblk=(32, 32), precincts=[(6, 6), (7, 7)], prog=4),
    dict(w=256, h=256, nc=3, bd=12, levels=3, cblk=(32, 32), precincts=[(6, 6), (7, 7)], prog=1, sop=True, eph=True),
    dict(w=256, h=256, nc=3, bd=12, levels=3, cblk=(32, 32), npasses=3),
    dict(w=255, h=253, nc=3, bd=12, levels=3, cblk=(32, 32), npasses=3, vsc=True),
    dict(w=255, h=253, nc=1, bd=16, levels=5, npasses=2),
    dict(w=300, h=200, nc=1, bd=16, levels=5, cblk=(128, 32)),
    dict(w=300, h=200, nc=1, bd=8, levels=5, cblk=(4, 4)),
    dict(w=300, h=200, nc=2, bd=8, levels=5, cblk=(1024, 4), mct=False),
    dict(w=300, h=200, nc=4, bd=12, levels=2, cblk=(16, 16)),
    dict(w=1, h=1, nc=1, bd=8, levels=0),
    dict(w=3, h=2, nc=1, bd=8, levels=2),
    dict(w=17, h=1, nc=3, bd=9, levels=5),
    dict(w=1, h=33, nc=1, bd=12, levels=3),
]


@pytest.mark.parametrize("cfg", STREAMS, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items())[:60])
def test_frame_decode_bit_exact(cfg):
    """Whole codestreams: oracle pixels == reference pixels, float 9/7 path included (compared as bytes),
    and the integer 9/7 path the reference takes under AV_CODEC_FLAG_BITEXACT."""
    cfg = dict(cfg)
    w, h, nc, bd = cfg.pop("w"), cfg.pop("h"), cfg.pop("nc"), cfg.pop("bd")
    planes = j2kgen.synth_image(w, h, nc, bd, seed=11)
    p = j2kgen.EncParams(**cfg)
    cs = j2kgen.encode(planes, bd, p)
    for bitexact in ((0,) if p.reversible else (0, 1)):
        r1, info1, out1, name = ref_decode(cs, bitexact)
        r2, info2, out2 = oracle_decode(cs, bitexact)
        assert r1 == 0 and r2 == 0
        assert info1[:2] == info2[:2]
        assert ORACLE_PIX[info2[2]] == name
        assert np.array_equal(out1, out2)
    if p.reversible and p.npasses == 1:
        bps = 2 if bd > 8 else 1
        img = out2.view(np.uint16 if bps == 2 else np.uint8).reshape(h, w, nc)
        shift = (16 - bd) if (bps == 2 and nc != 2) else 0
        assert np.array_equal(img >> shift, np.stack(planes, -1))


@pytest.mark.parametrize("cfg", [STREAMS[2], STREAMS[3], STREAMS[6], STREAMS[12], STREAMS[14], STREAMS[16]],
                         ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items())[:60])
def test_lowres_decode_bit_exact(cfg):
    """AVCodecContext.lowres = 1 .. levels + 1 (the last one is the reference's "reduction_factor too large"
    EINVAL; streams with an image offset are AVERROR_PATCHWELCOME): oracle == reference, dimensions,
    pixels and error codes."""
    cfg = dict(cfg)
    w, h, nc, bd = cfg.pop("w"), cfg.pop("h"), cfg.pop("nc"), cfg.pop("bd")
    planes = j2kgen.synth_image(w, h, nc, bd, seed=11)
    p = j2kgen.EncParams(**cfg)
    cs = j2kgen.encode(planes, bd, p)
    for lowres in range(1, p.levels + 2):
        for bitexact in ((0,) if p.reversible else (0, 1)):
            flags = bitexact | (lowres << 8)
            r1, info1, out1, name = ref_decode(cs, flags)
            r2, info2, out2 = oracle_decode(cs, flags)
            assert r1 == r2, (lowres, r1, r2)
            if r1 == 0:
                assert info1[:2] == info2[:2] and ORACLE_PIX[info2[2]] == name
                assert np.array_equal(out1, out2), lowres


# two quality layers: HT sets split over the layers, blocks first included in the second layer, triples of empty
# placeholder passes in front of the HT set in the same or in an earlier packet (jpeg2000dec.c:1256-1362)
LAYERED = [
    dict(w=128, h=96, nc=3, bd=12, levels=3, cblk=(32, 32)),
    dict(w=128, h=96, nc=3, bd=12, levels=3, cblk=(32, 32), npasses=3),
    dict(w=128, h=96, nc=1, bd=8, levels=2, cblk=(16, 16), npasses=2, vsc=True),
    dict(w=200, h=120, nc=3, bd=10, levels=4, npasses=3, prog=2, precincts=[(6, 6), (7, 7)], cblk=(32, 32)),
    dict(w=200, h=120, nc=3, bd=10, levels=4, npasses=3, prog=3, precincts=[(6, 6), (7, 7)], cblk=(32, 32)),
    dict(w=200, h=120, nc=3, bd=10, levels=4, npasses=3, prog=4, precincts=[(6, 6), (7, 7)], cblk=(32, 32), sop=True, eph=True),
    dict(w=200, h=120, nc=3, bd=10, levels=4, npasses=1, prog=1, tile=(128, 64)),
    dict(w=160, h=100, nc=3, bd=10, levels=3, reversible=False, base_step=1.0, npasses=3),
    dict(w=300, h=200, nc=1, bd=16, levels=5, cblk=(128, 32), npasses=3),
]


@pytest.mark.parametrize("cfg", LAYERED, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items())[:60])
def test_two_layer_streams_bit_exact(cfg):
    """The same picture coded in one and in two quality layers: the reference decodes both to identical
    pixels (this validates the generator), and the oracle equals the reference on the two-layer stream."""
    cfg = dict(cfg)
    w, h, nc, bd = cfg.pop("w"), cfg.pop("h"), cfg.pop("nc"), cfg.pop("bd")
    planes = j2kgen.synth_image(w, h, nc, bd, seed=4)
    one = j2kgen.encode(planes, bd, j2kgen.EncParams(**cfg))
    two = j2kgen.encode(planes, bd, j2kgen.EncParams(layers=2, **cfg))
    assert len(two) != len(one)
    for bitexact in ((0,) if cfg.get("reversible", True) else (0, 1)):
        r0, _, out0, _ = ref_decode(one, bitexact)
        r1, info1, out1, name = ref_decode(two, bitexact)
        r2, info2, out2 = oracle_decode(two, bitexact)
        assert r0 == 0 and r1 == 0 and r2 == 0
        assert np.array_equal(out0, out1)
        assert info1[:2] == info2[:2] and oracle_pix_name() == name and np.array_equal(out1, out2)


# progression order changes (POC marker in the main header, get_poc jpeg2000dec.c:760-818): every entry is
# (RSpoc, CSpoc, LYEpoc, REpoc, CEpoc, Ppoc); entries overlap, leave CEpoc 0 (= all) and REpoc 33, and span two segments
POC_STREAMS = [
    dict(cfg=dict(levels=3), poc=[(2, 0, 1, 4, 3, 0), (0, 0, 1, 2, 3, 1)]),
    dict(cfg=dict(levels=3, layers=2, npasses=3), poc=[(0, 0, 1, 4, 3, 1), (0, 0, 2, 4, 3, 4)]),
    dict(cfg=dict(levels=4, precincts=[(6, 6), (7, 7)], cblk=(32, 32)),
         poc=[(0, 0, 1, 3, 3, 2), (0, 0, 1, 5, 1, 4), (3, 1, 1, 5, 3, 3)]),
    dict(cfg=dict(levels=4, precincts=[(6, 6), (7, 7)], cblk=(32, 32), layers=2, npasses=3),
         poc=[(0, 0, 1, 5, 3, 4), (0, 0, 2, 2, 3, 2), (2, 0, 2, 33, 0, 3)], split=1),
    dict(cfg=dict(levels=3, tile=(128, 64), sop=True, eph=True), poc=[(1, 0, 1, 4, 3, 1), (0, 0, 1, 1, 3, 0)]),
    dict(cfg=dict(levels=2, reversible=False, base_step=1.0), poc=[(0, 1, 1, 3, 3, 1), (0, 0, 1, 3, 1, 0)]),
]


def poc_stream(t, w=200, h=120, nc=3, bd=10, seed=4):
    """-> (the stream with the POC marker, the same picture in the COD progression)"""
    planes = j2kgen.synth_image(w, h, nc, bd, seed=seed)
    base = j2kgen.encode(planes, bd, j2kgen.EncParams(**t["cfg"]))
    cs = j2kgen.encode(planes, bd, j2kgen.EncParams(poc=t["poc"], poc_split=t.get("split", 0), **t["cfg"]))
    return cs, base


@pytest.mark.parametrize("t", POC_STREAMS, ids=lambda t: "-".join("".join(map(str, e)) for e in t["poc"]))
def test_poc_streams_bit_exact(t):
    """The reference decodes the re-sequenced stream to the pixels of the plain one (this validates the
    generator's packet order), and the oracle follows the same entries."""
    cs, base = poc_stream(t)
    assert len(cs) > len(base) and cs[cs.index(b"\xff\x93"):] != base[base.index(b"\xff\x93"):]     # re-sequenced
    for bitexact in ((0,) if t["cfg"].get("reversible", True) else (0, 1)):
        r0, _, out0, _ = ref_decode(base, bitexact)
        r1, info1, out1, name = ref_decode(cs, bitexact)
        r2, info2, out2 = oracle_decode(cs, bitexact)
        assert r0 == 0 and r1 == 0 and r2 == 0
        assert np.array_equal(out0, out1)
        assert info1[:2] == info2[:2] and oracle_pix_name() == name and np.array_equal(out1, out2)


def poc_marker_cases():
    """Bad entries (RSpoc >= REpoc, REpoc > 33, CSpoc >= CEpoc, LYEpoc == 0), a segment length that is not
    2 + 7n, and more than 32 entries over several segments."""
    planes = j2kgen.synth_image(64, 48, 3, 8, seed=2)
    cs = j2kgen.encode(planes, 8, j2kgen.EncParams(levels=2, poc=[(0, 0, 1, 3, 3, 1)]))
    at = cs.index(b"\xff\x5f")
    assert cs[at + 2:at + 4] == b"\x00\x09"
    ent = at + 4
    cases = []
    for off, val in ((0, 3), (4, 34), (1, 3), (2, 0), (5, 1), (5, 4), (6, 7), (4, 1), (0, 2)):
        bad = bytearray(cs)
        bad[ent + off] = val
        if off == 2:
            bad[ent + 3] = 0
        cases.append(bytes(bad))
    cases.append(cs[:at + 3] + b"\x0a" + cs[at + 4:at + 11] + b"\x00" + cs[at + 11:])        # one stray byte
    seg = cs[at:at + 11]
    cases.append(cs[:at] + seg * 33 + cs[at + 11:])                                          # 33 entries
    cases.append(cs[:at] + seg * 32 + cs[at + 11:])                                          # exactly 32
    return cases


def test_poc_marker_validation_matches_reference():
    """get_poc's checks (jpeg2000dec.c:760-818): same verdict as the reference, same pixels where it decodes."""
    cases = poc_marker_cases()
    assert len(cases) == 12
    for bad in cases:
        r1, _, out1, _ = ref_decode(bad, 0)
        r2, _, out2 = oracle_decode(bad, 0)
        assert (r1 == 0) == (r2 == 0), (r1, r2)
        if r1 == 0:
            assert np.array_equal(out1, out2)
    assert sum(ref_decode(bad, 0)[0] == 0 for bad in cases) >= 3


# packed packet headers: PPT (tile-part header) and PPM (main header); chunk = payload bytes per marker segment
PACKED = [
    dict(cfg=dict(levels=3, tile=(128, 64), sop=True, eph=True), mode=1, chunk=0),
    dict(cfg=dict(levels=3, tile=(128, 64), sop=True, eph=True), mode=2, chunk=0),
    dict(cfg=dict(levels=4, cblk=(32, 32), layers=2, npasses=3, eph=True), mode=1, chunk=41),
    dict(cfg=dict(levels=4, cblk=(32, 32), layers=2, npasses=3), mode=2, chunk=41),
    dict(cfg=dict(levels=3, precincts=[(5, 5), (6, 6)], cblk=(16, 16), prog=2, tile=(96, 96)), mode=1, chunk=7),
    dict(cfg=dict(levels=3, precincts=[(5, 5), (6, 6)], cblk=(16, 16), prog=4, tile=(96, 96)), mode=2, chunk=300),
    dict(cfg=dict(levels=2, reversible=False, base_step=1.0, tile=(100, 60), sop=True), mode=1, chunk=0),
    dict(cfg=dict(levels=2, reversible=False, base_step=1.0, poc=[(0, 1, 1, 3, 3, 1), (0, 0, 1, 3, 1, 0)]), mode=2, chunk=64),
]


def packed_stream(t, w=200, h=120, nc=3, bd=10, seed=6):
    """-> (the stream with its packet headers in PPT / PPM marker segments, the same picture as a plain stream)"""
    planes = j2kgen.synth_image(w, h, nc, bd, seed=seed)
    base = j2kgen.encode(planes, bd, j2kgen.EncParams(**t["cfg"]))
    cs = j2kgen.encode(planes, bd, j2kgen.EncParams(packed_headers=t["mode"], ppx_chunk=t["chunk"], **t["cfg"]))
    return cs, base


@pytest.mark.parametrize("t", PACKED, ids=lambda t: f"{'ppt' if t['mode'] == 1 else 'ppm'}-{t['chunk']}-" + "-".join(
    f"{k}{v}" for k, v in t["cfg"].items())[:40])
def test_packed_header_streams_bit_exact(t):
    """PPT / PPM (jpeg2000dec.c:958-1014, :1099-1108, :1147-1153, :1474-1480, :2467-2476): the reference decodes
    the stream with packed headers to the pixels of the plain one, the oracle agrees."""
    cs, base = packed_stream(t)
    assert (b"\xff\x61" if t["mode"] == 1 else b"\xff\x60") in cs[:cs.index(b"\xff\x93")]
    for bitexact in ((0,) if t["cfg"].get("reversible", True) else (0, 1)):
        r0, _, out0, _ = ref_decode(base, bitexact)
        r1, info1, out1, name = ref_decode(cs, bitexact)
        r2, info2, out2 = oracle_decode(cs, bitexact)
        assert r0 == 0 and r1 == 0 and r2 == 0
        assert np.array_equal(out0, out1)
        assert info1[:2] == info2[:2] and oracle_pix_name() == name and np.array_equal(out1, out2)


def packed_marker_cases():
    """Marker-level damage: PPT in a HOMOGENEOUS stream (Ccap15 bit 11 clear), PPT next to PPM, a PPM run longer than
    the data, a truncated PPM / PPT payload (headers run dry), PPM inside a tile-part header, too short segments."""
    planes = j2kgen.synth_image(96, 64, 3, 8, seed=5)
    kw = dict(levels=2, tile=(64, 64), eph=True)
    ppt = j2kgen.encode(planes, 8, j2kgen.EncParams(packed_headers=1, **kw))
    ppm = j2kgen.encode(planes, 8, j2kgen.EncParams(packed_headers=2, **kw))
    cases = []
    cap = ppt.index(b"\xff\x50")
    b = bytearray(ppt); b[cap + 8] &= 0xF7; cases.append(bytes(b))                  # Ccap15 bit 11 cleared
    at = ppm.index(b"\xff\x60")
    n = int.from_bytes(ppm[at + 2:at + 4], "big")
    seg = ppm[at:at + 2 + n]
    sot = ppt.index(b"\xff\x90")
    cases.append(ppt[:sot] + seg + ppt[sot:])                                        # PPM in the main header + PPT
    b = bytearray(ppm); b[at + 5:at + 9] = (1 << 20).to_bytes(4, "big"); cases.append(bytes(b))   # Nppm too large
    cases.append(ppm[:at + 2] + (n - 9).to_bytes(2, "big") + ppm[at + 4:at + 2 + n - 9] + ppm[at + 2 + n:])   # PPM cut short
    pt = ppt.index(b"\xff\x61")
    m = int.from_bytes(ppt[pt + 2:pt + 4], "big")
    cases.append(ppt[:pt + 2] + (m - 5).to_bytes(2, "big") + ppt[pt + 4:pt + 2 + m - 5] + ppt[pt + 2 + m:])   # PPT cut short,
    #   Psot now 5 too large: the next SOT is missed, tile 1 never gets a first tile-part (init_tile fails, :1054)
    for cut in (2, 5, 20):                                                           # PPT cut short, Psot corrected:
        b = bytearray(ppt[:pt + 2] + (m - cut).to_bytes(2, "big") + ppt[pt + 4:pt + 2 + m - cut] + ppt[pt + 2 + m:])
        so = b.index(b"\xff\x90")                                                    # the last headers read as zeros
        b[so + 6:so + 10] = (int.from_bytes(b[so + 6:so + 10], "big") - cut).to_bytes(4, "big")
        cases.append(bytes(b))
    sod = ppm.index(b"\xff\x93")
    cases.append(ppm[:sod] + seg + ppm[sod:])                                        # PPM in a tile-part header
    cases.append(ppm[:at] + b"\xff\x60\x00\x02" + ppm[at:])                         # Lppm = 2
    cases.append(ppt[:pt] + b"\xff\x61\x00\x02" + ppt[pt:])                         # Lppt = 2
    cases.append(ppm[:at] + b"\xff\x60\x00\x03\x07" + ppm[at:])                     # empty PPM segment in front
    b = bytearray(ppm); b[at + 5:at + 9] = (0).to_bytes(4, "big"); cases.append(bytes(b))         # Nppm = 0
    return cases


def test_packed_header_marker_validation_matches_reference():
    cases = packed_marker_cases()
    verdicts = []
    for bad in cases:
        r1, _, out1, _ = ref_decode(bad, 0)
        r2, _, out2 = oracle_decode(bad, 0)
        verdicts.append(r1)
        assert (r1 == 0) == (r2 == 0), (len(verdicts), r1, r2)
        if r1 == 0:
            assert np.array_equal(out1, out2), len(verdicts)
    assert sum(v != 0 for v in verdicts) >= 4 and sum(v == 0 for v in verdicts) >= 2, verdicts


# several tile-parts per tile (a new one every `tp` packets, optionally interleaved across tiles), alone and with
# packed headers.  mode 2 (PPM) with more than one tile-part per tile runs into a reference quirk: select_header
# (jpeg2000dec.c:1099-1108) continues in the NEXT tile-part's *data* when a tile-part's PPM run is used up, so the
# reference does not decode such streams to the source picture; oracle and product follow it bit for bit.
TILE_PARTS = [
    dict(cfg=dict(levels=3, cblk=(32, 32), tile=(128, 64), eph=True, sop=True, layers=2, npasses=3), mode=0, tp=1, il=False),
    dict(cfg=dict(levels=3, cblk=(32, 32), tile=(128, 64), eph=True, sop=True, layers=2, npasses=3), mode=1, tp=3, il=True),
    dict(cfg=dict(levels=3, cblk=(16, 16), prog=2, precincts=[(5, 5), (6, 6)], reversible=False), mode=0, tp=7, il=False),
    dict(cfg=dict(levels=3, cblk=(16, 16), prog=2, precincts=[(5, 5), (6, 6)], reversible=False), mode=1, tp=1, il=False),
    dict(cfg=dict(levels=4, prog=4, tile=(100, 100), cblk=(16, 16)), mode=0, tp=1, il=True),
    dict(cfg=dict(levels=4, prog=4, tile=(100, 100), cblk=(16, 16)), mode=1, tp=5, il=True),
    dict(cfg=dict(levels=3, cblk=(32, 32), tile=(128, 64), eph=True, layers=2), mode=2, tp=1, il=False),
    dict(cfg=dict(levels=3, cblk=(16, 16), prog=2, precincts=[(5, 5), (6, 6)]), mode=2, tp=3, il=True),
    dict(cfg=dict(levels=4, prog=4, tile=(100, 100), cblk=(16, 16)), mode=2, tp=7, il=False),
]


def tile_part_stream(t, w=200, h=120, nc=3, bd=10, seed=3):
    planes = j2kgen.synth_image(w, h, nc, bd, seed=seed)
    base = j2kgen.encode(planes, bd, j2kgen.EncParams(**t["cfg"]))
    cs = j2kgen.encode(planes, bd, j2kgen.EncParams(packed_headers=t["mode"], ppx_chunk=53 if t["mode"] else 0,
                                                    tp_packets=t["tp"], tp_interleave=t["il"], **t["cfg"]))
    return cs, base


@pytest.mark.parametrize("t", TILE_PARTS, ids=lambda t: f"mode{t['mode']}-tp{t['tp']}-il{int(t['il'])}-" + "-".join(
    f"{k}{v}" for k, v in t["cfg"].items())[:30])
def test_tile_part_streams_match_reference(t):
    cs, base = tile_part_stream(t)
    assert cs.count(b"\xff\x90\x00\x0a") > base.count(b"\xff\x90\x00\x0a")
    for bitexact in ((0,) if t["cfg"].get("reversible", True) else (0, 1)):
        r0, _, out0, _ = ref_decode(base, bitexact)
        r1, info1, out1, name = ref_decode(cs, bitexact)
        r2, info2, out2 = oracle_decode(cs, bitexact)
        assert r0 == 0 and (r1 == 0) == (r2 == 0), (r1, r2)
        if t["mode"] != 2:
            assert r1 == 0 and np.array_equal(out0, out1)
        if r1 == 0:
            assert info1[:2] == info2[:2] and np.array_equal(out1, out2)


# coding / quantisation / progression styles of single tiles: COD / QCD / POC segments in the first tile-part header
# of every odd tile (Ccap15 bit 11 set: not the HOMOGENEOUS set).  var: 1 progression, 2 one level less, 4 narrower
# code-blocks, 8 guard bit / step, 16 tile POC
TILE_STYLES = [
    dict(cfg=dict(levels=3, cblk=(32, 32), tile=(64, 48)), var=1),
    dict(cfg=dict(levels=3, cblk=(32, 32), tile=(64, 48)), var=2),
    dict(cfg=dict(levels=3, cblk=(32, 32), tile=(64, 48)), var=4),
    dict(cfg=dict(levels=3, cblk=(32, 32), tile=(64, 48)), var=8),
    dict(cfg=dict(levels=3, cblk=(32, 32), tile=(64, 48)), var=16),
    dict(cfg=dict(levels=3, cblk=(32, 32), tile=(64, 48), packed_headers=1, ppx_chunk=61), var=31),
    dict(cfg=dict(levels=3, cblk=(32, 32), tile=(64, 48), reversible=False, base_step=1.0), var=12),
    dict(cfg=dict(levels=3, cblk=(32, 32), tile=(64, 48), reversible=False, base_step=1.0, tp_packets=3, tp_interleave=True), var=31),
    dict(cfg=dict(levels=4, cblk=(16, 16), tile=(100, 60), precincts=[(5, 5), (6, 6)], prog=1, layers=2, npasses=3, eph=True), var=19),
    dict(cfg=dict(levels=4, cblk=(16, 16), tile=(100, 60), precincts=[(5, 5), (6, 6)], prog=3, sop=True), var=7),
]


def tile_style_stream(t, w=200, h=120, nc=3, bd=10, seed=3):
    planes = j2kgen.synth_image(w, h, nc, bd, seed=seed)
    return j2kgen.encode(planes, bd, j2kgen.EncParams(tile_var=t["var"], **t["cfg"])), planes


@pytest.mark.parametrize("t", TILE_STYLES, ids=lambda t: f"var{t['var']}-" + "-".join(f"{k}{v}" for k, v in t["cfg"].items())[:40])
def test_tile_style_streams_bit_exact(t):
    """COD / QCD / POC in tile-part headers (get_sot's copy of the main styles, then get_cod / get_qcd / get_poc on the
    tile's copies, jpeg2000dec.c:862-869, :571-759, :760-818): reference == oracle; lossless ones give back the source."""
    cs, planes = tile_style_stream(t)
    assert cs.count(b"\xff\x52") + cs.count(b"\xff\x5c") + cs.count(b"\xff\x5f") > 2
    for bitexact in ((0,) if t["cfg"].get("reversible", True) else (0, 1)):
        r1, info1, out1, name = ref_decode(cs, bitexact)
        r2, info2, out2 = oracle_decode(cs, bitexact)
        assert r1 == 0 and r2 == 0
        assert info1[:2] == info2[:2] and oracle_pix_name() == name and np.array_equal(out1, out2)
        if t["cfg"].get("reversible", True) and t["cfg"].get("npasses", 1) == 1:
            assert np.array_equal(out1.view(np.uint16).reshape(120, 200, 3) >> 6, np.stack(planes, -1))


def tile_style_marker_cases():
    """Tile-part COD in a HOMOGENEOUS stream (Ccap15 bit 11 clear), a tile RGN (first and later tile-part), a tile QCC
    and COC, a second tile POC segment, a damaged tile COD."""
    planes = j2kgen.synth_image(128, 96, 3, 8, seed=4)
    kw = dict(levels=2, cblk=(32, 32), tile=(64, 48))
    cs = j2kgen.encode(planes, 8, j2kgen.EncParams(tile_var=17, **kw))

    def insert_in_tile_header(data, tile_sot_index, segment):
        sots = [i for i in range(len(data) - 3) if data[i:i + 4] == b"\xff\x90\x00\x0a"]
        at = sots[tile_sot_index]
        psot = int.from_bytes(data[at + 6:at + 10], "big") + len(segment)
        return data[:at + 6] + psot.to_bytes(4, "big") + data[at + 10:at + 12] + segment + data[at + 12:]

    cases = []
    cap = cs.index(b"\xff\x50")
    b = bytearray(cs); b[cap + 8] &= 0xF7; cases.append(bytes(b))                           # bit 11 cleared: INVALIDDATA
    cases.append(insert_in_tile_header(cs, 1, b"\xff\x5e\x00\x05\x01\x00\x03"))            # RGN in the RGNFREE set (bit 12 clear)
    b = bytearray(cs); b[cap + 8] |= 0x10; cs = bytes(b)                                    # from here on: bit 12 set
    cases.append(insert_in_tile_header(cs, 1, b"\xff\x5e\x00\x05\x01\x00\x03"))            # RGN, component 1, shift 3
    cases.append(insert_in_tile_header(cs, 2, b"\xff\x5e\x00\x05\x00\x00\x02"))            # RGN, component 0 (quirk Q6)
    cases.append(insert_in_tile_header(cs, 0, b"\xff\x5e\x00\x05\x00\x00\x1f"))            # shift 31: PATCHWELCOME
    qcd = cs.index(b"\xff\x5c")
    n = int.from_bytes(cs[qcd + 2:qcd + 4], "big")
    qcc = b"\xff\x5d" + (n + 1).to_bytes(2, "big") + b"\x02" + cs[qcd + 4:qcd + 2 + n]
    cases.append(insert_in_tile_header(cs, 3, qcc))                                         # tile QCC for component 2
    cod = cs.index(b"\xff\x52")
    coc = b"\xff\x53\x00\x09\x01" + cs[cod + 4:cod + 5] + cs[cod + 9:cod + 14]
    cases.append(insert_in_tile_header(cs, 2, coc))                                         # tile COC for component 1
    poc = b"\xff\x5f\x00\x09\x00\x00\x00\x01\x03\x03\x02"
    cases.append(insert_in_tile_header(cs, 1, poc))                                         # tile POC in front of the tile's own
    tile_cod = cs.index(b"\xff\x52", cs.index(b"\xff\x90"))
    b = bytearray(cs); b[tile_cod + 9] = 9; cases.append(bytes(b))                          # tile COD: code-block width 2^11
    multi = bytearray(j2kgen.encode(planes, 8, j2kgen.EncParams(tile_var=1, tp_packets=4, **kw)))
    multi[multi.index(b"\xff\x50") + 8] |= 0x10
    multi = bytes(multi)
    cases.append(insert_in_tile_header(multi, 1, b"\xff\x5e\x00\x05\x00\x00\x02"))         # RGN in a later tile-part
    return cases


def test_tile_style_marker_validation_matches_reference():
    verdicts = []
    for bad in tile_style_marker_cases():
        r1, _, out1, _ = ref_decode(bad, 0)
        r2, _, out2 = oracle_decode(bad, 0)
        verdicts.append(r1)
        assert (r1 == 0) == (r2 == 0), (len(verdicts), r1, r2)
        if r1 == 0:
            assert np.array_equal(out1, out2), len(verdicts)
    assert sum(v != 0 for v in verdicts) >= 3 and sum(v == 0 for v in verdicts) >= 4, verdicts


def psot_bound_cases():
    """get_sot's bound (jpeg2000dec.c:845-850): a tile-part must end two bytes before the end of the packet -- a stream
    without its EOC marker is INVALIDDATA -- and Psot = 0 means "up to the EOC marker"."""
    planes = j2kgen.synth_image(96, 64, 3, 8, seed=5)
    cs = j2kgen.encode(planes, 8, j2kgen.EncParams(levels=2, cblk=(32, 32), tile=(64, 64)))
    last = cs.rindex(b"\xff\x90\x00\x0a")
    zero = cs[:last + 6] + b"\0\0\0\0" + cs[last + 10:]
    return [cs[:-2], cs[:-1], cs + b"\0\0", zero, zero[:-2], zero + b"\xff\xd9", zero[:-2] + b"\x12\x34"]


def test_psot_bound_matches_reference():
    verdicts = []
    for i, c in enumerate(psot_bound_cases()):
        r1, info1, out1, _ = ref_decode(c, 0)
        r2, info2, out2 = oracle_decode(c, 0)
        verdicts.append(r1)
        assert r1 == r2, (i, r1, r2)
        if r1 == 0:
            assert np.array_equal(out1, out2), i
    assert verdicts[0] < 0 and verdicts[2] == 0 and verdicts[3] == 0, verdicts


# HT sets other than HTONLY (Ccap15 bits 14-15, SPcod bits 6-7; jpeg2000dec.c:443-459, :1060-1063, :1256-1340).
# ht_set: 1 HTDECLARED, 3 MIXED with every block written so that it reads as HT, 4 MIXED with HTONLY-style cleanup
# lengths -- the reference takes blocks whose length field has Lblock == 3 or its top bit set for Part-1 blocks and
# runs the MQ decoder over their HT bytes (no picture, but a deterministic answer) --, 5 the same under an HTONLY Ccap15
HT_SETS = [dict(cfg=c, ht_set=hs) for hs in (1, 3, 4, 5) for c in (
    dict(levels=3, cblk=(32, 32)),
    dict(levels=3, cblk=(16, 16), npasses=3, reversible=False, base_step=1.0),
    dict(levels=4, cblk=(32, 32), layers=2, npasses=3, tile=(100, 60)),
    dict(levels=2, cblk=(8, 8), prog=2, precincts=[(5, 5), (6, 6), (6, 6)], sop=True, eph=True),
    dict(levels=3, cblk=(64, 64), vsc=True, npasses=2),
)]
HT_SET_IDS = [f"set{t['ht_set']}-" + "-".join(f"{k}{v}" for k, v in t["cfg"].items())[:36] for t in HT_SETS]


def ht_set_stream(t, w=200, h=120, nc=3, bd=10, seed=3):
    planes = j2kgen.synth_image(w, h, nc, bd, seed=seed)
    return j2kgen.encode(planes, bd, j2kgen.EncParams(ht_set=t["ht_set"], **t["cfg"])), planes


@pytest.mark.parametrize("t", HT_SETS, ids=HT_SET_IDS)
def test_ht_set_streams_match_reference(t):
    """HTDECLARED and MIXED streams: reference == oracle (verdict, format, pixels); with the stripe-causal style bit
    the HT blocks run their SigProp pass causally and the Part-1 blocks form their contexts causally."""
    cs, planes = ht_set_stream(t)
    for bitexact in ((0,) if t["cfg"].get("reversible", True) else (0, 1)):
        r1, info1, out1, name = ref_decode(cs, bitexact)
        r2, info2, out2 = oracle_decode(cs, bitexact)
        assert r1 == 0
        assert r2 == 0 and info1[:2] == info2[:2] and oracle_pix_name() == name and np.array_equal(out1, out2)
        if t["ht_set"] <= 3 and t["cfg"].get("reversible", True) and t["cfg"].get("npasses", 1) == 1:
            assert np.array_equal(out1.view(np.uint16).reshape(120, 200, 3) >> 6, np.stack(planes, -1))
        if t["ht_set"] >= 4 and t["cfg"].get("npasses", 1) == 1:       # some blocks did go through the MQ decoder
            assert not np.array_equal(out1.view(np.uint16).reshape(120, 200, 3) >> 6, np.stack(planes, -1))


def ht_set_damaged_streams(n_per=10):
    rng = np.random.default_rng(91)
    for t in (HT_SETS[5], HT_SETS[10], HT_SETS[11], HT_SETS[13], HT_SETS[15]):
        cs, _ = ht_set_stream(t)
        sod = cs.index(b"\xff\x93")
        for k in range(n_per):
            bad = bytearray(cs)
            for _ in range(int(rng.integers(1, 6))):
                bad[int(rng.integers(sod + 2, len(bad) - 2))] = int(rng.integers(0, 256))
            if k % 4 == 3:
                bad = bad[:int(rng.integers(sod + 10, len(bad)))]
            yield bytes(bad)


def test_ht_set_damaged_streams_match_reference():
    """Byte damage inside the packets of MIXED streams: blocks flip between the two block coders, lengths go wrong."""
    n_ok = 0
    for i, bad in enumerate(ht_set_damaged_streams()):
        r1, info1, out1, _ = ref_decode(bad, 0)
        r2, info2, out2 = oracle_decode(bad, 0)
        assert (r1 == 0) == (r2 == 0), (i, r1, r2)
        if r1 == 0:
            assert np.array_equal(out1, out2), i
            n_ok += 1
    assert n_ok >= 20


# ---- Part-1 (EBCOT / MQ) code-blocks: SURVEY 8(f) row N4 ----------------------------------------------------
# The reference has no HT encoder and this repo's test encoder writes HT only, so Part-1 streams come from an
# independent encoder: OpenJPEG through Pillow (code-block style 0: no bypass / restart / reset / causal / segsym).
PART1 = [
    dict(mode="RGB", w=100, h=80, kw=dict(irreversible=False)),
    dict(mode="RGB", w=100, h=80, kw=dict(irreversible=True, quality_mode="rates", quality_layers=[8])),
    dict(mode="RGB", w=131, h=77, kw=dict(irreversible=False, num_resolutions=3, codeblock_size=(32, 32), progression="RPCL",
                                          precinct_size=(64, 64))),
    dict(mode="RGB", w=200, h=120, kw=dict(irreversible=False, quality_mode="rates", quality_layers=[20, 8, 1],
                                           tile_size=(64, 64))),
    dict(mode="RGB", w=200, h=120, kw=dict(irreversible=True, quality_mode="rates", quality_layers=[40, 10], progression="CPRL",
                                           tile_size=(128, 96), codeblock_size=(16, 64))),
    dict(mode="L", w=97, h=61, kw=dict(irreversible=False, num_resolutions=1)),
    dict(mode="L", w=160, h=90, kw=dict(irreversible=True, quality_mode="dB", quality_layers=[38], progression="PCRL")),
    dict(mode="I;16", w=150, h=100, kw=dict(irreversible=False, codeblock_size=(64, 64))),
    dict(mode="I;16", w=150, h=100, kw=dict(irreversible=True, quality_mode="rates", quality_layers=[6], progression="RLCP")),
    dict(mode="RGBA", w=120, h=70, kw=dict(irreversible=False, mct=1, codeblock_size=(4, 4), num_resolutions=2)),
    dict(mode="RGB", w=180, h=100, kw=dict(irreversible=False, offset=(13, 7), tile_offset=(5, 3), tile_size=(80, 60))),
    dict(mode="LA", w=64, h=64, kw=dict(irreversible=False)),
    dict(mode="RGB", w=256, h=200, kw=dict(irreversible=False, codeblock_size=(1024, 4), num_resolutions=2)),
]


def part1_stream(t, jp2=False, seed=9):
    """A Part-1 codestream (or JP2 file) of synthetic content from OpenJPEG via Pillow."""
    import io

    from PIL import Image
    nc = {"L": 1, "I;16": 1, "LA": 2, "RGB": 3, "RGBA": 4}[t["mode"]]
    bd = 16 if t["mode"] == "I;16" else 8
    planes = j2kgen.synth_image(t["w"], t["h"], nc, bd, seed=seed)
    if t["mode"] == "I;16":
        im = Image.fromarray(planes[0].astype(np.uint16))
    else:
        a = np.stack(planes, -1).astype(np.uint8)
        im = Image.fromarray(a[..., 0] if nc == 1 else a, t["mode"])
    b = io.BytesIO()
    im.save(b, "JPEG2000", no_jp2=not jp2, **t["kw"])
    return b.getvalue()


@pytest.mark.parametrize("t", PART1, ids=lambda t: f"{t['mode']}-{t['w']}x{t['h']}-" + "-".join(
    f"{k[:4]}{v}" for k, v in t["kw"].items())[:50])
def test_part1_streams_bit_exact(t):
    """MQ decoder, context formation, the three coding passes, T2 lengths of non-HT blocks
    (jpeg2000dec.c:1364-1368, :1872-2089; mqcdec.c): reference == oracle, float and fixed-point 9/7."""
    for jp2 in (False, True):
        cs = part1_stream(t, jp2)
        for bitexact in ((0,) if not t["kw"].get("irreversible") else (0, 1)):
            r1, info1, out1, name = ref_decode(cs, bitexact)
            r2, info2, out2 = oracle_decode(cs, bitexact)
            assert r1 == 0 and r2 == 0, (r1, r2)
            assert info1[:2] == info2[:2] and oracle_pix_name() == name and np.array_equal(out1, out2)


def test_part1_damaged_streams_match_reference():
    """Byte damage inside the packets: the MQ decoder runs into the 0xFF 0xFF sentinel, lengths go wrong, the
    bit-plane counter may leave its range -- same verdict and same pixels as the reference."""
    rng = np.random.default_rng(77)
    n_ok = 0
    for t in (PART1[0], PART1[3], PART1[7], PART1[4]):
        cs = part1_stream(t)
        sod = cs.index(b"\xff\x93")
        for k in range(12):
            bad = bytearray(cs)
            for _ in range(int(rng.integers(1, 6))):
                bad[int(rng.integers(sod + 2, len(bad) - 2))] = int(rng.integers(0, 256))
            if k % 4 == 3:
                bad = bad[:int(rng.integers(sod + 10, len(bad)))]
            bad = bytes(bad)
            r1, info1, out1, _ = ref_decode(bad, 0)
            r2, info2, out2 = oracle_decode(bad, 0)
            assert (r1 == 0) == (r2 == 0), (t["mode"], k, r1, r2)
            if r1 == 0:
                assert np.array_equal(out1, out2), (t["mode"], k)
                n_ok += 1
    assert n_ok >= 20


def part1_styled_streams():
    """OpenJPEG writes code-block style 0 only and nothing else in the container writes Part-1 blocks, so the other
    styles are reached by changing SPcod of a style-0 stream: the same bytes decoded under context reset (0x02),
    vertically causal contexts (0x08), predictable termination (0x10), segmentation symbols (0x20) and combinations
    give no picture, but every decision of the block decoder goes through the changed rules -- a deterministic
    answer to agree on.  Bypass (0x01) and termination on each pass (0x04) change the tier-2 lengths: refused here."""
    for ti in (0, 1, 3, 4, 7, 9, 10):
        cs = part1_stream(PART1[ti])
        cod = cs.index(b"\xff\x52")
        for style in (0x02, 0x08, 0x20, 0x10, 0x2A, 0x3A, 0x01, 0x04):
            b = bytearray(cs)
            b[cod + 12] = style
            yield f"part1[{ti}] style {style:#04x}", bytes(b)


def test_part1_styled_streams_match_reference():
    n_ok = n_refused = 0
    plain = {}
    for name, cs in part1_styled_streams():
        r1, info1, out1, pix = ref_decode(cs, 0)
        r2, info2, out2 = oracle_decode(cs, 0)
        if int(name[-2:], 16) & 0x05:
            assert r2 == -1163346256, name                # the reference itself may decode or fail here
            n_refused += 1
            continue
        assert r1 == 0 and r2 == 0 and oracle_pix_name() == pix and np.array_equal(out1, out2), name
        plain.setdefault(name[:8], out1)
        n_ok += 1
    assert n_ok == 42 and n_refused == 14


# chroma-subsampled streams (XRsiz / YRsiz > 1): (name, [(XRsiz, YRsiz)], bit depth) -> the planar YUV(A) formats
SUBSAMPLED = [("420", [(1, 1), (2, 2), (2, 2)], 8), ("422", [(1, 1), (2, 1), (2, 1)], 10), ("440", [(1, 1), (1, 2), (1, 2)], 8),
              ("411", [(1, 1), (4, 1), (4, 1)], 8), ("410", [(1, 1), (4, 4), (4, 4)], 8),
              ("420a", [(1, 1), (2, 2), (2, 2), (1, 1)], 8), ("420-12", [(1, 1), (2, 2), (2, 2)], 12),
              ("422-16", [(1, 1), (2, 1), (2, 1)], 16), ("420-9", [(1, 1), (2, 2), (2, 2)], 9)]
SUBSAMPLED_GEOMS = [(96, 64, dict(levels=2)), (131, 77, dict(levels=3, tile=(64, 48))),
                    (200, 120, dict(levels=4, reversible=False, base_step=1.0, prog=1))]


def subsampled_stream(subs, bd, w, h, kw, seed=1):
    """-> (codestream, planes): every component on its own grid, no component transform"""
    full = j2kgen.synth_image(w, h, len(subs), bd, seed=seed)
    planes = [np.ascontiguousarray(a[::sy, ::sx]) for a, (sx, sy) in zip(full, subs)]
    p = j2kgen.EncParams(mct=False, subsampling=subs, **kw)
    return j2kgen.encode(planes, bd, p), planes, p


@pytest.mark.parametrize("name,subs,bd", SUBSAMPLED, ids=[s[0] for s in SUBSAMPLED])
def test_subsampled_decode_bit_exact(name, subs, bd):
    """Planar output: plane sizes follow the AVPixelFormat's chroma shift, 9..16-bit samples are not
    left-aligned (write_frame_16 with precision = s->precision): oracle == reference, format name included."""
    for w, h, kw in SUBSAMPLED_GEOMS:
        cs, planes, p = subsampled_stream(subs, bd, w, h, kw)
        for bitexact in ((0,) if p.reversible else (0, 1)):
            r1, info1, out1, pix = ref_decode(cs, bitexact)
            r2, info2, out2 = oracle_decode(cs, bitexact)
            assert r1 == 0 and r2 == 0
            assert info1[:2] == info2[:2] and oracle_pix_name() == pix
            assert np.array_equal(out1, out2)
        if p.reversible:
            dt = np.uint16 if bd > 8 else np.uint8
            flat, pos = out2.view(dt), 0
            for a in planes:
                assert np.array_equal(flat[pos: pos + a.size].reshape(a.shape), a)
                pos += a.size


def test_jp2_wrapped_decode_bit_exact():
    """JP2 / JPH files: the codestream behind the boxes, the colour specification's pixel-format list (planar
    YUV for sYCC), channel definitions that remap components to planes, a palette (PAL8 for one component,
    only steering the format choice otherwise), the ten-box search limit -- oracle == reference in dimensions, AVPixelFormat and pixels."""
    n = 0
    for nc, bd, sub in [(3, 8, None), (3, 12, None), (1, 8, None), (1, 16, None), (2, 8, None), (4, 8, None), (3, 10, None),
                        (4, 16, None), (3, 8, "420"), (3, 10, "422")]:
        planes = j2kgen.synth_image(96, 64, nc, bd, seed=3)
        cs = j2kgen.encode(planes, bd, j2kgen.EncParams(levels=2, mct=nc >= 3))
        variants = [j2kgen.wrap_jp2(cs, 96, 64, nc, bd, colour_space=csp) for csp in (None, 16, 17, 18, 99)]
        variants.append(j2kgen.wrap_jp2(cs, 96, 64, nc, bd, colour_space=16, palette=(16, 8)))
        variants.append(j2kgen.wrap_jp2(cs, 96, 64, nc, bd, colour_space=18, cdef=[(0, 0, 2), (1, 0, 1), (2, 0, 3), (3, 1, 0)][:nc]))
        variants.append(j2kgen.wrap_jp2(cs, 96, 64, nc, bd, colour_space=18, cdef=[(0, 0, 3), (2, 0, 1)][:nc]))
        variants.append(j2kgen.wrap_jp2(cs, 96, 64, nc, bd, extra_before=11))
        variants.append(j2kgen.wrap_jp2(cs, 96, 64, nc, bd, extra_before=8))
        for data in variants:
            r1, info1, out1, name = ref_decode(data)
            r2, info2, out2 = oracle_decode(data)
            assert r1 == r2, (nc, bd, r1, r2)
            if r1 == 0:
                assert info1[:2] == info2[:2] and oracle_pix_name() == name
                assert np.array_equal(out1, out2), (nc, bd, name)
            n += 1
    assert n > 80


def dcinema_cases():
    """Rsiz 3 / 4 (AV_PROFILE_JPEG2000_DCINEMA_2K / _4K): get_siz walks the XYZ list -- XYZ12 for three components of
    9..12 bits, then the planar YUV formats -- whatever a JP2 colour specification says (jpeg2000dec.c:330-333); the
    samples are written like RGB48, << (16 - cbps) (:2386-2391)."""
    def with_rsiz(cs, r):
        siz = cs.index(b"\xff\x51")
        return cs[:siz + 4] + r.to_bytes(2, "big") + cs[siz + 6:]
    for nc, bd, kw in ((3, 12, dict(levels=3)), (3, 12, dict(levels=3, reversible=False, base_step=1.0, tile=(64, 64))),
                       (3, 10, dict(levels=2)), (3, 16, dict(levels=2)), (1, 12, dict(levels=2)), (3, 8, dict(levels=2)),
                       (4, 12, dict(levels=2, mct=False)), (3, 9, dict(levels=2, npasses=3))):
        planes = j2kgen.synth_image(96, 64, nc, bd, seed=3)
        cs = j2kgen.encode(planes, bd, j2kgen.EncParams(**kw))
        yield f"{nc}x{bd} rsiz 3", with_rsiz(cs, 3)
        yield f"{nc}x{bd} rsiz 4", with_rsiz(cs, 4)
        yield f"{nc}x{bd} jp2 rsiz 3", j2kgen.wrap_jp2(with_rsiz(cs, 3), 96, 64, nc, bd, colour_space=16)
    yield "part1 16-bit grey rsiz 3", with_rsiz(part1_stream(PART1[7]), 3)
    yield "part1 rgb 8-bit rsiz 4", with_rsiz(part1_stream(PART1[0]), 4)


def test_dcinema_profile_formats_match_reference():
    names = set()
    for name, cs in dcinema_cases():
        r1, info1, out1, pix = ref_decode(cs)
        r2, info2, out2 = oracle_decode(cs)
        if r1 == 0 and pix == "xyz12le" and ("x8" in name or "8-bit" in name):
            assert r2 == -1163346256, name     # bytes written into half of a 16-bit frame: refused, not restated
            continue
        assert r1 == r2, name
        if r1 == 0:
            assert info1[:2] == info2[:2] and oracle_pix_name() == pix and np.array_equal(out1, out2), name
            names.add(pix)
    assert {"xyz12le", "yuv444p16le", "gray16le"} <= names


# capture / display resolution boxes: (vnum, vden, hnum, hden, vexp, hexp) -> AVCodecContext.sample_aspect_ratio
RES_BOXES = [(300, 1, 300, 1, 0, 0), (300, 1, 150, 1, 0, 0), (72, 1, 96, 1, 2, 1), (65535, 3, 7, 65535, 9, 0), (1, 1, 1, 1, 0, 18),
             (1, 1, 1, 1, 0, 19), (5, 0, 5, 1, 0, 0), (65535, 1, 65535, 1, 200, 3), (1000, 7, 999, 13, 3, 5),
             (65535, 65534, 65533, 65532, 0, 12), (1, 65535, 65535, 1, 0, 0), (59, 7, 61, 11, 4, 4)]


def res_box_file(res):
    planes = j2kgen.synth_image(64, 48, 3, 8, seed=3)
    return j2kgen.wrap_jp2(j2kgen.encode(planes, 8, j2kgen.EncParams(levels=2)), 64, 48, 3, 8, res=res)


def test_sample_aspect_ratio_matches_reference():
    """'res ' box -> s->sar through av_reduce with the reference's double arithmetic and overflow guards
    (jpeg2000dec.c:2762-2796); published only when both terms are non-zero (:2867-2868)."""
    seen = set()
    for res in RES_BOXES:
        r1, info1, out1, _ = ref_decode(res_box_file(res))
        r2, info2, out2 = oracle_decode(res_box_file(res))
        assert r1 == 0 and r2 == 0 and np.array_equal(out1, out2)
        assert (info2[7] >> 32, info2[7] & 0xFFFFFFFF) == (info1[5], info1[6]), res
        seen.add((info1[5], info1[6]))
    assert (0, 0) in seen and (2, 1) in seen and (1, 2147483647) in seen and len(seen) >= 7


# palettised JP2 files: (bit depth, colour space, (entries, palette depth))
PAL8_CASES = [(8, 16, (200, 8)), (8, None, (256, 12)), (6, 16, (7, 4)), (8, 99, (256, 16)), (5, 16, (32, 1)), (8, 17, (16, 8)),
              (8, 16, (257, 8)), (1, 16, (2, 8)), (8, 18, (100, 9))]


def pal8_file(bd, csp, pal, w=96, h=64, **kw):
    planes = j2kgen.synth_image(w, h, 1, bd, seed=3)
    cs = j2kgen.encode(planes, bd, j2kgen.EncParams(levels=2, cblk=(32, 32), mct=False, **kw))
    return j2kgen.wrap_jp2(cs, w, h, 1, bd, colour_space=csp, palette=pal), planes


@pytest.mark.parametrize("bd,csp,pal", PAL8_CASES)
def test_pal8_files_match_reference(bd, csp, pal):
    """One component of at most 8 bits + a valid 'pclr' box outside the greyscale list: AV_PIX_FMT_PAL8 -- the index
    plane written like gray8 and 256 palette words in data[1] (jpeg2000dec.c:2711-2752, :2900-2901)."""
    for kw in (dict(), dict(tile=(64, 32))):
        data, planes = pal8_file(bd, csp, pal, **kw)
        r1, info1, out1, name = ref_decode(data)
        r2, info2, out2 = oracle_decode(data)
        assert r1 == 0 and r2 == 0 and info1[:2] == info2[:2] and oracle_pix_name() == name
        assert np.array_equal(out1, out2)
        if name == "pal8":
            assert out1.size == 96 * 64 + 1024
            assert np.array_equal(out1[:96 * 64].reshape(64, 96) >> (8 - bd), planes[0])
    assert (name == "pal8") == (csp not in (17, 18) and pal[0] <= 256)      # grey and sYCC lists hold no PAL8


def test_error_codes_match():
    """Truncated / damaged codestreams: both fail, or both succeed with identical pixels."""
    planes = j2kgen.synth_image(96, 80, 3, 12, seed=3)
    cs = j2kgen.encode(planes, 12, j2kgen.EncParams(levels=3, cblk=(32, 32)))
    rng = np.random.default_rng(5)
    for _ in range(60):
        bad = bytearray(cs)
        mode = rng.integers(0, 3)
        if mode == 0:
            bad = bad[: int(rng.integers(2, len(bad)))]
        elif mode == 1:
            for _ in range(3):
                bad[int(rng.integers(110, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        else:
            pos = int(rng.integers(110, len(bad) - 8))
            bad[pos:pos + 8] = rng.integers(0, 256, 8, dtype=np.uint8).tobytes()
        r1, _, out1, _ = ref_decode(bytes(bad))
        r2, _, out2 = oracle_decode(bytes(bad))
        assert (r1 == 0) == (r2 == 0), (mode, r1, r2)
        if r1 == 0:
            assert np.array_equal(out1, out2)
