"""Golden vectors produced by the unmodified reference (tests/golden/make_golden.py) vs the oracle.
CPU only; the GPU counterparts live in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from helpers import GOLDEN, ORACLE_PIX, ht_block, oracle, oracle_decode, oracle_pix_name


def _groups(npz, fields):
    keys = sorted({k.rsplit("|", 1)[0] for k in npz.files})
    for k in keys:
        yield k, [npz[f"{k}|{f}"] for f in fields]


def test_golden_frames():
    z = np.load(GOLDEN / "frames.npz")
    n = 0
    for key, (cs, meta, pix, out) in _groups(z, ["cs", "meta", "pix", "out"]):
        bitexact = int(key.split("|")[1])
        r, info, got = oracle_decode(cs.tobytes(), bitexact)
        assert r == 0, key
        assert info[0] == meta[0] and info[1] == meta[1], key
        assert ORACLE_PIX[info[2]] == pix.tobytes().decode(), key
        assert np.array_equal(got, out), key
        n += 1
    assert n >= 20


def test_golden_frames_wide():
    """lowres, JP2 / JPH files, chroma-subsampled streams: reference outputs and error codes"""
    z = np.load(GOLDEN / "frames_wide.npz")
    n = 0
    for key, (cs, meta, pix, out) in _groups(z, ["cs", "meta", "pix", "out"]):
        flags = int(key.split("|")[1])
        r, info, got = oracle_decode(cs.tobytes(), flags)
        assert r == meta[2], key
        if r == 0:
            assert info[0] == meta[0] and info[1] == meta[1], key
            assert oracle_pix_name() == pix.tobytes().decode(), key
            assert np.array_equal(got, out), key
        n += 1
    assert n >= 35


def test_golden_frames_packed():
    """PPT / PPM streams, damaged PPT / PPM markers, tiles without a first tile-part: reference verdicts + pixels"""
    z = np.load(GOLDEN / "frames_packed.npz")
    n = 0
    for key, (cs, meta, pix, out) in _groups(z, ["cs", "meta", "pix", "out"]):
        r, info, got = oracle_decode(cs.tobytes(), int(key.split("|")[1]))
        assert (r == 0) == (meta[2] == 0), key
        if r == 0:
            assert info[0] == meta[0] and info[1] == meta[1], key
            assert oracle_pix_name() == pix.tobytes().decode(), key
            assert np.array_equal(got, out), key
        n += 1
    assert n >= 25


@pytest.mark.parametrize("name", ["frames_part1.npz", "frames_tiles.npz", "frames_htsets.npz", "frames_fuzz.npz"])
def test_golden_frames_part1_and_tiles(name):
    """Part-1 (EBCOT / MQ) streams from OpenJPEG, plain and damaged; several tile-parts per tile, tiles with their own
    COD / QCD / POC, damaged tile-header markers; the findings of the differential fuzzer (tests/golden/
    make_fuzz_golden.py): reference verdicts + pixels"""
    import hashlib
    z = np.load(GOLDEN / name)
    n = 0
    for key, (cs, meta, pix, out) in _groups(z, ["cs", "meta", "pix", "out"]):
        r, info, got = oracle_decode(cs.tobytes(), int(key.split("|")[1]))
        assert (r == 0) == (meta[2] == 0), key
        if r == 0:
            assert info[0] == meta[0] and info[1] == meta[1], key
            assert oracle_pix_name() == pix.tobytes().decode(), key
            if out.size == 20 and got.size != 20:      # frames_tiles.npz keeps SHA-1 digests of the pictures
                assert hashlib.sha1(got.tobytes()).digest() == out.tobytes(), key
            else:
                assert np.array_equal(got, out), key
        n += 1
    assert n >= (28 if name == "frames_fuzz.npz" else 25)


def test_golden_kat_values():
    """The survey's hand-checked numbers on top of the byte comparison."""
    z = np.load(GOLDEN / "frames.npz")
    k1 = z["K1|0|out"].reshape(64, 64)
    assert k1[0, 0] == 116 and (k1 != 128).sum() == 1
    assert z["K2|0|out"].reshape(64, 64)[0, 0] == 114
    k3 = z["K3|0|out"].reshape(64, 64)
    assert k3[0, 0] == 114 and k3[1, 0] == 122
    assert z["K4|0|out"].reshape(64, 64)[0, 0] == 108
    k5 = z["K5|0|out"].reshape(64, 64)
    assert k5[:2, :3].tolist() == [[113, 120, 129], [120, 124, 128]]
    k6 = z["K6|0|out"].reshape(64, 64)
    assert k6[:2, :3].tolist() == [[112, 143, 121], [143, 113, 135]]
    k7 = z["K7|0|out"].view(np.uint16).reshape(64, 64, 3)
    assert k7[0, 0].tolist() == [34304, 34304, 28160] and k7[1, 1].tolist() == [33152, 33152, 31616]
    k8 = z["K8|0|out"].reshape(64, 128)
    assert k8[0, 64] == 116 and (k8 != 128).sum() == 1


def test_golden_blocks():
    z = np.load(GOLDEN / "blocks.npz")
    n = 0
    for key, (data, par, out) in _groups(z, ["data", "par", "out"]):
        npasses, Lcup, Lref, zbp, w, h, M_b, roi, vsc, ret = [int(v) for v in par]
        r, got = ht_block(oracle().oracle_ht_block, data.tobytes(), npasses, Lcup, Lref, zbp, w, h, M_b, roi, vsc)
        assert r == ret, key
        assert np.array_equal(got, out), key
        n += 1
    assert n == 120


def test_golden_idwt():
    z = np.load(GOLDEN / "idwt.npz")
    n = 0
    for key, (a, par, out) in _groups(z, ["in", "par", "out"]):
        t = int(key.split("|")[0])
        b = a.copy()
        border = np.ascontiguousarray(par[:4], np.int32)
        assert oracle().oracle_idwt(t, border.ctypes.data, int(par[4]), b.ctypes.data) == 0
        assert np.array_equal(b.view(np.uint32), out.view(np.uint32)), key
        n += 1
    assert n == 36
