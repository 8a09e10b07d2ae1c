"""Host side of the product (codestream -> flat plan, ffmpeg-ht_b200/csrc/j2k_parse.cc) checked on the
CPU: the plan's tables are fed through the ORACLE's stage functions (block decode, IDWT, MCT) with a
numpy restatement of dequantisation and packing, and the picture must equal the oracle's own decode of
the same codestream -- i.e. offsets, lengths, zbp, M_b, step sizes, Mallat placement, level bounds and
pack parameters are all what the reference would use.  No GPU."""
import ctypes

import numpy as np
import pytest

import ffmpeg_ht_b200 as fb
import j2kgen
from helpers import ROOT, ht_block, mq_block, oracle, oracle_decode, oracle_pix_name
from test_oracle_vs_ref import (dcinema_cases, RES_BOXES, res_box_file, part1_styled_streams, PAL8_CASES, pal8_file, psot_bound_cases, HT_SETS, HT_SET_IDS, ht_set_damaged_streams, ht_set_stream, LAYERED, PACKED, PART1, POC_STREAMS, STREAMS, SUBSAMPLED, SUBSAMPLED_GEOMS, TILE_PARTS,
                                TILE_STYLES, part1_stream, tile_part_stream, tile_style_marker_cases, tile_style_stream,
                                packed_marker_cases, packed_stream, poc_marker_cases, poc_stream,
                                subsampled_stream)


def picture_from_plan(cs: bytes, flags: int):
    P = fb.plan(cs, flags)
    info = P["info"]
    data = np.frombuffer(cs, np.uint8)
    allbytes = np.concatenate([data, P["extra"], np.zeros(64, np.uint8)])
    coef = np.zeros(int(P["coef_elems"]) + 64, np.uint32)
    for d in P["cblk"]:
        if int(d["flags"]) & 0x80:                  # J2K_CBLK_SKIP: resolution dropped by lowres
            continue
        w, h, M_b = int(d["w"]), int(d["h"]), int(d["M_b"])
        seg = allbytes[int(d["data_off"]): int(d["data_off"]) + int(d["length"])].tobytes()
        if int(d["flags"]) & 0x08:                  # J2K_CBLK_MQ: Lcup = nonzerobits, Lref = band orientation
            f = int(d["flags"])                            # descriptor flags -> SPcod style bits of a Part-1 block
            style = (0x08 if f & 1 else 0) | (0x02 if f & 0x10 else 0) | (0x20 if f & 0x20 else 0)
            ret, sm = mq_block(seg, int(d["npasses"]), int(d["Lcup"]), w, h, int(d["Lref"]), M_b, int(d["roi_shift"]), style)
        else:
            ret, sm = ht_block(oracle().oracle_ht_block, seg, int(d["npasses"]), int(d["Lcup"]), int(d["Lref"]),
                               int(d["zbp"]), w, h, M_b, int(d["roi_shift"]), int(d["flags"]) & 1)
        band = P["bands"][int(d["band"])]
        mode = (int(d["flags"]) >> 1) & 3
        sm = sm.astype(np.int64)
        mag = sm & 0x7FFFFFFF
        neg = sm < 0
        if ret == 0 or (ret < 0 and not int(d["flags"]) & 0x08):   # a failed MQ block keeps what it decoded
            vals = np.zeros((h, w), np.uint32)
        elif mode == 1:
            v = mag >> (31 - M_b)
            v = np.where(neg, -v, v)
            if int(band["i_stepsize"]) != 32768:
                v = np.trunc(v * int(band["i_stepsize"]) / 65536).astype(np.int64)
            vals = v.astype(np.int32).view(np.uint32)
        elif mode == 0:
            v = np.where(neg, -mag, mag).astype(np.int32).astype(np.float32)
            vals = (v * np.float32(band["fscale"])).astype(np.float32).view(np.uint32)
        else:
            v = np.where(neg, -mag, mag)
            v = (v + 32) >> 6
            m = v * int(band["iscale"])
            vals = np.where(m > 0, (m + 32768) >> 16, (m + 32767) >> 16).astype(np.int32).view(np.uint32)
        stride, off = int(d["dst_stride"]), int(d["dst_off"])
        for y in range(h):
            coef[off + y * stride: off + y * stride + w] = vals[y]
    planes = []
    for tc in P["tcs"]:
        n = int(tc["width"]) * int(tc["height"])
        pl = coef[int(tc["coef_off"]): int(tc["coef_off"]) + n].copy()
        lv = tc["lev"][int(tc["nlevels"])]
        border = np.array([lv["u0"], lv["u1"], lv["v0"], lv["v1"]], np.int32)
        assert border[1] - border[0] == tc["width"] and border[3] - border[2] == tc["height"]
        if n:
            assert oracle().oracle_idwt(int(tc["transform"]), border.ctypes.data, int(tc["nlevels"]), pl.ctypes.data) == 0
        planes.append(pl)
    pic = np.zeros(int(P["picture_bytes"]), np.uint8)
    if info.pix_name == "pal8":                            # plane 1 = the palette, filled in by the host
        pic[P["plane_off"][1]:] = P["palette"].view(np.uint8)
    rowb = [info.plane_row_bytes[i] for i in range(4)]
    for g in P["groups"]:
        tcs = [P["tcs"][int(g["tc"][c])] for c in range(int(g["ncomp"]))]
        pls = [planes[int(g["tc"][c])] for c in range(int(g["ncomp"]))]
        t = int(tcs[0]["transform"])
        if int(g["mct"]):
            oracle().oracle_mct(t, pls[0].ctypes.data, pls[1].ctypes.data, pls[2].ctypes.data, pls[0].size)
        wide = int(g["wide"])
        for c, (tc, pl) in enumerate(zip(tcs, pls)):
            plane = int(g["comp_plane"][c])
            view = pic[P["plane_off"][plane]:]
            W, H = int(tc["width"]), int(tc["height"])
            if t == 0:
                val = np.rint(pl.view(np.float32)).astype(np.int64)
            else:   # (oracle_idwt already undid the 2^8 pre-shift of the fixed-point 9/7 path)
                val = pl.view(np.int32).astype(np.int64)
            cb = int(tc["cbps"])
            val = np.clip(val + (1 << (cb - 1)), 0, (1 << cb) - 1) << int(g["shift"][c])
            val = val.reshape(H, W)
            rb = rowb[plane]
            ps, so = int(g["pixel_stride"]), int(g["sample_off"][c])
            for y in range(H):
                row = view[(int(tc["out_y"]) + y) * rb: (int(tc["out_y"]) + y + 1) * rb]
                r = row.view(np.uint16) if wide else row
                r[int(tc["out_x"]) * ps + so: (int(tc["out_x"]) + W) * ps + so: ps] = val[y]
    return info, pic


LOWRES_STREAMS = [
    dict(w=256, h=200, nc=3, bd=12, levels=5),
    dict(w=250, h=131, nc=3, bd=12, levels=5, tile=(128, 64)),
    dict(w=256, h=256, nc=3, bd=10, levels=3, reversible=False, base_step=1.0),
    dict(w=255, h=253, nc=3, bd=12, levels=3, cblk=(32, 32), npasses=3, vsc=True),
    dict(w=300, h=200, nc=1, bd=16, levels=5, cblk=(128, 32)),
    dict(w=256, h=256, nc=3, bd=12, levels=3, cblk=(32, 32), precincts=[(6, 6), (7, 7)], prog=3),
]


@pytest.mark.parametrize("cfg", LOWRES_STREAMS, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items())[:60])
def test_plan_with_lowres_reproduces_oracle_picture(cfg):
    """AVCodecContext.lowres: the dropped resolutions are parsed but flagged, planes / DWT levels / output
    shrink as the reference's do (jpeg2000dec.c:314-320, :509-520, :1047-1050)."""
    cfg = dict(cfg)
    w, h, nc, bd = cfg.pop("w"), cfg.pop("h"), cfg.pop("nc"), cfg.pop("bd")
    planes = j2kgen.synth_image(w, h, nc, bd, seed=12)
    p = j2kgen.EncParams(**cfg)
    cs = j2kgen.encode(planes, bd, p)
    for lowres in range(1, p.levels + 2):
        for bitexact in ((0,) if p.reversible else (0, 1)):
            flags = (fb.FLAG_BITEXACT if bitexact else 0) | fb.flag_lowres(lowres)
            r, oinfo, opic = oracle_decode(cs, bitexact | (lowres << 8))
            if r < 0:
                assert lowres > p.levels and r == -22
                with pytest.raises(fb.DecodeError) as e:
                    fb.plan(cs, flags)
                assert e.value.code == r
                continue
            info, pic = picture_from_plan(cs, flags)
            assert (info.width, info.height, info.pix_fmt) == (oinfo[0], oinfo[1], oinfo[2])
            assert np.array_equal(pic, opic), (lowres, bitexact)


@pytest.mark.parametrize("cfg", STREAMS, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items())[:60])
def test_plan_reproduces_oracle_picture(cfg):
    cfg = dict(cfg)
    w, h, nc, bd = cfg.pop("w"), cfg.pop("h"), cfg.pop("nc"), cfg.pop("bd")
    planes = j2kgen.synth_image(w, h, nc, bd, seed=11)
    p = j2kgen.EncParams(**cfg)
    cs = j2kgen.encode(planes, bd, p)
    for bitexact in ((0,) if p.reversible else (0, 1)):
        r, oinfo, opic = oracle_decode(cs, bitexact)
        assert r == 0
        info, pic = picture_from_plan(cs, fb.FLAG_BITEXACT if bitexact else 0)
        assert (info.width, info.height) == (oinfo[0], oinfo[1])
        assert info.pix_fmt == oinfo[2]          # same numbering on both sides
        assert np.array_equal(pic, opic)


def test_geometry_cache_across_frames():
    """Second frame with the same main header reuses the cached geometry; a different header rebuilds."""
    a = j2kgen.encode(j2kgen.synth_image(96, 64, 3, 12, seed=1), 12, j2kgen.EncParams(levels=3, cblk=(32, 32)))
    b = j2kgen.encode(j2kgen.synth_image(96, 64, 3, 12, seed=2), 12, j2kgen.EncParams(levels=3, cblk=(32, 32)))
    c = j2kgen.encode(j2kgen.synth_image(80, 64, 1, 8, seed=3), 8, j2kgen.EncParams(levels=2, cblk=(16, 16)))
    for cs in (a, b, c, a):
        r, _, opic = oracle_decode(cs)
        _, pic = picture_from_plan(cs, 0)
        assert r == 0 and np.array_equal(pic, opic)


def test_error_behaviour_matches_oracle():
    """Truncated / damaged codestreams: the host parser fails when the reference's front end fails."""
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
        r, _, opic = oracle_decode(bytes(bad))
        try:
            _, pic = picture_from_plan(bytes(bad), 0)
            ok = True
        except fb.DecodeError as e:
            ok = False
            assert e.code < 0
        assert ok == (r == 0)
        if ok:
            assert np.array_equal(pic, opic)


def test_library_exports_every_declared_symbol():
    L = fb.lib()
    header = (fb._DIR.parent / "include" / "b200j2k.h").read_text()
    import re
    declared = set(re.findall(r"\b(b200j2k_[a-z0-9_]+)\s*\(", header))
    declared -= {"b200j2k_decoder"}
    assert declared == set(fb.EXPORTS), declared ^ set(fb.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.b200j2k_version().startswith(b"b200j2k")


def test_frame_info_layout_matches_the_header(tmp_path):
    """The ctypes mirror of b200j2k_frame_info (what every Python caller reads) against the C header: size and the
    offsets of the fields behind the arrays, taken from a C compiler."""
    import ctypes
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "b200j2k.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu\\n", sizeof(b200j2k_frame_info), '
                   'offsetof(b200j2k_frame_info, plane_row_bytes), offsetof(b200j2k_frame_info, bytes_consumed), '
                   'offsetof(b200j2k_frame_info, av_pix_fmt_name), offsetof(b200j2k_frame_info, sar_num)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)])
    want = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    F = fb.FrameInfo
    assert want == [ctypes.sizeof(F), F.plane_row_bytes.offset, F.bytes_consumed.offset, F.av_pix_fmt_name.offset,
                    F.sar_num.offset]


def test_no_device_fails_loudly():
    """No CPU fallback: without a CUDA device the compute entry points return the device error."""
    L = fb.lib()
    if L.b200j2k_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(fb.DecodeError) as e:
        fb.Jpeg2000Decoder()
    assert e.value.code == fb.ERR_DEVICE
    with pytest.raises(fb.DecodeError):
        fb.idwt(1, [0, 8, 0, 8], 1, np.zeros(64, np.int32))


def test_bytes_consumed_matches_reference():
    """decode()'s return value: the reference returns bytestream2_tell() of the last tile-part sub-stream."""
    from helpers import have_ref, ref_decode
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    for cfg in STREAMS:
        cfg = dict(cfg)
        w, h, nc, bd = cfg.pop("w"), cfg.pop("h"), cfg.pop("nc"), cfg.pop("bd")
        cs = j2kgen.encode(j2kgen.synth_image(w, h, nc, bd, seed=11), bd, j2kgen.EncParams(**cfg))
        r, info, _, _ = ref_decode(cs, 0)
        assert r == 0
        assert fb.plan(cs, 0)["info"].bytes_consumed == info[4]


@pytest.mark.parametrize("cfg", LAYERED, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items())[:60])
def test_plan_of_two_layer_streams(cfg):
    """Two quality layers: pass counts with placeholder triples, Lcup / Lref accumulated over the packets,
    segments of one block gathered behind the codestream when its contributions are not adjacent."""
    cfg = dict(cfg)
    w, h, nc, bd = cfg.pop("w"), cfg.pop("h"), cfg.pop("nc"), cfg.pop("bd")
    planes = j2kgen.synth_image(w, h, nc, bd, seed=4)
    cs = j2kgen.encode(planes, bd, j2kgen.EncParams(layers=2, **cfg))
    P = fb.plan(cs, 0)
    npass = P["cblk"]["npasses"]
    assert (npass > 3).any()                                  # placeholder triples do occur
    if cfg.get("npasses", 1) > 1:
        assert len(P["extra"]) > 0                            # and so do HT sets split over two packets
    for bitexact in ((0,) if cfg.get("reversible", True) else (0, 1)):
        r, oinfo, opic = oracle_decode(cs, bitexact)
        assert r == 0
        info, pic = picture_from_plan(cs, fb.FLAG_BITEXACT if bitexact else 0)
        assert np.array_equal(pic, opic)


@pytest.mark.parametrize("t", POC_STREAMS, ids=lambda t: "-".join("".join(map(str, e)) for e in t["poc"]))
def test_plan_of_poc_streams(t):
    """POC marker: the host parser walks the entries like jpeg2000_decode_packets (jpeg2000dec.c:1835-1869)."""
    cs, _ = poc_stream(t)
    for bitexact in ((0,) if t["cfg"].get("reversible", True) else (0, 1)):
        r, oinfo, opic = oracle_decode(cs, bitexact)
        assert r == 0
        info, pic = picture_from_plan(cs, fb.FLAG_BITEXACT if bitexact else 0)
        assert np.array_equal(pic, opic)


def test_poc_marker_validation_matches_oracle():
    for bad in poc_marker_cases():
        r, _, opic = oracle_decode(bad)
        try:
            _, pic = picture_from_plan(bad, 0)
            assert r == 0 and np.array_equal(pic, opic)
        except fb.DecodeError as e:
            assert r < 0 and e.code < 0


@pytest.mark.parametrize("t", PACKED, ids=lambda t: f"{'ppt' if t['mode'] == 1 else 'ppm'}-{t['chunk']}")
def test_plan_of_packed_header_streams(t):
    """PPT / PPM: packet headers are read from the marker segments, bodies stay referenced in place."""
    cs, _ = packed_stream(t)
    for bitexact in ((0,) if t["cfg"].get("reversible", True) else (0, 1)):
        r, oinfo, opic = oracle_decode(cs, bitexact)
        assert r == 0
        info, pic = picture_from_plan(cs, fb.FLAG_BITEXACT if bitexact else 0)
        assert np.array_equal(pic, opic)


@pytest.mark.parametrize("t", TILE_PARTS, ids=lambda t: f"mode{t['mode']}-tp{t['tp']}-il{int(t['il'])}")
def test_plan_of_tile_part_streams(t):
    """Several (interleaved) tile-parts per tile, with PPT / PPM: same verdict and picture as the oracle, the
    select_header quirk of PPM streams included."""
    cs, _ = tile_part_stream(t)
    r, oinfo, opic = oracle_decode(cs, 0)
    try:
        info, pic = picture_from_plan(cs, 0)
        assert r == 0 and np.array_equal(pic, opic)
    except fb.DecodeError as e:
        assert r < 0 and e.code < 0


@pytest.mark.parametrize("t", TILE_STYLES, ids=lambda t: f"var{t['var']}-" + "-".join(f"{k}{v}" for k, v in t["cfg"].items())[:30])
def test_plan_of_tile_style_streams(t):
    """COD / QCD / POC in tile-part headers: geometry, step sizes and packet order per tile."""
    cs, _ = tile_style_stream(t)
    for bitexact in ((0,) if t["cfg"].get("reversible", True) else (0, 1)):
        r, oinfo, opic = oracle_decode(cs, bitexact)
        assert r == 0
        info, pic = picture_from_plan(cs, fb.FLAG_BITEXACT if bitexact else 0)
        assert np.array_equal(pic, opic)


def test_tile_style_marker_validation_matches_oracle():
    n_ok = 0
    for bad in tile_style_marker_cases():
        r, _, opic = oracle_decode(bad)
        try:
            _, pic = picture_from_plan(bad, 0)
            assert r == 0 and np.array_equal(pic, opic)
            n_ok += 1
        except fb.DecodeError as e:
            assert r < 0 and e.code == r
    assert n_ok >= 4


def test_geometry_cache_sees_tile_header_changes():
    """Two frames with the same main header but different tile-part COD segments must not share cached geometry."""
    import ctypes
    planes = j2kgen.synth_image(128, 96, 3, 8, seed=4)
    kw = dict(levels=2, cblk=(32, 32), tile=(64, 48))
    a = j2kgen.encode(planes, 8, j2kgen.EncParams(tile_var=4, **kw))       # odd tiles: narrower code-blocks
    b = j2kgen.encode(planes, 8, j2kgen.EncParams(tile_var=2, **kw))       # odd tiles: one level less
    for cs in (a, b, a, b):
        r, _, opic = oracle_decode(cs)
        _, pic = picture_from_plan(cs, 0)
        assert r == 0 and np.array_equal(pic, opic)


def test_plan_of_dcinema_profile_streams():
    n = 0
    for name, cs in dcinema_cases():
        r, _, opic = oracle_decode(cs)
        try:
            info, pic = picture_from_plan(cs, 0)
            assert r == 0 and info.pix_name == oracle_pix_name() and np.array_equal(pic, opic), name
            n += 1
        except fb.DecodeError as e:
            assert e.code == r, name
    assert n >= 18


def test_sample_aspect_ratio_of_res_boxes():
    for res in RES_BOXES:
        data = res_box_file(res)
        r, oinfo, _ = oracle_decode(data)
        i = fb.probe(data)
        assert r == 0 and (i.sar_num, i.sar_den) == (oinfo[7] >> 32, oinfo[7] & 0xFFFFFFFF), res


def test_plan_of_part1_styled_streams():
    """Part-1 blocks under context reset / causal contexts / segmentation symbols: descriptor flags carry the style."""
    n = 0
    for name, cs in part1_styled_streams():
        r, _, opic = oracle_decode(cs)
        try:
            _, pic = picture_from_plan(cs, 0)
            assert r == 0 and np.array_equal(pic, opic), name
            n += 1
        except fb.DecodeError as e:
            assert e.code == r, name
    assert n >= 20


@pytest.mark.parametrize("bd,csp,pal", PAL8_CASES)
def test_plan_of_pal8_files(bd, csp, pal):
    """PAL8: plane 0 = indices (gray8 layout), plane 1 = one row of 256 palette words."""
    for kw in (dict(), dict(tile=(64, 32))):
        data, _ = pal8_file(bd, csp, pal, **kw)
        r, oinfo, opic = oracle_decode(data)
        info, pic = picture_from_plan(data, 0)
        assert r == 0 and info.pix_name == oracle_pix_name() and np.array_equal(pic, opic)
        if info.pix_name == "pal8":
            assert info.nplanes == 2 and info.plane_row_bytes[1] == 1024 and info.plane_height[1] == 1


def test_psot_bound_matches_oracle():
    for i, c in enumerate(psot_bound_cases()):
        r, _, opic = oracle_decode(c)
        try:
            _, pic = picture_from_plan(c, 0)
            assert r == 0 and np.array_equal(pic, opic), i
        except fb.DecodeError as e:
            assert e.code == r, i


@pytest.mark.parametrize("t", HT_SETS, ids=HT_SET_IDS)
def test_plan_of_ht_set_streams(t):
    """HTDECLARED / MIXED streams: which block coder a block belongs to is decided per block in tier 2."""
    cs, _ = ht_set_stream(t)
    for bitexact in ((0,) if t["cfg"].get("reversible", True) else (0, 1)):
        r, oinfo, opic = oracle_decode(cs, bitexact)
        try:
            info, pic = picture_from_plan(cs, fb.FLAG_BITEXACT if bitexact else 0)
            assert r == 0 and np.array_equal(pic, opic)
        except fb.DecodeError as e:
            assert r < 0 and e.code == r


def test_plan_of_damaged_ht_set_streams():
    n_ok = 0
    for i, bad in enumerate(ht_set_damaged_streams()):
        r, _, opic = oracle_decode(bad)
        try:
            _, pic = picture_from_plan(bad, 0)
            assert r == 0 and np.array_equal(pic, opic), i
            n_ok += 1
        except fb.DecodeError as e:
            assert r < 0 and e.code < 0, i
    assert n_ok >= 20


def test_packed_header_marker_validation_matches_oracle():
    n_ok = 0
    for bad in packed_marker_cases():
        r, _, opic = oracle_decode(bad)
        try:
            _, pic = picture_from_plan(bad, 0)
            assert r == 0 and np.array_equal(pic, opic)
            n_ok += 1
        except fb.DecodeError as e:
            assert r < 0 and e.code < 0
    assert n_ok >= 2


@pytest.mark.parametrize("t", PART1, ids=lambda t: f"{t['mode']}-{t['w']}x{t['h']}")
def test_plan_of_part1_streams(t):
    """Part-1 (MQ) code-blocks: T2 lengths, nonzerobits, band orientation and placement in the flat plan."""
    for jp2 in (False, True):
        cs = part1_stream(t, jp2)
        for bitexact in ((0,) if not t["kw"].get("irreversible") else (0, 1)):
            r, oinfo, opic = oracle_decode(cs, bitexact)
            assert r == 0
            info, pic = picture_from_plan(cs, fb.FLAG_BITEXACT if bitexact else 0)
            assert (info.width, info.height) == (oinfo[0], oinfo[1])
            assert np.array_equal(pic, opic)


def test_tile_without_first_tile_part_fails_like_the_reference():
    """A tile that never gets a tile-part with TPsot = 0 has no coding style: the reference's init_tile returns
    AVERROR_INVALIDDATA for the frame (jpeg2000dec.c:862-869, :1054-1055)."""
    planes = j2kgen.synth_image(96, 64, 3, 8, seed=5)
    cs = j2kgen.encode(planes, 8, j2kgen.EncParams(levels=2, tile=(64, 64)))
    sot = [i for i in range(len(cs) - 3) if cs[i:i + 4] == b"\xff\x90\x00\x0a"]
    assert len(sot) == 2
    renumbered = bytearray(cs)
    renumbered[sot[1] + 10] = 1                                   # the second tile only has "tile-part 1"
    for bad in (cs[:sot[1]], cs[:sot[1] + 3], bytes(renumbered)):
        assert oracle_decode(bad)[0] == fb.ERR_INVALIDDATA
        with pytest.raises(fb.DecodeError) as e:
            fb.plan(bad, 0)
        assert e.value.code == fb.ERR_INVALIDDATA


@pytest.mark.parametrize("name,subs,bd", SUBSAMPLED, ids=[s[0] for s in SUBSAMPLED])
def test_plan_of_subsampled_and_jp2_streams(name, subs, bd):
    """Planar output through the plan: component -> plane by the channel definitions, plane sizes of the
    chosen AVPixelFormat; the same stream wrapped as JP2 with a cdef box that swaps the chroma planes."""
    for w, h, kw in SUBSAMPLED_GEOMS[:2]:
        cs, planes, p = subsampled_stream(subs, bd, w, h, kw)
        nc = len(subs)
        swapped = j2kgen.wrap_jp2(cs, w, h, nc, bd, colour_space=18,
                                  cdef=[(0, 0, 1), (1, 0, 3), (2, 0, 2), (3, 1, 0)][:nc])
        for data in (cs, swapped):
            r, oinfo, opic = oracle_decode(data, 0)
            assert r == 0
            info, pic = picture_from_plan(data, 0)
            assert (info.width, info.height) == (oinfo[0], oinfo[1])
            assert np.array_equal(pic, opic)
    # 4:4:4 with the component transform, delivered planar because the JP2 header says sYCC
    nc = min(len(subs), 4)
    full = j2kgen.synth_image(120, 70, nc, bd, seed=2)
    cs = j2kgen.encode(full, bd, j2kgen.EncParams(levels=3, mct=True))
    data = j2kgen.wrap_jp2(cs, 120, 70, nc, bd, colour_space=18)
    r, oinfo, opic = oracle_decode(data, 0)
    assert r == 0
    info, pic = picture_from_plan(data, 0)
    assert info.pix_fmt in (fb.lib() and 9, 10) and np.array_equal(pic, opic)


def test_jp2_wrapper_headers_match_reference():
    """JP2 / JPH files (jp2_find_codestream, jpeg2000dec.c:2658-2804): the codestream is found behind the
    boxes, the enumerated colour space picks the pixel-format list (sRGB / grey / sYCC), a palette steers
    the choice, the search gives up after ten foreign boxes -- dimensions, AVPixelFormat name and error code
    equal the reference's for every combination."""
    from helpers import have_ref, ref_decode
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    n = 0
    for nc, bd in [(3, 8), (3, 12), (1, 8), (1, 16), (2, 8), (4, 8), (3, 10), (4, 16), (2, 16), (1, 12)]:
        planes = j2kgen.synth_image(64, 48, nc, bd, seed=3)
        cs = j2kgen.encode(planes, bd, j2kgen.EncParams(levels=2, mct=nc >= 3))
        variants = [cs]
        for csp in (None, 16, 17, 18, 99):
            variants.append(j2kgen.wrap_jp2(cs, 64, 48, nc, bd, colour_space=csp))
        variants.append(j2kgen.wrap_jp2(cs, 64, 48, nc, bd, colour_space=16, palette=(16, 8)))
        variants.append(j2kgen.wrap_jp2(cs, 64, 48, nc, bd, colour_space=18, cdef=[(0, 0, 2), (1, 0, 1), (2, 0, 3)][:nc]))
        variants.append(j2kgen.wrap_jp2(cs, 64, 48, nc, bd, extra_before=11))
        variants.append(j2kgen.wrap_jp2(cs, 64, 48, nc, bd, extra_before=8))
        variants.append(j2kgen.wrap_jp2(cs, 64, 48, nc, bd)[:40])
        for data in variants:
            r, info, out, name = ref_decode(data)
            try:
                i = fb.probe(data)
                mine = (0, i.width, i.height, i.pix_name)
            except fb.DecodeError as e:
                mine = (e.code,)
            assert mine == ((r, info[0], info[1], name) if r == 0 else (r,)), (nc, bd, len(data))
            n += 1
    assert n == 110


def test_component_must_fit_its_output_plane():
    """A 'cdef' box names the output plane of every component of a planar picture (write_frame, jpeg2000dec.c:2325-2326).
    A crafted one sends the full-size luma component of a 4:2:0 stream to a half-size chroma plane: the reference then
    writes past the end of that AVFrame plane.  The plan refuses such a file (INVALIDDATA) instead of carrying the
    out-of-bounds write over to the device; honest channel definitions of the same stream still pass."""
    subs, bd = [(1, 1), (2, 2), (2, 2)], 8
    for w, h, kw in SUBSAMPLED_GEOMS[:2]:
        cs, planes, p = subsampled_stream(subs, bd, w, h, kw)
        good = j2kgen.wrap_jp2(cs, w, h, 3, bd, colour_space=18, cdef=[(0, 0, 1), (1, 0, 3), (2, 0, 2)])
        assert fb.plan(good, 0)["info"].width == w
        for cdef in ([(0, 0, 2), (1, 0, 1), (2, 0, 3)], [(0, 0, 3), (1, 0, 2), (2, 0, 1)]):
            bad = j2kgen.wrap_jp2(cs, w, h, 3, bd, colour_space=18, cdef=cdef)
            with pytest.raises(fb.DecodeError) as e:
                fb.plan(bad, 0)
            assert e.value.code == fb.ERR_INVALIDDATA


def test_plan_of_fuzz_findings_matches_reference():
    """tests/golden/frames_fuzz.npz: the damaged streams on which the differential fuzzer (tests/fuzz_host.py) caught the
    parser deviating from the reference -- verdict and picture of the REFERENCE, reproduced from the product's tables."""
    import hashlib
    z = np.load(ROOT / "tests" / "golden" / "frames_fuzz.npz")
    n = 0
    for key in sorted(k[:-3] for k in z.files if k.endswith("|cs")):
        cs, meta, out = z[key + "|cs"].tobytes(), z[key + "|meta"], z[key + "|out"]
        if meta[2] < 0:
            with pytest.raises(fb.DecodeError) as e:
                picture_from_plan(cs, 0)
            assert e.value.code == meta[2], key                  # the bare -1 of the reference, too
        else:
            info, pic = picture_from_plan(cs, 0)
            assert (info.width, info.height) == (meta[0], meta[1]), key
            assert hashlib.sha1(pic.tobytes()).digest() == out.tobytes(), key
            assert info.bytes_consumed == meta[3], key           # what FFCodec.cb.decode returns
        n += 1
    assert n == 28
