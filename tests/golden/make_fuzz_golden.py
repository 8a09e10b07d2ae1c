#!/usr/bin/env python3
"""Generates tests/golden/frames_fuzz.npz with the REFERENCE (oracle/_ref/libj2kref.so built from /root/reference):
the damaged streams on which tests/fuzz_host.py (reference vs oracle vs product, differentially) found the oracle or
the product's parser deviating from the reference in round 2 -- each a generated stream plus a few damaged bytes --
with the reference's verdict and a SHA-1 digest of its picture.  Same layout as frames_tiles.npz (key = name|flags).
What each one pinned down:
  segs     more than six / eight codeword segments in one contribution of a block (lengthinc[] is as long as the
           pass count, jpeg2000dec.c:1227)
  qcd_len  a QCD whose style reads fewer bytes than the segment declares: the bare -1 of the marker loop (:2629-2633)
  qsty     an undefined quantisation style: parsed like style 2, step size 0 (jpeg2000.c:246-249)
  dwt      a wavelet byte that is neither 0 nor 1: accepted by get_cox, refused with -1 when the tile is set up
           (jpeg2000dwt.c:590)
  ppt_sod  packed packet headers of a tile whose first tile-part never reaches its SOD: an empty header stream
           (:2474-2476), every packet of the tile is empty
  tp_gap   a gap in the TPsot numbering: select_stream walks over empty slots (:1117-1126)
  lost     36 bytes lost in a tile-part header plus two damaged bytes
  layers   a layer count of 0xFF01: the reference keeps SGcod's 16-bit field in a uint8_t (jpeg2000.h:152) -- one layer
  jp2size  a JP2 box of absurd size: jp2_find_codestream returns an error code, which its caller takes for "found"
           (:2846) -- the codestream is then found by the scan for SOC
  bypass0  a Part-1 style with arithmetic-coding bypass announced, but no block ever contributes (no SOD): the style is
           only refused (PATCHWELCOME, not served) when a block of it is met in a packet
  plt      a PPT marker turned into a PLT whose last byte has the continuation bit set: get_plt refuses it (:951)
  levels32 32 decomposition levels (of a 1 x 1 picture): the reference's limit is JPEG2000_MAX_RESLEVELS = 34, not 33
  nosod    the only tile-part never reaches its SOD: the picture is the DC level and the decoder reports 0 bytes consumed
           (the position in the tile-part stream selected last: an empty slot, :1111-1134, :2903)
  tp1cod   a tile whose only tile-part says TPsot = 1 but carries its own COD / QCD: the styles are initialised, it decodes
  eoc2     an unknown progression order (no packet is read) and a run of bytes behind EOC: every tile still ends with
           the 2-byte EOC skip on the stream that is current -- here the whole packet: 184 bytes consumed, not 182
  badlen   no packet is read (unknown progression order) and the marker loop gives up on a length field behind the
           tile-part: the two bytes of that field count as consumed (:2485-2494)
  noqcd    no QCD at all (exponents and guard bits zero: M_b = -1) and no SOD: nothing decodes, the picture is the
           DC level -- not an error
Run where the reference is mounted:  python tests/golden/make_fuzz_golden.py"""
import hashlib
import importlib.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "tools"), str(ROOT / "tests")]
from helpers import ref_decode  # noqa: E402

spec = importlib.util.spec_from_file_location("fuzz_host", ROOT / "tests" / "fuzz_host.py")
fuzz_host = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fuzz_host)

# name: (family of tests/fuzz_host.py, index in it, operations in order: ("set", position, byte) / ("del", position, count))
CASES = {
    "segs": ("ht_sets", 2, [("set", 42009, 83)]),
    "qcd_len": ("plain", 19, [("set", 73, 146), ("set", 93, 244)]),
    "qsty": ("plain", 21, [("set", 73, 255), ("set", 107, 154), ("set", 116, 22)]),
    "dwt": ("plain", 19, [("set", 68, 9), ("set", 127, 64)]),
    "ppt_sod": ("packed", 4, [("set", 49370, 138)]),
    "tp_gap": ("tile_styles", 7, [("set", 5050, 3), ("set", 6200, 255), ("set", 15763, 238)]),
    "lost": ("tile_parts", 3, [("del", 472, 36), ("set", 3991, 68), ("set", 4614, 34)]),
    "noqcd": ("plain", 18, [("set", 69, 53), ("set", 88, 131)]),
    "layers": ("part1", 13, [("set", 136, 255), ("set", 595, 11)]),
    "jp2size": ("part1", 13, [("set", 32, 255)]),
    "nosod": ("plain", 18, [("set", 88, 254)]),
    "badlen": ("part1", 23, [("dup", 782, 145), ("set", 160, 248)]),
    "tp1cod": ("tile_styles", 1, [("set", 6809, 1)]),
    "eoc2": ("plain", 21, [("dup", 113, 69), ("set", 60, 197), ("set", 105, 80)]),
    "levels32": ("plain", 18, [("set", 64, 0x20)]),
    "plt": ("packed", 4, [("set", 37122, 0x58)]),
    "bypass0": ("plain", 18, [("set", 67, 19), ("set", 87, 247), ("set", 95, 255)]),
}


def marker(code, payload):
    return bytes([0xFF, code]) + (len(payload) + 2).to_bytes(2, "big") + bytes(payload)


# TLM / PLT / CRG segments spliced in front of the first SOT of a three-component stream: the reference only walks over
# them, but checks what it walks over (get_tlm :901-934 with its uint8_t entry count, get_plt :936-955, read_crg :875-883)
SPLICED = {
    "tlm_ok": marker(0x55, [0, 0x40] + [0] * 4 * 3),             # Stlm: ST 0, SP 1 -> 4-byte entries
    "tlm_st3": marker(0x55, [0, 0x30] + [0] * 6),                # ST = 3
    "tlm_odd": marker(0x55, [0, 0x10] + [0] * 7),                # 3-byte entries, 7 bytes: consumption != length -> -1
    "tlm_300": marker(0x55, [0, 0x00] + [0] * 2 * 300),          # 300 entries: the count wraps to 44 -> -1
    "plt_ok": marker(0x58, [0, 0x81, 0x02, 0x7F]),
    "plt_open": marker(0x58, [0, 0x05, 0x81]),                   # last length not terminated
    "plt_short": marker(0x58, [0]),                              # n < 4
    "crg_ok": marker(0x63, [0] * 12),
    "crg_bad": marker(0x63, [0] * 8),
    "cpf": marker(0x59, [0, 1]),
    "plm": marker(0x57, [0, 2, 1, 1]),
}
for _name, _seg in SPLICED.items():
    CASES[_name] = ("plain", 20, [("ins_before_sot", 0, _seg)])


def main():
    families = {}
    pack = {}
    for name, (fam, idx, ops) in CASES.items():
        if fam not in families:
            families[fam] = fuzz_host.family(fam)
        cs = bytearray(families[fam][idx])
        for op, pos, v in ops:
            if op == "set":
                cs[pos] = v
            elif op == "dup":
                cs[pos:pos] = cs[pos: pos + v]
            elif op == "ins_before_sot":
                at = bytes(cs).index(b"\xff\x90")
                cs[at:at] = v
            else:
                del cs[pos: pos + v]
        cs = bytes(cs)
        r, info, out, pix = ref_decode(cs, 0)
        k = f"{name}|0"
        pack[k + "|cs"] = np.frombuffer(cs, np.uint8)
        pack[k + "|meta"] = np.array([info[0], info[1], r, info[4]], np.int64)     # [3]: bytes consumed (decode's return)
        pack[k + "|pix"] = np.frombuffer((pix or "none").encode(), np.uint8)
        pack[k + "|out"] = np.frombuffer(hashlib.sha1(out.tobytes()).digest(), np.uint8) if r == 0 else np.zeros(0, np.uint8)
        print(f"{name}: {len(cs)} bytes, reference {r}" + (f", {info[0]}x{info[1]} {pix}" if r == 0 else ""))
    np.savez_compressed(ROOT / "tests" / "golden" / "frames_fuzz.npz", **pack)


if __name__ == "__main__":
    main()
