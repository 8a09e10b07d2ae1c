"""Damaged input must never make the host side read out of bounds or hand the device a descriptor that leaves
the packet: tools/fuzz_host.cc under AddressSanitizer + UBSan on mutated codestreams (plan builder + frame
splitter).  CPU only; skipped where the sanitizer runtime is missing."""
import os
import subprocess

import pytest

import j2kgen
from helpers import ROOT
from test_oracle_vs_ref import (HT_SETS, PACKED, PAL8_CASES, PART1, POC_STREAMS, TILE_PARTS, TILE_STYLES, ht_set_stream,
                                packed_stream, pal8_file, part1_stream, poc_stream, tile_part_stream, tile_style_stream)


def test_host_parser_and_splitter_under_sanitizers(tmp_path):
    exe = tmp_path / "fuzz_host"
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined,float-cast-overflow",
           "-fno-sanitize-recover=undefined", "-w",
           f"-I{ROOT / 'ffmpeg-ht_b200' / 'csrc'}", f"-I{ROOT / 'include'}", "-I/usr/local/cuda/include",
           str(ROOT / "tools" / "fuzz_host.cc"), str(ROOT / "ffmpeg-ht_b200" / "csrc" / "j2k_parse.cc"),
           str(ROOT / "ffmpeg-ht_b200" / "csrc" / "j2k_splitter.cc"), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build not available here: " + r.stderr[-200:])
    planes = j2kgen.synth_image(96, 64, 3, 8, seed=5)
    seeds = [j2kgen.encode(planes, 8, j2kgen.EncParams(levels=2, tile=(64, 64), sop=True, eph=True)),
             j2kgen.encode(planes, 8, j2kgen.EncParams(levels=3, cblk=(16, 16), precincts=[(4, 4), (5, 5)], prog=2,
                                                       npasses=3, layers=2))]
    seeds.append(j2kgen.wrap_jp2(seeds[0], 96, 64, 3, 8, colour_space=16, cdef=[(0, 0, 1), (1, 0, 2), (2, 0, 3)]))
    seeds += [packed_stream(t, w=90, h=70)[0] for t in (PACKED[0], PACKED[3], PACKED[5])]
    seeds.append(poc_stream(POC_STREAMS[2], w=90, h=70)[0])
    seeds += [part1_stream(PART1[0]), part1_stream(PART1[3], jp2=True), part1_stream(PART1[9])]
    seeds += [tile_part_stream(t, w=90, h=70)[0] for t in (TILE_PARTS[1], TILE_PARTS[7])]
    seeds += [tile_style_stream(t, w=90, h=70)[0] for t in (TILE_STYLES[5], TILE_STYLES[8])]
    seeds += [ht_set_stream(t, w=90, h=70)[0] for t in (HT_SETS[7], HT_SETS[10], HT_SETS[12], HT_SETS[18])]
    seeds.append(pal8_file(*PAL8_CASES[1], tile=(64, 32))[0])
    files = []
    for i, s in enumerate(seeds):
        f = tmp_path / f"seed{i}.j2k"
        f.write_bytes(s)
        files.append(str(f))
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([str(exe), "7", "2500"] + files, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "runtime error" not in r.stderr and "ERROR" not in r.stderr, r.stderr[-2000:]
    ok = int(r.stdout.split()[1])
    assert ok > 50, r.stdout            # the mutations must not all be rejected outright
