"""The fuzz tests once more against a DEBUG build of the HT kernels: in-kernel bounds assertions compiled in
(-DB200J2K_DEBUG_BOUNDS, ffmpeg-ht_b200/csrc/ht_common.cuh) and guard zones around every device buffer
(B200J2K_REDZONE=1, decoder.cu).  compute-sanitizer is not available on the GPU pool; with this, a kernel that reads
its un-stuffed arrays out of range or writes next to a buffer on damaged input fails the decode instead of passing
unnoticed.  Runs in a child process because the library and the switch are chosen when the library is loaded."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
DBG = ROOT / "ffmpeg-ht_b200" / "variants" / "lib_dbg.so"
pytestmark = pytest.mark.gpu


def debug_library():
    if not DBG.exists():
        r = subprocess.run([str(ROOT / "tools" / "build_variant.sh"), "dbg", "ht_cleanup.cu", "-DB200J2K_DEBUG_BOUNDS"],
                           capture_output=True, text=True)
        if r.returncode != 0 or not DBG.exists():
            pytest.skip("debug variant of the library could not be built here: " + r.stderr[-200:])
    return DBG


def test_fuzzed_inputs_against_the_debug_build():
    env = dict(os.environ, B200J2K_LIB=str(debug_library()), B200J2K_REDZONE="1")
    sel = ("damaged or fuzz or absurd or passes_but_no_bytes or failed_block_count or blocks_vs_oracle or "
           "htset_damaged or layered")
    r = subprocess.run([sys.executable, "-m", "pytest", str(ROOT / "tests" / "test_gpu_parity.py"), "-m", "gpu", "-q", "-x",
                        "-k", sel, "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, cwd=str(ROOT))
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "guard bytes" not in r.stderr and "assertions failed" not in r.stderr, tail


def test_guard_zones_catch_a_stray_write():
    """The checker itself: a child process decodes one frame with guard zones on and then lets the library's own
    self-test poke one byte behind the picture buffer (b200j2k_debug_poke): the next wait must fail."""
    code = (
        "import sys; sys.path[:0] = [%r, %r]\n"
        "import numpy as np, ffmpeg_ht_b200 as fb, j2kgen\n"
        "cs = j2kgen.encode(j2kgen.synth_image(128, 96, 3, 12, seed=1), 12, j2kgen.EncParams(levels=2))\n"
        "dec = fb.Jpeg2000Decoder(device=0, slots=1)\n"
        "info, pic = dec.decode(cs)\n"
        "assert fb.lib().b200j2k_debug_poke(dec._h, 0) == 0\n"
        "try:\n"
        "    dec.decode(cs)\n"
        "    print('NOT CAUGHT')\n"
        "except fb.DecodeError as e:\n"
        "    print('CAUGHT', e.code)\n" % (str(ROOT), str(ROOT / "tools")))
    env = dict(os.environ, B200J2K_REDZONE="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=str(ROOT))
    assert "CAUGHT -19" in r.stdout, (r.stdout + r.stderr)[-800:]
    assert "guard bytes of the picture buffer" in r.stderr
