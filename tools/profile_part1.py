#!/usr/bin/env python3
"""ncu / timing driver for the Part-1 (MQ) path: stages one OpenJPEG-encoded 4K RGB 8-bit lossless frame and
re-runs its kernels.   python tools/profile_part1.py [runs]"""
import io
import sys
from pathlib import Path

import numpy as np
from PIL import Image

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import ffmpeg_ht_b200 as fb  # noqa: E402
import j2kgen  # noqa: E402

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
planes = j2kgen.synth_image(3840, 2160, 3, 8, seed=11)
b = io.BytesIO()
Image.fromarray(np.stack(planes, -1).astype(np.uint8), "RGB").save(b, "JPEG2000", no_jp2=True, irreversible=False,
                                                                 codeblock_size=(64, 64))
cs = b.getvalue()
with fb.Jpeg2000Decoder(device=0, slots=1) as dec:
    info = dec.stage(0, cs)
    dec.run_staged_timed(0)
    ms = np.mean([dec.run_staged_timed(0) for _ in range(runs)], axis=0)
    print("part1 4K", info.ncodeblocks, "blocks", len(cs), "bytes; stage ms [-, -, mq blocks, -, idwt levels, final]",
          [round(float(x), 3) for x in ms], flush=True)
