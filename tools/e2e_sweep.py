#!/usr/bin/env python3
"""Time the blocking host-buffer decode (H2D + kernel + D2H) for several pipeline chunk sizes / output formats.  Diagnostic."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import ldpcgputegra_b200 as pkg

code = pkg.Code.load("576x288")
F = 65536
src = pkg.PinnedArray((F, code.n), np.int8)
tmp = pkg.CGPUDecoder(code, nb_frames=F)
src.array[:] = tmp.awgn(F, pkg.sigma_for(2.0, 0.5), 5)
tmp.close()
for packed in (0, 1):
    dst = pkg.PinnedArray((F, code.n if not packed else code.n // 8), np.uint8)
    for waves in (1, 2, 3, 5):
        dec = pkg.CGPUDecoder(code, nb_frames=F, chunk_waves=waves, out_format=packed)
        for _ in range(3):
            dec.decode(src.array, 10, out=dst.array)
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            dec.decode(src.array, 10, out=dst.array)
        dt = (time.perf_counter() - t0) / reps
        print(json.dumps(dict(packed=packed, chunk_waves=waves, ms=dt * 1e3, mframes_s=F / dt / 1e6, info_gbps=F * code.k_info / dt / 1e9)))
        dec.close()
