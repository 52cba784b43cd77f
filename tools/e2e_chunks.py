#!/usr/bin/env python3
"""Blocking host-buffer decode of the flagship workload for several pipeline chunk sizes (LDPC_B200_CHUNK_FRAMES, an experiment knob of
the library; default = one kernel wave = 6808 frames).  One JSON line per point.  Diagnostic."""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
body = """
import sys, time, json, os
sys.path.insert(0, %r)
import numpy as np, ldpcgputegra_b200 as pkg
code = pkg.Code.load("576x288"); F = 65536
dec = pkg.CGPUDecoder(code, nb_frames=F, out_format=int(os.environ.get("PACKED", "0")))
src = [pkg.PinnedArray((F, code.n), np.int8) for _ in range(2)]
for s in src: s.array[:] = dec.awgn(F, pkg.sigma_for(2.0, 0.5), 5)
dst = pkg.PinnedArray((F, dec.hard_row_bytes), np.uint8)
for i in range(5): dec.decode(src[i %% 2].array, 10, out=dst.array)
best = 1e9
for rep in range(5):
    t0 = time.perf_counter()
    for i in range(20): dec.decode(src[i %% 2].array, 10, out=dst.array)
    best = min(best, (time.perf_counter() - t0) / 20)
print(json.dumps(dict(chunk_frames=os.environ.get("LDPC_B200_CHUNK_FRAMES", "default"), packed=int(os.environ.get("PACKED", "0")), ms=best * 1e3, info_gbps=F * code.k_info / best / 1e9)))
""" % str(ROOT)
for packed in ("0", "1"):
    for cf in ("", "1702", "3404", "4608", "6808", "13616", "32768"):
        env = dict(os.environ, PACKED=packed)
        if cf: env["LDPC_B200_CHUNK_FRAMES"] = cf
        r = subprocess.run([sys.executable, "-c", body], capture_output=True, text=True, env=env)
        print(r.stdout.strip() or r.stderr[-300:], flush=True)
