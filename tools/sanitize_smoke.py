#!/usr/bin/env python3
"""Small invocations of every kernel family, meant to be run under `compute-sanitizer --tool memcheck` (and racecheck)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import ldpcgputegra_b200 as pkg

rng = np.random.Generator(np.random.Philox(1))
def llr8(code, F): return rng.integers(-31, 32, size=(F, code.n), dtype=np.int8)

c576, c200, c1944 = pkg.Code.load("576x288"), pkg.Code.load("200x100"), pkg.Code.load("1944x972")
runs = [
    (c576, dict(kernel=2), 301), (c576, dict(kernel=2, early_term=1, out_format=1), 77), (c576, dict(kernel=2, no_static=1), 129),
    (c200, dict(kernel=2), 63), (c1944, dict(kernel=2), 40),
    (c576, dict(kernel=1), 130), (c576, dict(kernel=1, early_term=1), 70), (c576, dict(kernel=4), 700), (c1944, dict(kernel=4, out_format=1), 600),
    (c576, dict(kernel=3), 100), (c576, dict(kernel=5), 100), (c576, dict(schedule="FLOODING", early_term=1), 75), (c200, dict(schedule="FLOODING", kernel=3), 33),
]
for code, kw, F in runs:
    dec = pkg.CGPUDecoder(code, nb_frames=F, device=0, **kw)
    dec.set_debug(True)
    h, it = dec.decode(llr8(code, F), 4, want_iters=True)
    dec.debug_state(F)
    print("ok", code.n, kw, dec.info(pkg.INFO_KERNEL), int(h.sum()))
    dec.close()
for dt, kw in (("F32", dict(algo="NMS", schedule="FLOODING", early_term=1)), ("F32", dict(algo="OMS", kernel=3)), ("I16", dict(semantics="UNIFORM", sat_var=4095, sat_msg=1023))):
    dec = pkg.CGPUDecoder(c576, nb_frames=90, device=0, dtype=dt, **kw)
    y = dec.awgn(90, 0.8, seed=3)
    dec.decode(y, 5)
    print("ok", dt, kw, dec.info(pkg.INFO_KERNEL))
    dec.close()
enc = pkg.Encoder(c576); cw = enc.encode(rng.integers(0, 2, size=(45, 288), dtype=np.uint8)); enc.close()
enc = pkg.Encoder(pkg.Code.load("64800x32400")); cw = enc.encode(rng.integers(0, 2, size=(33, 32400), dtype=np.uint8)); enc.close()
dec = pkg.CGPUDecoder(c576, nb_frames=4096, device=0)
q = dec.awgn(1000, 0.8, seed=1); dec.decode(q, 3); dec.quantize(np.linspace(-5, 5, 1001, dtype=np.float32)); dec.close()
print("all done")
