#!/usr/bin/env python3
"""Blocking host-buffer decode (H2D + interleave + kernel + de-interleave + D2H) of the frame-parallel kernels: the batch quartered over
the four stream slots (default) against one chunk (chunk_waves=1: a wave of these kernels is 300 Ki frames).  Diagnostic."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import ldpcgputegra_b200 as pkg

for name, F in (("64800x32400", 32768), ("64800x32400", 65536), ("4000x2000", 131072)):
    code = pkg.Code.load(name)
    src = pkg.PinnedArray((F, code.n), np.int8)
    dst = pkg.PinnedArray((F, code.n), np.uint8)
    tmp = pkg.CGPUDecoder(code, nb_frames=F)
    src.array[:] = tmp.awgn(F, pkg.sigma_for(2.0, code.k_info / code.n), 5)
    tmp.close()
    ref = None
    for kw in (dict(chunk_waves=1), dict()):
        dec = pkg.CGPUDecoder(code, nb_frames=F, **kw)
        for _ in range(2):
            dec.decode(src.array, 10, out=dst.array)
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            dec.decode(src.array, 10, out=dst.array)
        dt = (time.perf_counter() - t0) / reps
        same = True if ref is None else bool(np.array_equal(ref, dst.array))
        if ref is None: ref = dst.array.copy()
        print(json.dumps(dict(code=name, frames=F, kernel=dec.info(pkg.INFO_KERNEL), **{k: str(v) for k, v in kw.items()}, ms=dt * 1e3, mframes_s=F / dt / 1e6,
                              info_gbps=F * code.k_info / dt / 1e9, pcie_gbs_each_way=F * code.n / dt / 1e9, same_as_single_chunk=same)), flush=True)
        dec.close()
    del src, dst
