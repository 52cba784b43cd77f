#!/usr/bin/env python3
"""Quick parity check of the staged kernel's tensor-map paths (message map, gather4) against the plain bulk-copy path."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1])); sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import numpy as np
import ldpcgputegra_b200 as pkg
from _helpers import awgn_llr, stress_llr, oracle_decode
for name, F in (("576x288", 1500), ("4000x2000", 700), ("1200x600", 600)):
    code = pkg.Code.load(name)
    llr = np.concatenate([awgn_llr(code, F, 2.0, 5), stress_llr(code, 100, 6)])
    ref = None
    for tma, g4 in ((1, 1), (2, 1), (1, 2), (2, 2)):
        dec = pkg.CGPUDecoder(code, nb_frames=llr.shape[0], kernel=4, fs_tma=tma, fs_g4=g4)
        dec.set_debug(True)
        hard = dec.decode(llr, 6)
        post, msgs = dec.debug_state(llr.shape[0])
        dec.close()
        if ref is None:
            o = oracle_decode(code, pkg.default_params(), llr, 6)
            ref = (o["hard"], o["post"], o["msgs"])
        ok = all(np.array_equal(a, b) for a, b in zip((hard, post, msgs), ref))
        print(name, "tma", tma, "g4", g4, "bit-exact" if ok else "MISMATCH", flush=True)
