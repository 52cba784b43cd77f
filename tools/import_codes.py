#!/usr/bin/env python3
"""Convert the reference's compile-time code tables into this library's binary .ldpc tables.

Run HERE (the container that has /root/reference); the GPU box only sees the generated files.  The parser is the product's
own H-matrix loader (ldpc_b200_load_code_header) so the conversion also exercises it.  x86-tree and gpu-tree tables of the
same code are checked to be identical.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from ldpcgputegra_b200 import Code, CODES_DIR

REF = Path("/root/reference/code")
X86 = {"576x288": "576x288", "1944x972": "1944x972", "2048x384": "2048x384", "2304x1152": "2304x1152", "4000x2000": "4000x2000",
       "64800x32400": "64800x32400.dvb-s2", "64800x7200": "64800x7200.dvb-s2", "64800x6480": "64800x6480.dvb-s2"}


# tables that only the ARM tree carries (code/ldpc_decoder_arm/Constantes; SURVEY App. C last row)
ARM_ONLY = {"155x93": "155x93", "2640x1320": "2640x1320", "1920x960": "802.11e.1920x960"}


def load_gpu(d: Path):
    """gpu_fixed flavour: macros in constantes_gpu.h + table in constantes_decoder.h; a few dirs only carry an x86-style header."""
    if (d / "constantes_decoder.h").exists():
        return Code.from_header(d / "constantes_gpu.h", d / "constantes_decoder.h")
    if (d / "constantes_sse.h").exists():
        return Code.from_header(d / "constantes_sse.h")
    raise FileNotFoundError(d)


def main():
    CODES_DIR.mkdir(exist_ok=True)
    done = {}
    for name, d in X86.items():
        done[name] = Code.from_header(REF / "x86/Constantes" / d / "constantes_sse.h")
    for d in sorted(p for p in (REF / "gpu_fixed/matrix").iterdir() if p.is_dir()):
        name = d.name
        try:
            g = load_gpu(d)
        except Exception as e:  # incomplete table in the reference tree
            print(f"{name}: skipped ({e})")
            continue
        if name in done:
            x = done[name]
            same = (x.n, x.n_checks, x.deg, x.rows) == (g.n, g.n_checks, g.deg, g.rows) and np.array_equal(x.pos, g.pos)
            print(f"{name}: x86 and gpu tables identical: {same}")
            if not same:
                g.save(CODES_DIR / f"{name}.gpu.ldpc")
        else:
            done[name] = g
    for name, d in ARM_ONLY.items():
        done[name] = Code.from_header(REF / "ldpc_decoder_arm/Constantes" / d / "constantes_sse.h")
    for name, c in done.items():
        c.save(CODES_DIR / f"{name}.ldpc")
        lv, _ = c.level_schedule()
        print(f"{name}: n={c.n} checks={c.n_checks} m={c.m} deg={c.deg} rows={c.rows} levels={lv}")


if __name__ == "__main__":
    main()
