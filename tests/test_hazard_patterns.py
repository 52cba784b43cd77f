"""CPU-side check of the GPU parity suite's reach: the staged frame-parallel kernel (kernel_fs.cuh) treats a row according to its
hazard pattern — no edge written within the last 16 rows, exactly one such edge whose writer is 1..4 rows back (copied from the
forwarding ring), one that is farther back (re-read from global memory), several (the per-edge path).  The codes the GPU tests run
through that kernel must exhibit every pattern, and DVB-S2 must be the pure staircase the fast path was written for.  This
restates the host-side classification of ldpc_b200_create (csrc/ldpc_b200.cu: the pos2 loop) in numpy."""
import collections
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ldpcgputegra_b200 as pkg

FS_HAZARD, FS_FWD = 16, 4      # kernel_fs.cuh


def hazard_patterns(code):
    pos = np.asarray(code.pos)
    last = np.full(code.n, -(1 << 30), dtype=np.int64)
    stats = collections.Counter()
    n_checks = int(sum(code.rows))
    for lap in range(2):                      # lap 0 warms `last` up: the window is cyclic over the iteration boundary
        q, e = lap * n_checks, 0
        for D, R in zip(code.deg, code.rows):
            for _ in range(R):
                kinds = []
                for j in range(D):
                    back = q - last[pos[e + j]]
                    if back <= FS_HAZARD:
                        kinds.append("fwd%d" % back if back <= FS_FWD else "far")
                    last[pos[e + j]] = q
                if lap == 1:
                    stats["none" if not kinds else kinds[0] if len(kinds) == 1 else "multi"] += 1
                q += 1; e += D
    return stats


def test_gpu_tested_codes_cover_every_hazard_pattern():
    seen = collections.Counter()
    for name in ("576x288", "4000x2000", "1200x600", "816x408"):      # tests/test_parity_gpu.py: test_staged_kernel_*, test_other_codes
        seen.update(hazard_patterns(pkg.Code.load(name)))
    for kind in ("none", "fwd1", "fwd2", "fwd3", "fwd4", "far", "multi"):
        assert seen[kind] > 0, f"no row with hazard pattern {kind!r} in the codes the staged kernel is tested on"


def test_dvbs2_is_a_pure_staircase():
    c = pkg.Code.load("64800x32400")
    s = hazard_patterns(c)
    assert s == {"fwd1": c.n_checks - 1, "none": 1}
