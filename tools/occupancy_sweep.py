#!/usr/bin/env python3
"""What are resident warps worth to the flagship on-chip kernel?  Runs 576x288 / 64 Ki frames / 10 iterations with the number of
frame pairs per SM capped below what shared memory allows (LDPC_B200_RP_MAX_SLOTS, an experiment knob of the library): 4 pairs = 3 warps.
The slope of this curve at the top is the most a smaller (compressed) state could buy; DESIGN.md 3.1 sets it against what the
re-expansion of compressed messages would cost in issue slots.  One JSON line per point."""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
runs = json.dumps([{"code": "576x288", "frames": 65536, "iters": 10, "reps": 20, "rotate": 6}])
for slots in (4, 8, 12, 16, 20, 23, 24):
    env = dict(os.environ, LDPC_B200_RP_MAX_SLOTS=str(slots))
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "kernel_sweep.py"), "--runs", runs], capture_output=True, text=True, env=env)
    for line in r.stdout.strip().splitlines():
        d = json.loads(line)
        print(json.dumps(dict(max_pairs_per_sm=slots, pairs_per_sm=d["frames_per_cta"] // 2, warps_per_sm=3 * ((d["frames_per_cta"] // 2 + 3) // 4), ms=d["ms"], mframes_s=d["mframes_s"])), flush=True)
