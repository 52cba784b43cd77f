#!/usr/bin/env python3
"""Times the reference's own gpu_fixed OMS kernel (oracle/_ref/libref_gpu_<code>.so: unmodified sources, sm_100a build) on AWGN
frames, in its own process — "the kernel to beat" line of bench.py.  Prints one JSON object.  Checker-side tool (uses tests/_helpers)."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from _helpers import Code, awgn_llr, ref_gpu, ref_gpu_decode   # noqa: E402

code_name, frames, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
L = ref_gpu(code_name)
if L is None:
    print(json.dumps({"unavailable": "oracle/_ref/libref_gpu_%s.so not built" % code_name})); sys.exit(0)
code = Code.load(code_name)
llr = awgn_llr(code, frames, 2.0, seed=2024)
ref_gpu_decode(L, "OMS", llr[:4096], iters, want_state=False)
best = min((ref_gpu_decode(L, "OMS", llr, iters, want_state=False) for _ in range(3)), key=lambda r: r["kernel_ms"])
k = code.k_info
print(json.dumps({"what": "LDPC_Sched_Stage_1_OMS_SIMD (gpu_fixed, unmodified, sm_100a build) on this GPU: decode kernel only / H2D..D2H with its interleavers",
                  "frames": frames, "kernel_ms": best["kernel_ms"], "total_ms": best["total_ms"],
                  "info_gbps_kernel": frames * k / best["kernel_ms"] / 1e6, "info_gbps_total": frames * k / best["total_ms"] / 1e6}))
