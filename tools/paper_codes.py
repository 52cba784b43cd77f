#!/usr/bin/env python3
"""The reference paper's benchmark codes (paper/ldpcGpuTegra.tex:296-356) on this GPU: decode kernel on
device-resident AWGN frames at 10 and 5 iterations, int8 layered OMS (gpu_fixed semantics, as in the paper), air throughput in the
paper's own unit (coded Mbps = N * frames / s) next to the best published figure."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import ldpcgputegra_b200 as pkg

# best published air Mbps per code (10 it / 5 it) and where
PAPER = {"576x288": (127, 217, "GTX680, 3 streams"), "1024x518": (20, None, "Tegra K1, 3 streams"), "1200x600": (20, None, "Tegra K1, 3 streams"),
         "1944x972": (16, None, "Tegra K1, 3 streams"), "2304x1152": (132, 226, "GTX680, 3 streams"), "4000x2000": (131, 230, "GTX680, 3 streams"),
         "8000x4000": (32.76, None, "Tegra K1, 3 streams"), "9972x4986": (26, None, "Tegra K1, 3 streams")}

def run(name, frames, iters, reps=5):
    code = pkg.Code.load(name)
    dec = pkg.CGPUDecoder(code, nb_frames=frames, algo="OMS", semantics="GPU_FIXED")
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
    d_llr = torch.empty((frames, code.n), dtype=torch.int8, device="cuda"); d_hard = torch.empty((frames, code.n), dtype=torch.uint8, device="cuda")
    dec.awgn_device(d_llr.data_ptr(), frames, pkg.sigma_for(2.0, code.k_info / code.n), 1, 0, ts.cuda_stream)
    for _ in range(3): dec.decode_device(d_llr.data_ptr(), d_hard.data_ptr(), frames, iters, stream=ts.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): dec.decode_device(d_llr.data_ptr(), d_hard.data_ptr(), frames, iters, stream=ts.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    p10, p5, where = PAPER[name]
    pub = p10 if iters == 10 else p5
    out = dict(code=name, frames=frames, iters=iters, ms=ms, mframes_s=frames / ms / 1e3, air_mbps=frames * code.n / ms / 1e3, info_gbps=frames * code.k_info / ms / 1e6,
               kernel=dec.info(pkg.INFO_KERNEL), levels=dec.info(pkg.INFO_LEVELS), paper_best_air_mbps=pub, paper_hw=where,
               speedup_vs_paper=(frames * code.n / ms / 1e3 / pub) if pub else None)
    print(json.dumps(out), flush=True); dec.close()

for name, frames in [("576x288", 262144), ("1024x518", 131072), ("1200x600", 131072), ("1944x972", 65536), ("2304x1152", 65536), ("4000x2000", 131072), ("8000x4000", 131072), ("9972x4986", 131072)]:
    for iters in (10, 5):
        run(name, frames, iters)
