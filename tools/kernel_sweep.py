#!/usr/bin/env python3
"""Time the on-chip kernel on device-resident frames for a few (G, P) groupings / codes.  Diagnostic, not the bench."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import ldpcgputegra_b200 as pkg

def run(code_name, frames, iters, reps=12, rotate=6, **kw):
    code = pkg.Code.load(code_name)
    dec = pkg.CGPUDecoder(code, nb_frames=frames, **kw)
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
    rotate = max(1, min(rotate, int(2e9 // (frames * code.n))))
    tdt = {0: torch.int8, 1: torch.int16, 2: torch.float32}[dec.params.dtype]
    rotate = max(1, min(rotate, int(2e9 // (frames * code.n * torch.empty((), dtype=tdt).element_size()))))
    d_llrs = [torch.empty((frames, code.n), dtype=tdt, device="cuda") for _ in range(rotate)]
    d_hard = torch.empty((frames, code.n), dtype=torch.uint8, device="cuda")
    for b, d in enumerate(d_llrs):
        dec.awgn_device(d.data_ptr(), frames, pkg.sigma_for(2.0, code.k_info / code.n), 1, b * frames, ts.cuda_stream)
    for i in range(3):
        dec.decode_device(d_llrs[i % rotate].data_ptr(), d_hard.data_ptr(), frames, iters, stream=ts.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        dec.decode_device(d_llrs[i % rotate].data_ptr(), d_hard.data_ptr(), frames, iters, stream=ts.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fe = int(d_hard[:, :code.k_info].any(dim=1).sum())
    out = dict(code=code_name, frames=frames, iters=iters, ms=ms, mframes_s=frames / ms / 1e3, info_gbps=frames * code.k_info / ms / 1e6,
               kernel_used=dec.info(pkg.INFO_KERNEL), frames_per_cta=dec.info(pkg.INFO_FRAMES_PER_CTA), smem=dec.info(pkg.INFO_SMEM_BYTES), fer=fe / frames, rotate=rotate, **{k: str(v) for k, v in kw.items()})
    print(json.dumps(out), flush=True); dec.close()
    del d_llrs, d_hard
    torch.cuda.empty_cache()

if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")        # e.g. "576x288:3:4" -> one run, for ncu
    ap.add_argument("--frames", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--runs", default="")        # JSON list of dicts: {"code":..., "frames":..., "iters":..., + CGPUDecoder keywords}
    a = ap.parse_args()
    if a.runs:
        for r in json.loads(a.runs):
            r = dict(r)
            if "group" in r: r["group"] = tuple(r["group"])
            run(r.pop("code"), r.pop("frames", a.frames), r.pop("iters", 10), reps=r.pop("reps", a.reps), rotate=r.pop("rotate", 6), **r)
        sys.exit(0)
    if a.only:
        name, G, P = a.only.split(":")
        run(name, a.frames, 10, reps=a.reps, rotate=1, group=(int(G), int(P)))
        sys.exit(0)
    for grp in [(1, 1), (3, 4)]:
        run("576x288", 65536, 10, group=grp, rotate=1)
        run("576x288", 65536, 10, group=grp, rotate=6)
    for name in ["1248x624", "1944x972", "2304x1152", "4000x2000", "1200x600"]:
        run(name, 16384, 10)
    run("576x288", 65536, 10, kernel=1)
    run("64800x32400", 2048, 10)
