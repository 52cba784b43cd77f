#!/usr/bin/env python3
"""End-to-end (host buffers) throughput of the flagship workload with every rank of a torchrun job decoding at once, for the knobs
that could matter on a multi-GPU host: ordinary vs write-combined pinned input buffers, ranks pinned to disjoint host cores or not,
byte-per-bit vs bit-packed output — next to the raw pinned-copy rate of the box with all ranks copying (bench.host_link_probe).
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/e2e_multi.py
One JSON object per line on rank 0.  Diagnostic, not the bench."""
import json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
import bench as B
import ldpcgputegra_b200 as pkg

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
ranks = B.Ranks(torch, local, world)
code = pkg.Code.load("576x288")
F, n, k = 65536, code.n, code.k_info
all_cores = sorted(os.sched_getaffinity(0))
tmp = pkg.CGPUDecoder(code, nb_frames=F, device=local)
llr = tmp.awgn(F, pkg.sigma_for(2.0, 0.5), seed=2024, first_frame=rank * F)
link = B.host_link_probe(torch, ranks, F * n, tmp.info(pkg.INFO_FRAMES_PER_CTA) * 148 * n)
tmp.close()
if rank == 0:
    print(json.dumps(dict(what="host_link", n_gpus=world, **link)), flush=True)
for wc in (0, 1):
    for pin in (0, 1):
        if pin:
            per = max(1, len(all_cores) // world)
            os.sched_setaffinity(0, set(all_cores[local * per:(local + 1) * per]))
        else:
            os.sched_setaffinity(0, set(all_cores))
        for packed in (0, 1):
            dec = pkg.CGPUDecoder(code, nb_frames=F, device=local, out_format=packed)
            src = [pkg.PinnedArray((F, n), np.int8, write_combined=bool(wc)) for _ in range(2)]
            for s in src:
                s.array[:] = llr
            dst = pkg.PinnedArray((F, (n + 7) // 8 if packed else n), np.uint8)
            sec = B.time_e2e(torch, ranks, dec, src, dst, 10, 20, 3)
            fps = world * F / sec
            if rank == 0:
                print(json.dumps(dict(what="e2e", n_gpus=world, write_combined_input=wc, ranks_pinned_to_cores=pin, packed_output=packed, frames_per_s=fps,
                                      info_gbps=fps * k / 1e9, h2d_gbs_total=fps * n / 1e9, d2h_gbs_total=fps * ((n + 7) // 8 if packed else n) / 1e9)), flush=True)
            dec.close()
            for s in src:
                s.free()
            dst.free()
ranks.close()
