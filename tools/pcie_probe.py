#!/usr/bin/env python3
"""Raw pinned-memory PCIe rates on this box (diagnostic for bench.py's e2e number): H2D alone, D2H alone, both at once, in
chunks of the size ldpc_b200_decode() pipelines (one kernel wave of 576x288 = 6808 frames x 576 B)."""
import json, sys, time
import torch

def rate(fn, nbytes, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9

total = 65536 * 576
for chunk in (6808 * 576, 4 * 6808 * 576, total):
    n = max(1, total // chunk)
    h_in = torch.empty(total, dtype=torch.uint8).pin_memory(); h_out = torch.empty(total, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(total, dtype=torch.uint8, device="cuda"); d_out = torch.empty(total, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def h2d():
        with torch.cuda.stream(s1):
            for i in range(n): d_in[i*chunk:(i+1)*chunk].copy_(h_in[i*chunk:(i+1)*chunk], non_blocking=True)
    def d2h():
        with torch.cuda.stream(s2):
            for i in range(n): h_out[i*chunk:(i+1)*chunk].copy_(d_out[i*chunk:(i+1)*chunk], non_blocking=True)
    def both(): h2d(); d2h()
    print(json.dumps(dict(chunk_bytes=chunk, chunks=n, h2d_gbs=rate(h2d, n*chunk), d2h_gbs=rate(d2h, n*chunk), both_each_gbs=rate(both, n*chunk))))
