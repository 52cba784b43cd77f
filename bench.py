#!/usr/bin/env python3
"""bench.py — decoded info Gb/s at 10 iterations, int8 layered offset-min-sum, 576x288, 64 Ki frames per GPU (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W]            one JSON line (rank 0)
  python bench.py --impl reference ...                           the reference's own x86 decoder on the host cores, same line shape

A step = one pass of the hot path over one 65 536-frame batch of synthetic BPSK/AWGN frames (Eb/N0 = 2 dB) per GPU.
`value`     : device-resident inputs, the decode kernel(s) only, CUDA events on the launching stream, max over ranks.
`e2e`       : the same batch through the blocking C-ABI call ldpc_b200_decode with pinned HOST buffers — H2D, decode, D2H
              inside the timed region (the reference boundary: CGPU_Decoder_OMS_SIMD::decode, gpu_fixed/decoder_oms/...cu:97-149).
`roofline`  : HBM roofline of the decode kernel (LLR-in + bits-out bytes per frame) against MEASURED_PEAKS.json;
`sm_roofline`: the binding roof for on-chip codes — canonical scalar-int work 18 ops x I x M per frame against
              148 SM x 128 int lanes x clock.
`cpu_baseline`: the reference x86 SSE decoder (oracle/_ref, kind "reference") or the C port (kind "port") on all host cores, rank 0, N=1.
Multi-GPU: frames are independent -> each rank decodes its own batch, no collective on the data path (scaling "weak");
only the timing/BER counters are reduced.
"""
import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

CODE = "576x288"
FRAMES = 65536
ITERS = 10
EBN0 = 2.0
NBUF = 6          # input batches rotated between steps: 6 x 37.7 MB = 226 MB > 126 MB L2


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the decode kernel, from the newest committed `ncu --set full`
    summary of this workload (profiles/rNN_ncu_rp_v*.json, written by tools/ncu_summary.py).  None when no capture is committed."""
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    best = None
    for f in sorted((ROOT / "profiles").glob("r*_ncu_rp_v*.json")):
        try:
            l0 = json.loads(f.read_text())["launches"][0]
            tot = sum(float(l0[k]["value"]) * unit[l0[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            best = (tot, f.name)
        except Exception:
            continue
    return best


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line), via NVML."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample_once(self):
        if not self.nv:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            pass

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_run(llr: np.ndarray, min_seconds: float, threads: int):
    """Times the reference x86 SSE decoder (or the C port when oracle/_ref is absent) on `threads` host threads.
    Returns (frames_per_s, kind, sample description, hard decisions of the last pass)."""
    from _helpers import ref_x86, oracle_decode_mt, Code, default_params, ALGO
    F = llr.shape[0]
    hard = np.empty((F, llr.shape[1]), np.uint8)
    L = ref_x86(CODE)
    reps, spent = 0, 0.0
    if L is not None:
        kind = "reference"
        L.ref_x86_decode_mt(ALGO["OMS"], 1, llr.ctypes.data, hard.ctypes.data, min(F, 16 * threads * 4), ITERS, threads)   # warm-up
        while reps == 0 or spent < min_seconds:
            t = L.ref_x86_decode_mt(ALGO["OMS"], 1, llr.ctypes.data, hard.ctypes.data, F, ITERS, threads)
            if t < 0:
                raise RuntimeError("ref_x86_decode_mt failed")
            spent += t; reps += 1
    else:
        kind = "port"
        code = Code.load(CODE)
        prm = default_params()
        while reps == 0 or spent < min_seconds:
            t0 = time.perf_counter()
            hard = oracle_decode_mt(code, prm, llr, ITERS, threads)
            spent += time.perf_counter() - t0; reps += 1
    return F * reps / max(spent, 1e-9), kind, f"{F} frames x {reps} passes ({spent:.1f} s wall on {threads} threads)", hard


def claim_stdout():
    """Rank 0 must print ONE JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its version banner there
    when the communicator is created), so fd 1 is pointed at stderr for the whole run and the line goes to a private duplicate."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    out = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    steps, warmup = max(args.steps, 1), max(args.warmup, 3 if args.impl == "b200" else 0)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    F = args.frames

    import ldpcgputegra_b200 as pkg
    code = pkg.Code.load(CODE)
    k_info, n, m = code.k_info, code.n, code.m
    sigma = pkg.sigma_for(EBN0, k_info / n)
    config = {"workload": "int8 layered offset-min-sum (offset 1, x86-SSE semantics), 802.16e 576x288 rate 1/2, 10 iterations, "
                          "65536 synthetic BPSK/AWGN frames per GPU at Eb/N0 = 2 dB (BASELINE.json configs[1])",
              "code": CODE, "frames_per_gpu": F, "iterations": ITERS, "ebn0_db": EBN0, "semantics": "X86_SSE", "algo": "OMS",
              "l2": f"inputs rotate over {NBUF} distinct batches ({NBUF * F * n / 1e6:.0f} MB > 126 MB L2)", "parallelism": f"frame-sharded x{world}, no collective"}

    if args.impl == "reference":
        # the reference's own CPU implementation of the path, all host threads, rank 0 only
        if rank != 0:
            return 0
        from _helpers import awgn_llr
        threads = host_threads()
        llr = awgn_llr(code, F, EBN0, seed=2024)
        for _ in range(warmup):
            cpu_reference_run(llr[: 16 * threads * 8], 0.0, threads)
        t_tot, frames_tot, kind, sample = 0.0, 0, None, ""
        per_step = max(0.5, min(4.0, 60.0 / steps))
        for _ in range(steps):
            fps, kind, sample, _ = cpu_reference_run(llr, per_step, threads)
            t_tot += 1.0; frames_tot += fps
        fps = frames_tot / steps
        val = fps * k_info / 1e9
        line = {"impl": "reference", "metric": "decoded info throughput at 10 iterations", "value": val, "unit": "Gb/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * F / fps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "i8", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "Gb/s", "cores": threads, "kind": kind, "sample": sample + " per step"},
                "e2e": {"value": val, "unit": "Gb/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "frames_per_s": fps, "air_gbps": fps * n / 1e9}
        print(json.dumps(line), file=out, flush=True)
        return 0

    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the decoder has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    dec = pkg.CGPUDecoder(code, nb_frames=F, device=local_rank)      # defaults = the reference's: OMS, offset 1, 127/31
    kernel = dec.info(pkg.INFO_KERNEL)
    ts = torch.cuda.Stream()                 # a real (non-null) stream: the kernels and the timing events share it
    torch.cuda.set_stream(ts)
    stream = ts.cuda_stream
    d_llr = [torch.empty((F, n), dtype=torch.int8, device="cuda") for _ in range(NBUF)]
    d_hard = torch.empty((F, n), dtype=torch.uint8, device="cuda")
    for b in range(NBUF):   # every rank and every buffer gets its own frames of the counter-based generator
        dec.awgn_device(d_llr[b].data_ptr(), F, sigma, seed=2024, first_frame=(rank * NBUF + b) * F, stream=stream)
    torch.cuda.synchronize()

    # ---- value: device-resident, kernel only ----
    sampler = ClockSampler(local_rank)           # NVML is initialised BEFORE the warm-up so no idle gap precedes the timed region
    barrier()
    for i in range(warmup):
        dec.decode_device(d_llr[i % NBUF].data_ptr(), d_hard.data_ptr(), F, ITERS, stream=stream)
    sampler.start()
    l0 = dec.info(pkg.INFO_LAUNCHES)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        dec.decode_device(d_llr[i % NBUF].data_ptr(), d_hard.data_ptr(), F, ITERS, stream=stream)
    e1.record()
    sampler.sample_once()                        # the launches above are asynchronous: the GPU is inside the timed region here
    torch.cuda.synchronize()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = dec.info(pkg.INFO_LAUNCHES) - l0
    barrier()
    clocks = sampler.result()
    ms_per_step = ms_total / steps
    fps = world * F / (ms_per_step * 1e-3)
    be, fe = dec.count_errors_device(d_hard.data_ptr(), F, stream)

    # ---- e2e: pinned host buffers through the blocking C-ABI decode (H2D + decode + D2H inside the timed region) ----
    h_llr = [pkg.PinnedArray((F, n), np.int8) for _ in range(2)]
    h_hard = pkg.PinnedArray((F, n), np.uint8)
    for b in range(2):
        h_llr[b].array[:] = d_llr[b].cpu().numpy()
    for i in range(max(warmup, 3)):
        dec.decode(h_llr[i % 2].array, ITERS, out=h_hard.array)
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        dec.decode(h_llr[i % 2].array, ITERS, out=h_hard.array)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_fps = world * F * steps / e2e_s
    host_fe = int(h_hard.array[:, :k_info].any(axis=1).sum())

    # the same call with bit-packed output (the new 1-bit-per-bit format: D2H is 8x smaller) — reported beside the headline e2e
    decp = pkg.CGPUDecoder(code, nb_frames=F, device=local_rank, out_format=1)
    h_pack = pkg.PinnedArray((F, (n + 7) // 8), np.uint8)
    for i in range(3):
        decp.decode(h_llr[i % 2].array, ITERS, out=h_pack.array)
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        decp.decode(h_llr[i % 2].array, ITERS, out=h_pack.array)
    torch.cuda.synchronize()
    e2e_packed_fps = world * F * steps / max_over_ranks(time.perf_counter() - t0)
    barrier()
    packed_ok = bool(np.array_equal(np.unpackbits(h_pack.array, axis=1, bitorder="little")[:, :n], dec.decode(h_llr[(steps - 1) % 2].array, ITERS)))
    decp.close()

    hbm_peak, sm_max_mhz, peak_src = peaks()
    traffic = ncu_traffic() if F == FRAMES else None
    bytes_per_frame = n + n                      # int8 LLR in + one byte per bit out (the reference's output format)
    per_gpu_fps = fps / world
    achieved_gbs = per_gpu_fps * bytes_per_frame / 1e9
    int_peak = 148 * 128 * sm_max_mhz * 1e6      # scalar-int lanes x clock (both issue pipes, profiles/r01_pipe_microbench.jsonl)
    sm_ops = per_gpu_fps * ITERS * m * 18.0
    line = {"metric": "decoded info throughput at 10 iterations", "value": fps * k_info / 1e9, "unit": "Gb/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i8", "data": "synthetic",
            "config": config,
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                         "traffic": traffic[0] if traffic else None, "traffic_source": traffic[1] if traffic else None,
                         "algorithmic_bytes_per_launch": F * bytes_per_frame,
                         "peak_source": peak_src, "algorithmic_bytes_per_frame": bytes_per_frame,
                         "note": "on-chip state: HBM sees LLR-in + bits-out only, so this fraction is small by construction; the binding roof is sm_roofline"},
            "sm_roofline": {"bound": "sm_int_issue", "achieved": sm_ops / 1e12, "peak": int_peak / 1e12, "unit": "Tint-op/s", "frac": sm_ops / int_peak,
                            "edge_updates_per_s": per_gpu_fps * ITERS * m, "ops_per_edge_update": 18,
                            "note": "canonical scalar-int cost of SURVEY 8d; two frames per instruction (f16x2) may exceed 1.0"},
            "e2e": {"value": e2e_fps * k_info / 1e9, "unit": "Gb/s", "h2d_bytes_per_step": F * n, "d2h_bytes_per_step": F * n,
                    "frames_per_s": e2e_fps, "api": "ldpc_b200_decode (blocking, pinned host buffers, 4 stream slots)",
                    "bound": "PCIe: 37.7 MB each way per step; this box moves 41.7 GB/s per direction when both are busy (tools/pcie_probe.py) = 20.8 Gb/s",
                    "packed_output": {"value": e2e_packed_fps * k_info / 1e9, "unit": "Gb/s", "d2h_bytes_per_step": F * ((n + 7) // 8), "equals_byte_output": packed_ok}},
            "gpu_launches": int(launches), "kernel": {1: "frame-parallel (HBM state)", 2: "row-parallel on-chip", 3: "generic engine", 4: "frame-parallel, bulk-copy staged"}[kernel],
            "clocks": clocks, "frames_per_s": fps, "air_gbps": fps * n / 1e9,
            "ber_fer": {"frames": F, "bit_errors": be, "frame_errors": fe, "fer": fe / F, "e2e_frame_errors_last_batch": host_fe}}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        llr_host = np.ascontiguousarray(h_llr[0].array)
        cpu_fps, kind, sample, cpu_hard = cpu_reference_run(llr_host, 3.0, threads)
        line["cpu_baseline"] = {"value": cpu_fps * k_info / 1e9, "unit": "Gb/s", "cores": threads, "kind": kind, "sample": sample,
                                "frames_per_s": cpu_fps, "agrees_with_gpu": bool(np.array_equal(cpu_hard, dec.decode(llr_host, ITERS)))}
        # the kernel to beat: the reference's own gpu_fixed OMS kernel (unmodified, built for sm_100a into oracle/_ref), same workload,
        # in its own process so that nothing it does can touch this one's CUDA context
        try:
            import subprocess
            r = subprocess.run([sys.executable, str(ROOT / "tools" / "ref_gpu_time.py"), CODE, str(F), str(ITERS)], capture_output=True, text=True, timeout=180)
            line["reference_gpu_kernel"] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 and r.stdout.strip() else {"unavailable": (r.stderr or "no output")[-200:]}
        except Exception as e:          # the comparison is a courtesy, never a reason to lose the bench line
            line["reference_gpu_kernel"] = {"unavailable": str(e)[:200]}
    if rank == 0:
        print(json.dumps(line), file=out, flush=True)
    dec.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
