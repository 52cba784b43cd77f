#!/usr/bin/env python3
"""bench.py — decoded info Gb/s at 10 iterations, int8 layered offset-min-sum, 576x288, 64 Ki frames per GPU (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W]            one JSON line (rank 0)
  python bench.py --impl reference ...                           the reference's own x86 decoder on the host cores, same line shape

A step = one pass of the hot path over one 65 536-frame batch of synthetic BPSK/AWGN frames (Eb/N0 = 2 dB) per GPU.
`value`     : device-resident inputs, the decode kernel(s) only, CUDA events on the launching stream, max over ranks.
`e2e`       : the same batch through the blocking C-ABI call ldpc_b200_decode with pinned HOST buffers — H2D, decode, D2H
              inside the timed region (the reference boundary: CGPU_Decoder_OMS_SIMD::decode, gpu_fixed/decoder_oms/...cu:97-149);
              `e2e.host_link` is the pinned-copy rate of THIS box measured in the same process with every rank copying both ways
              at once, `e2e.ceiling` what that rate allows, `e2e.packed_output` the same call with the bit-packed output format,
              `e2e.async_slots` the same buffers through the non-blocking slot API (four batches in flight).
`roofline`  : HBM roofline of the decode kernel (LLR-in + bits-out bytes per frame) against MEASURED_PEAKS.json;
`sm_roofline`: the binding roof for on-chip codes — canonical scalar-int work 18 ops x I x M per frame against
              148 SM x 128 int lanes x clock.
`cpu_baseline`: the reference x86 SSE decoder (oracle/_ref, kind "reference") or the C port (kind "port") on all host cores, rank 0, N=1.
`extra_configs`: BASELINE.json configs[4] (DVB-S2 64800x32400, HBM-resident state staged by the bulk-copy engine: the balanced
              303 104-frame batch AND the 65 536-frame batch) and configs[2] (float normalised min-sum, flooding schedule, per-frame
              syndrome stop, 576x288, 64 Ki frames), measured in the same process after the flagship, each with its own roofline,
              e2e and clocks.  `value` stays on configs[1].
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
DVBS2 = "64800x32400"
DVBS2_FRAMES = 303104        # 148 SMs x 2 CTAs x 256 consumer threads x 4 frames: every SM carries the same load
FLOAT_ITERS = 10             # configs[2] at the flagship's iteration budget (the stop criterion may end a frame earlier)
# scalar operations per edge and iteration of the float flooding decoder with the stop criterion: check-node pass
# 12 (sub, abs, 2-smallest search 3, sign flag 2; select, sign 3, store-side negate 1), variable-node pass 2 (add, amortised clamp),
# stop criterion 2 (compare, xor) — the float counterpart of SURVEY 8d's 18 integer operations per layered edge update
FLOAT_FLOODING_OPS_PER_EDGE = 16


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


def ncu_traffic(pattern="r*_ncu_rp_v*.json"):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the decode kernel, from the newest committed `ncu --set full`
    summary of this workload (profiles/rNN_ncu_*_v*.json, written by tools/ncu_summary.py).  None when no capture is committed."""
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    best = None
    for f in sorted((ROOT / "profiles").glob(pattern)):
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
        return {"sm_mhz": float(np.median(self.samples)), "sm_min_mhz": float(np.min(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_run(llr: np.ndarray, min_seconds: float, threads: int, code_name: str = CODE):
    """Times the reference x86 SSE decoder (or the C port when oracle/_ref is absent) on `threads` host threads.
    Returns (frames_per_s, kind, sample description, hard decisions of the last pass).  Touches oracle/ only — never the product library."""
    from _helpers import ref_x86, oracle_decode_mt, read_ldpc_table, reference_default_params, ALGO
    F = llr.shape[0]
    hard = np.empty((F, llr.shape[1]), np.uint8)
    L = ref_x86(code_name)
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
        code = read_ldpc_table(code_name)       # no product library on this path: pure-Python table reader, parameters restated
        prm = reference_default_params()
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


def flagship_config(world, F):
    return {"workload": "int8 layered offset-min-sum (offset 1, x86-SSE semantics), 802.16e 576x288 rate 1/2, 10 iterations, "
                        "65536 synthetic BPSK/AWGN frames per GPU at Eb/N0 = 2 dB (BASELINE.json configs[1])",
            "code": CODE, "frames_per_gpu": F, "iterations": ITERS, "ebn0_db": EBN0, "semantics": "X86_SSE", "algo": "OMS",
            "l2": f"inputs rotate over {NBUF} distinct batches ({NBUF * F * 576 / 1e6:.0f} MB > 126 MB L2)", "parallelism": f"frame-sharded x{world}, no collective"}


def reference_arm(args, out, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref, else the C port) on all host threads, rank 0
    only.  Nothing here loads libldpc_b200.so: the code table comes from the reference library itself (or the pure-Python reader of
    the .ldpc file), sigma and the quantiser from numpy and oracle/."""
    if rank != 0:
        return 0
    from _helpers import awgn_llr, ref_x86, ref_x86_code, read_ldpc_table
    L = ref_x86(CODE)
    code = ref_x86_code(L) if L is not None else read_ldpc_table(CODE)
    F, steps, warmup = args.frames, max(args.steps, 1), max(args.warmup, 0)
    k_info, n = code.n - code.n_checks, code.n
    threads = host_threads()
    llr = awgn_llr(code, F, EBN0, seed=2024)
    for _ in range(warmup):
        cpu_reference_run(llr[: 16 * threads * 8], 0.0, threads)
    frames_tot, kind, sample = 0.0, None, ""
    per_step = max(0.5, min(4.0, 60.0 / steps))
    for _ in range(steps):
        fps, kind, sample, _ = cpu_reference_run(llr, per_step, threads)
        frames_tot += fps
    fps = frames_tot / steps
    val = fps * k_info / 1e9
    line = {"impl": "reference", "metric": "decoded info throughput at 10 iterations", "value": val, "unit": "Gb/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * F / fps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "i8", "data": "synthetic", "config": flagship_config(world, F),
            "cpu_baseline": {"value": val, "unit": "Gb/s", "cores": threads, "kind": kind, "sample": sample + " per step"},
            "e2e": {"value": val, "unit": "Gb/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "frames_per_s": fps, "air_gbps": fps * n / 1e9}
    if os.environ.get("LDPC_BENCH_REPORT_MAPS"):      # tests/test_abi.py: which of this repo's shared objects did this process map?
        with open("/proc/self/maps") as f:
            line["mapped_objects"] = sorted({ln.split()[-1] for ln in f if ".so" in ln and str(ROOT) in ln})
    print(json.dumps(line), file=out, flush=True)
    return 0


class Ranks:
    """torch.distributed plumbing: barrier + max / sum over ranks (NCCL); a no-op at N = 1."""

    def __init__(self, torch, local_rank, world):
        self.torch, self.world, self.dist = torch, world, None
        if world > 1:
            import torch.distributed as dist_mod
            self.dist = dist_mod
            self.dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t.item())

    def max(self, x):
        return self._reduce(x, "MAX")

    def sum(self, x):
        return self._reduce(x, "SUM")

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def time_device(torch, pkg, ranks, dec, d_llrs, d_hard, F, iters, steps, warmup, stream, local_rank, d_iters=0):
    """`steps` device-resident decodes, CUDA events on the launching stream, max over ranks.  Returns (ms per step, launches, clocks)."""
    sampler = ClockSampler(local_rank)           # NVML is initialised BEFORE the warm-up so no idle gap precedes the timed region
    ranks.barrier()
    for i in range(warmup):
        dec.decode_device(d_llrs[i % len(d_llrs)].data_ptr(), d_hard.data_ptr(), F, iters, d_iters=d_iters, stream=stream)
    sampler.start()
    l0 = dec.info(pkg.INFO_LAUNCHES)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        dec.decode_device(d_llrs[i % len(d_llrs)].data_ptr(), d_hard.data_ptr(), F, iters, d_iters=d_iters, stream=stream)
    e1.record()
    sampler.sample_once()                        # the launches above are asynchronous: the GPU is inside the timed region here
    torch.cuda.synchronize()
    ms_total = ranks.max(e0.elapsed_time(e1))
    launches = dec.info(pkg.INFO_LAUNCHES) - l0
    ranks.barrier()
    return ms_total / steps, int(launches), sampler.result()


def time_e2e(torch, ranks, dec, h_llrs, h_out, iters, steps, warmup):
    """`steps` blocking host-buffer decodes (H2D + decode + D2H inside), wall clock around the loop, max over ranks."""
    for i in range(warmup):
        dec.decode(h_llrs[i % len(h_llrs)].array, iters, out=h_out.array)
    ranks.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        dec.decode(h_llrs[i % len(h_llrs)].array, iters, out=h_out.array)
    torch.cuda.synchronize()
    s = ranks.max(time.perf_counter() - t0)
    ranks.barrier()
    return s / steps


def time_e2e_async(torch, ranks, dec, h_llrs, h_outs, iters, steps, warmup):
    """The same host-buffer decodes through the NON-blocking slot API (ldpc_b200_decode_async / ldpc_b200_sync): step i goes to stream
    slot i % 4 as soon as that slot's previous batch has been waited for, so up to four batches are in flight and one batch's copies
    overlap another's kernel.  Every step still moves its input H2D and its result D2H; wall clock around the loop, max over ranks."""
    S = len(h_outs)
    def loop(k):
        for i in range(k):
            dec.sync(i % S)                                   # the result of step i - S is in h_outs[i % S] from here on
            dec.decode_async(i % S, h_llrs[i % len(h_llrs)].array, h_outs[i % S].array, iters)
        dec.sync()
    loop(warmup)
    ranks.barrier()
    t0 = time.perf_counter()
    loop(steps)
    torch.cuda.synchronize()
    s = ranks.max(time.perf_counter() - t0)
    ranks.barrier()
    return s / steps


def host_link_probe(torch, ranks, nbytes, chunk, reps=8):
    """Pinned-memory copy rate of this box with EVERY rank copying at once (what e2e is bounded by): H2D alone, D2H alone, both
    directions together, in chunks of the size decode() pipelines.  GB/s per GPU and per direction, the slowest rank's."""
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda"); d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    n = max(1, nbytes // chunk)

    def h2d():
        with torch.cuda.stream(s1):
            for i in range(n):
                d_in[i * chunk:(i + 1) * chunk].copy_(h_in[i * chunk:(i + 1) * chunk], non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            for i in range(n):
                h_out[i * chunk:(i + 1) * chunk].copy_(d_out[i * chunk:(i + 1) * chunk], non_blocking=True)

    def rate(fns):
        for f in fns:
            f()
        ranks.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            for f in fns:
                f()
        torch.cuda.synchronize()
        dt = ranks.max(time.perf_counter() - t0)
        ranks.barrier()
        return n * chunk * reps / dt / 1e9

    r = {"h2d_gbs": rate([h2d]), "d2h_gbs": rate([d2h]), "both_each_gbs": rate([h2d, d2h]), "ranks_copying": ranks.world,
         "bytes": n * chunk, "chunk_bytes": chunk, "how": "pinned cudaMemcpyAsync, all ranks at once, slowest rank, per GPU and direction"}
    del h_in, h_out, d_in, d_out
    return r


def run_dvbs2(torch, pkg, ranks, rank, local_rank, F, steps, with_e2e, cpu_baseline):
    """BASELINE configs[4]: DVB-S2 64800x32400 in reference row order, int8 layered OMS, state resident in HBM and staged through
    shared memory by cp.async.bulk (kernel 4).  Algorithmic HBM bytes per frame: N in + N out + I*4*M."""
    world = ranks.world
    code = pkg.Code.load(DVBS2)
    n, m, k_info = code.n, code.m, code.k_info
    sigma = pkg.sigma_for(EBN0, k_info / n)
    dec = pkg.CGPUDecoder(code, nb_frames=F, device=local_rank)
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); stream = ts.cuda_stream
    d_llr = torch.empty((F, n), dtype=torch.int8, device="cuda")          # 4-20 GB: far beyond the L2, no rotation needed
    d_hard = torch.empty((F, n), dtype=torch.uint8, device="cuda")
    dec.awgn_device(d_llr.data_ptr(), F, sigma, seed=2024, first_frame=rank * F, stream=stream)
    torch.cuda.synchronize()
    ms, launches, clocks = time_device(torch, pkg, ranks, dec, [d_llr], d_hard, F, ITERS, steps, 3, stream, local_rank)
    be, fe = dec.count_errors_device(d_hard.data_ptr(), F, stream)
    kernel = dec.info(pkg.INFO_KERNEL)
    fps = world * F / (ms * 1e-3)
    hbm_peak, _, peak_src = peaks()
    bytes_per_frame = 2 * n + ITERS * 4 * m
    achieved = (fps / world) * bytes_per_frame / 1e9
    traffic = ncu_traffic("r02_ncu_fs_v*.json")
    res = {"workload": f"int8 layered offset-min-sum (x86-SSE semantics), DVB-S2 64800x32400 rate 1/2 in reference row order, 10 iterations, {F} synthetic "
                       "BPSK/AWGN frames per GPU at 2 dB, decoder state (291 KB per frame) resident in HBM (BASELINE.json configs[4])",
           "value": fps * k_info / 1e9, "unit": "Gb/s", "ms_per_step": ms, "steps": steps, "warmup": 3, "frames_per_gpu": F, "frames_per_s": fps, "air_gbps": fps * n / 1e9,
           "l2": f"state {F * (n + m) / 1e9:.0f} GB and inputs {F * n / 1e9:.1f} GB per GPU: far larger than the L2",
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "peak_source": peak_src,
                        "traffic": None, "traffic_note": (f"{traffic[1]}: {traffic[0] / 1e9:.0f} GB of DRAM traffic in the profiled launch (its iteration count is in the file)" if traffic else
                                                          "no 10-iteration capture: profiles/r01_ncu_fs_v3.txt (2 iterations, 303 104 frames) measured 442 GB = 0.85 x the algorithmic bytes of those iterations"),
                        "algorithmic_bytes_per_frame": bytes_per_frame, "algorithmic_bytes_per_launch": F * bytes_per_frame,
                        "note": "N in + N out + I*4*M (posterior and message read+write per edge, 1 B each, uncompressed); the SM-issue roof coincides for this code (SURVEY 8d)"},
           "gpu_launches": launches, "kernel": {1: "frame-parallel (HBM state)", 4: "frame-parallel, bulk-copy staged (cp.async.bulk + mbarrier ring)"}.get(kernel, str(kernel)),
           "clocks": clocks, "ber_fer": {"frames": F, "bit_errors": be, "frame_errors": fe}}
    del d_llr, d_hard
    if with_e2e:
        # blocking host call, pinned buffers: H2D + interleave + decode + de-interleave + D2H in the timed region
        h_llr, h_hard = pkg.PinnedArray((F, n), np.int8), pkg.PinnedArray((F, n), np.uint8)
        h_llr.array[:] = dec.awgn(F, sigma, seed=2024, first_frame=rank * F)
        s = time_e2e(torch, ranks, dec, [h_llr], h_hard, ITERS, steps, 2)
        res["e2e"] = {"value": world * F / s * k_info / 1e9, "unit": "Gb/s", "h2d_bytes_per_step": F * n, "d2h_bytes_per_step": F * n, "frames_per_step": F,
                      "api": "ldpc_b200_decode (blocking, pinned host buffers, batch quartered over the 4 stream slots)"}
        if cpu_baseline:
            from _helpers import ref_x86, ALGO
            L = ref_x86(DVBS2)
            threads = host_threads()
            if L is not None:
                Fc = 16 * threads * 2
                llr = np.ascontiguousarray(h_llr.array[:Fc]); hard = np.empty((Fc, n), np.uint8)
                L.ref_x86_decode_mt(ALGO["OMS"], 1, llr.ctypes.data, hard.ctypes.data, Fc, ITERS, threads)
                spent, reps = 0.0, 0
                while spent < 2.0:
                    spent += L.ref_x86_decode_mt(ALGO["OMS"], 1, llr.ctypes.data, hard.ctypes.data, Fc, ITERS, threads); reps += 1
                res["cpu_baseline"] = {"value": Fc * reps / spent * k_info / 1e9, "unit": "Gb/s", "cores": threads, "kind": "reference",
                                       "sample": f"{Fc} frames x {reps} passes ({spent:.1f} s wall on {threads} threads)",
                                       "agrees_with_gpu": bool(np.array_equal(hard, h_hard.array[:Fc]))}
        h_llr.free(); h_hard.free()
    dec.close()
    torch.cuda.empty_cache()
    return res


def run_float_flooding(torch, pkg, ranks, rank, local_rank, F, steps, warmup):
    """BASELINE configs[2]: float normalised min-sum (factor 0.75), flooding schedule, per-frame syndrome early termination, 576x288,
    64 Ki frames at 2 dB, iteration budget 10.  Roof: SM issue (state on chip) — DESIGN.md 3.2d: FLOAT_FLOODING_OPS_PER_EDGE scalar
    operations per edge and EXECUTED iteration (the stop criterion ends frames early) against 148 SM x 128 lanes x clock."""
    world = ranks.world
    code = pkg.Code.load(CODE)
    n, m, k_info = code.n, code.m, code.k_info
    sigma = pkg.sigma_for(EBN0, k_info / n)
    dec = pkg.CGPUDecoder(code, nb_frames=F, device=local_rank, dtype="F32", algo="NMS", factor1=0.75, schedule="FLOODING", early_term=1)
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); stream = ts.cuda_stream
    nbuf = 2                                                              # 2 x 151 MB of float LLRs > 126 MB L2
    d_llrs = [torch.empty((F, n), dtype=torch.float32, device="cuda") for _ in range(nbuf)]
    d_hard = torch.empty((F, n), dtype=torch.uint8, device="cuda")
    d_iters = torch.zeros(F, dtype=torch.uint8, device="cuda")
    for b in range(nbuf):
        dec.awgn_device(d_llrs[b].data_ptr(), F, sigma, seed=2024, first_frame=(rank * nbuf + b) * F, stream=stream)
    torch.cuda.synchronize()
    ms, launches, clocks = time_device(torch, pkg, ranks, dec, d_llrs, d_hard, F, FLOAT_ITERS, steps, warmup, stream, local_rank, d_iters=d_iters.data_ptr())
    be, fe = dec.count_errors_device(d_hard.data_ptr(), F, stream)
    mean_iters = float(d_iters.float().mean().item())
    kernel = dec.info(pkg.INFO_KERNEL)
    fps = world * F / (ms * 1e-3)
    hbm_peak, sm_max_mhz, peak_src = peaks()
    issue_peak = 148 * 128 * sm_max_mhz * 1e6
    ops = (fps / world) * mean_iters * m * FLOAT_FLOODING_OPS_PER_EDGE
    bytes_per_frame = 4 * n + n
    res = {"workload": f"float normalised min-sum (factor 0.75), flooding schedule, per-frame syndrome early termination, 576x288, iteration budget {FLOAT_ITERS}, "
                       f"{F} synthetic BPSK/AWGN frames per GPU at 2 dB (BASELINE.json configs[2]); parity unpinned: the reference has no float or flooding decoder",
           "value": fps * k_info / 1e9, "unit": "Gb/s", "ms_per_step": ms, "steps": steps, "warmup": warmup, "frames_per_gpu": F, "frames_per_s": fps, "dtype": "f32",
           "mean_iterations_executed": mean_iters, "l2": f"inputs rotate over {nbuf} batches ({nbuf * F * n * 4 / 1e6:.0f} MB > 126 MB L2)",
           "roofline": {"bound": "sm_issue", "achieved": ops / 1e12, "peak": issue_peak / 1e12, "unit": "Top/s", "frac": ops / issue_peak,
                        "ops_per_edge_iteration": FLOAT_FLOODING_OPS_PER_EDGE, "edge_iterations_per_s": (fps / world) * mean_iters * m,
                        "note": "state on chip: check-node pass 12 + variable-node pass 2 + stop criterion 2 scalar operations per edge and executed iteration (DESIGN.md 3.2d) "
                                "against 148 SM x 128 lanes x clock"},
           "hbm_roofline": {"bound": "hbm", "achieved": (fps / world) * bytes_per_frame / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": (fps / world) * bytes_per_frame / 1e9 / hbm_peak,
                            "algorithmic_bytes_per_frame": bytes_per_frame, "peak_source": peak_src},
           "gpu_launches": launches, "kernel": {3: "generic engine (HBM state)", 5: "generic engine, on-chip state, a CTA owns F frames", 6: "generic engine, on-chip state, one warp per frame (work queue)"}.get(kernel, str(kernel)),
           "clocks": clocks, "ber_fer": {"frames": F, "bit_errors": be, "frame_errors": fe, "fer": fe / F}}
    # e2e: float LLRs in from pinned host memory (4 bytes per bit), hard-decision bytes out
    h_llr, h_hard = pkg.PinnedArray((F, n), np.float32), pkg.PinnedArray((F, n), np.uint8)
    h_llr.array[:] = d_llrs[0].cpu().numpy()
    s = time_e2e(torch, ranks, dec, [h_llr], h_hard, FLOAT_ITERS, steps, 2)
    res["e2e"] = {"value": world * F / s * k_info / 1e9, "unit": "Gb/s", "h2d_bytes_per_step": F * n * 4, "d2h_bytes_per_step": F * n, "api": "ldpc_b200_decode (blocking, pinned host buffers)"}
    h_llr.free(); h_hard.free()
    del d_llrs, d_hard, d_iters
    dec.close()
    torch.cuda.empty_cache()
    return res


def main():
    out = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip extra_configs (DVB-S2, float flooding)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":        # decided BEFORE anything touches the product library
        return reference_arm(args, out, rank, world)

    steps, warmup = max(args.steps, 1), max(args.warmup, 3)
    F = args.frames
    import torch
    import ldpcgputegra_b200 as pkg
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the decoder has no CPU fallback")
    torch.cuda.set_device(local_rank)
    ranks = Ranks(torch, local_rank, world)
    code = pkg.Code.load(CODE)
    k_info, n, m = code.k_info, code.n, code.m
    sigma = pkg.sigma_for(EBN0, k_info / n)
    config = flagship_config(world, F)

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
    ms_per_step, launches, clocks = time_device(torch, pkg, ranks, dec, d_llr, d_hard, F, ITERS, steps, warmup, stream, local_rank)
    fps = world * F / (ms_per_step * 1e-3)
    be, fe = dec.count_errors_device(d_hard.data_ptr(), F, stream)

    # ---- e2e: pinned host buffers through the blocking C-ABI decode (H2D + decode + D2H inside the timed region) ----
    h_llr = [pkg.PinnedArray((F, n), np.int8) for _ in range(2)]
    h_hard = pkg.PinnedArray((F, n), np.uint8)
    for b in range(2):
        h_llr[b].array[:] = d_llr[b].cpu().numpy()
    e2e_s = time_e2e(torch, ranks, dec, h_llr, h_hard, ITERS, steps, max(warmup, 3))
    e2e_fps = world * F / e2e_s
    host_fe = int(h_hard.array[:, :k_info].any(axis=1).sum())

    # the non-blocking slot API on the same buffers: four batches in flight (the reference's decode_stream, CGPU_Decoder_MS_SIMD.cu:219-275,
    # copies synchronously; this one does not)
    h_outs = [h_hard] + [pkg.PinnedArray((F, n), np.uint8) for _ in range(3)]
    n_async = max(steps, 12)
    e2e_async_fps = world * F / time_e2e_async(torch, ranks, dec, h_llr, h_outs, ITERS, n_async, 4)
    async_ok = all(bool(np.array_equal(h_outs[(n_async - 1 - k) % 4].array, dec.decode(h_llr[(n_async - 1 - k) % 2].array, ITERS))) for k in range(2))
    for h in h_outs[1:]:
        h.free()

    # the same call with bit-packed output (the new 1-bit-per-bit format: D2H is 8x smaller) — the second, first-class e2e
    decp = pkg.CGPUDecoder(code, nb_frames=F, device=local_rank, out_format=1)
    h_pack = pkg.PinnedArray((F, (n + 7) // 8), np.uint8)
    e2e_packed_fps = world * F / time_e2e(torch, ranks, decp, h_llr, h_pack, ITERS, steps, 3)
    packed_ok = bool(np.array_equal(np.unpackbits(h_pack.array, axis=1, bitorder="little")[:, :n], dec.decode(h_llr[(steps - 1) % 2].array, ITERS)))
    decp.close()

    # what the host side of THIS box can move with every rank copying at once: the ceiling e2e is graded against
    link = host_link_probe(torch, ranks, F * n, dec.info(pkg.INFO_FRAMES_PER_CTA) * 148 * n)
    per_gpu_kernel_fps = fps / world
    ceil_bytes_fps = world * min(per_gpu_kernel_fps, link["both_each_gbs"] * 1e9 / n)                 # n bytes in and n bytes out per frame, both directions busy
    ceil_packed_fps = world * min(per_gpu_kernel_fps, link["h2d_gbs"] * 1e9 / n)                      # packed: the way out is 8x smaller, the way in binds

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
                    "host_link": link,
                    "ceiling": {"value": ceil_bytes_fps * k_info / 1e9, "unit": "Gb/s", "frac_of_ceiling": e2e_fps / ceil_bytes_fps,
                                "how": f"{world} x min(kernel rate, host_link.both_each_gbs / {n} B): every frame moves {n} B in and {n} B of byte-per-bit decisions out, "
                                       "both directions busy; measured with all ranks copying at once, same process"},
                    "bound": f"host link of this box with {world} rank(s) copying at once: {link['both_each_gbs']:.1f} GB/s per GPU and direction (host_link)",
                    "async_slots": {"value": e2e_async_fps * k_info / 1e9, "unit": "Gb/s", "frames_per_s": e2e_async_fps, "steps": n_async, "in_flight": 4,
                                    "frac_of_ceiling": e2e_async_fps / ceil_bytes_fps, "equals_blocking_call": async_ok,
                                    "api": "ldpc_b200_decode_async + ldpc_b200_sync, one batch per stream slot, same pinned host buffers and bytes per step"},
                    "packed_output": {"value": e2e_packed_fps * k_info / 1e9, "unit": "Gb/s", "h2d_bytes_per_step": F * n, "d2h_bytes_per_step": F * ((n + 7) // 8),
                                      "frames_per_s": e2e_packed_fps, "equals_byte_output": packed_ok,
                                      "ceiling": {"value": ceil_packed_fps * k_info / 1e9, "unit": "Gb/s", "frac_of_ceiling": e2e_packed_fps / ceil_packed_fps,
                                                  "how": f"{world} x min(kernel rate, host_link.h2d_gbs / {n} B)"},
                                      "api": "ldpc_b200_decode with out_format = LDPC_OUT_PACKED (LSB-first bits; the reference has no packed format)"}},
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
    dec.close()
    for h in h_llr:
        h.free()
    h_hard.free(); h_pack.free()
    del d_llr, d_hard
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs on the same line (every rank takes part: the collectives inside must match) ----
    if not args.no_extra and F == FRAMES:
        extra = {}
        want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
        for key, fn in (("dvbs2", lambda: run_dvbs2(torch, pkg, ranks, rank, local_rank, DVBS2_FRAMES, 3, False, False)),
                        ("dvbs2_64k", lambda: run_dvbs2(torch, pkg, ranks, rank, local_rank, 65536, 3, True, want_cpu)),
                        ("float_flooding_et", lambda: run_float_flooding(torch, pkg, ranks, rank, local_rank, FRAMES, min(steps, 20), 3))):
            try:
                extra[key] = fn()
            except Exception as e:      # an extra config must never cost the flagship line
                extra[key] = {"error": repr(e)[:300]}
                if world > 1:           # the ranks are no longer in step: stop here rather than hang in a collective
                    break
        line["extra_configs"] = extra
    if rank == 0:
        print(json.dumps(line), file=out, flush=True)
    ranks.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
