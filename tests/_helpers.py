"""Test-side access to the checker (oracle/) and, when built, the reference's own decoders (oracle/_ref).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
ORACLE_DIR = ROOT / "oracle"
REF_DIR = ORACLE_DIR / "_ref"

import sys
sys.path.insert(0, str(ROOT))
from ldpcgputegra_b200 import Code, CodeT, ParamsT, default_params, ALGO, SEM  # noqa: E402

_oracle = None


def build_oracle():
    r = subprocess.run(["make", "-C", str(ORACLE_DIR), "liboracle.so"], capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)


def oracle():
    global _oracle
    if _oracle is None:
        if not (ORACLE_DIR / "liboracle.so").exists():
            build_oracle()
        L = C.CDLL(str(ORACLE_DIR / "liboracle.so"))
        vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
        L.oracle_decode_fixed.restype = i32
        L.oracle_decode_fixed.argtypes = [C.POINTER(CodeT), C.POINTER(ParamsT), vp, vp, vp, vp, vp, sz, i32, i32]
        L.oracle_decode_fixed_mt.restype = i32
        L.oracle_decode_fixed_mt.argtypes = [C.POINTER(CodeT), C.POINTER(ParamsT), vp, vp, sz, i32, i32, i32]
        L.oracle_decode_float.restype = i32
        L.oracle_decode_float.argtypes = [C.POINTER(CodeT), C.POINTER(ParamsT), vp, vp, vp, vp, vp, sz, i32]
        L.oracle_quantize.restype = None
        L.oracle_quantize.argtypes = [vp, vp, sz, i32, i32]
        L.oracle_pack_bits.restype = None
        L.oracle_pack_bits.argtypes = [vp, vp, sz, i32]
        _oracle = L
    return _oracle


def oracle_decode(code: Code, prm: ParamsT, llr: np.ndarray, iters: int, want_state=True, want_iters=True):
    """Returns dict(hard, post, msgs, iters) from the CPU restatement (int8 or int16 llr)."""
    llr = np.ascontiguousarray(llr)
    eb = llr.dtype.itemsize
    F = llr.shape[0]
    hard = np.empty((F, code.n), np.uint8)
    post = np.empty((F, code.n), llr.dtype) if want_state else None
    msgs = np.empty((F, code.m), llr.dtype) if want_state else None
    it = np.empty(F, np.uint8) if want_iters else None
    c = code.c_struct()
    rc = oracle().oracle_decode_fixed(C.byref(c), C.byref(prm), llr.ctypes.data, hard.ctypes.data,
                                      post.ctypes.data if want_state else None, msgs.ctypes.data if want_state else None,
                                      it.ctypes.data if want_iters else None, F, iters, eb)
    if rc:
        raise RuntimeError(f"oracle_decode_fixed -> {rc}")
    return dict(hard=hard, post=post, msgs=msgs, iters=it)


def oracle_decode_float(code: Code, prm: ParamsT, llr: np.ndarray, iters: int):
    """Float min-sum of the CPU restatement (own definition, unpinned): dict(hard, post, msgs, iters)."""
    llr = np.ascontiguousarray(llr, np.float32)
    F = llr.shape[0]
    hard = np.empty((F, code.n), np.uint8); post = np.empty((F, code.n), np.float32); msgs = np.empty((F, code.m), np.float32)
    it = np.empty(F, np.uint8)
    c = code.c_struct()
    rc = oracle().oracle_decode_float(C.byref(c), C.byref(prm), llr.ctypes.data, hard.ctypes.data, post.ctypes.data, msgs.ctypes.data, it.ctypes.data, F, iters)
    if rc:
        raise RuntimeError(f"oracle_decode_float -> {rc}")
    return dict(hard=hard, post=post, msgs=msgs, iters=it)


def oracle_decode_mt(code: Code, prm: ParamsT, llr: np.ndarray, iters: int, threads: int):
    llr = np.ascontiguousarray(llr)
    F = llr.shape[0]
    hard = np.empty((F, code.n), np.uint8)
    c = code.c_struct()
    rc = oracle().oracle_decode_fixed_mt(C.byref(c), C.byref(prm), llr.ctypes.data, hard.ctypes.data, F, iters, llr.dtype.itemsize, threads)
    if rc:
        raise RuntimeError(f"oracle_decode_fixed_mt -> {rc}")
    return hard


def oracle_quantize(y: np.ndarray, scale=8, sat=31):
    y = np.ascontiguousarray(y, np.float32)
    q = np.empty(y.shape, np.int8)
    oracle().oracle_quantize(y.ctypes.data, q.ctypes.data, y.size, scale, sat)
    return q


def oracle_pack(hard: np.ndarray, n: int):
    F = hard.shape[0]
    out = np.empty((F, (n + 7) // 8), np.uint8)
    oracle().oracle_pack_bits(np.ascontiguousarray(hard).ctypes.data, out.ctypes.data, F, n)
    return out


# ---- code tables and parameters WITHOUT the product library (bench.py --impl reference must not map libldpc_b200.so) --------
def read_ldpc_table(name_or_path) -> Code:
    """Pure-Python reader of this repo's .ldpc table files (layout: ldpcgputegra_b200/csrc/code_table.cpp, save_code_table):
    8-byte magic, int32 {n, n_checks, m, nb_deg, deg[8], rows[8], index_bytes}, then m indices of 2 or 4 bytes."""
    p = Path(name_or_path)
    if not p.exists():
        p = ROOT / "ldpcgputegra_b200" / "codes" / f"{name_or_path}.ldpc"
    raw = p.read_bytes()
    hdr = np.frombuffer(raw, np.int32, 4 + 2 * 8 + 1, offset=8)
    n, n_checks, m, nb_deg = (int(x) for x in hdr[:4])
    ib = int(hdr[20])
    pos = np.frombuffer(raw, np.uint16 if ib == 2 else np.uint32, m, offset=8 + 4 * 21).astype(np.uint32)
    return Code(n, n_checks, [int(x) for x in hdr[4:4 + nb_deg]], [int(x) for x in hdr[12:12 + nb_deg]], pos)


def reference_default_params() -> ParamsT:
    """ldpc_b200_default_params restated in Python (ref: code/x86/main_p.cpp:90-104,133-139) for callers that must not load the product."""
    p = ParamsT()
    p.algo, p.schedule, p.dtype, p.semantics = ALGO["OMS"], 0, 0, SEM["X86_SSE"]
    p.offset, p.factor_q5, p.factor1, p.factor2 = 1, 29, 0.75, 0.875
    p.sat_var, p.sat_msg, p.llr_scale, p.sat_llr = 127, 31, 8, 31
    return p


# ---- the reference's own decoders, compiled from /root/reference into oracle/_ref (optional) -----------------------
def ref_x86(code_name: str):
    p = REF_DIR / f"libref_x86_{code_name}.so"
    if not p.exists():
        return None
    L = C.CDLL(str(p))
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    L.ref_x86_info.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.ref_x86_table.argtypes = [vp]
    L.ref_x86_decode.restype = i32
    L.ref_x86_decode.argtypes = [i32, i32, vp, vp, vp, vp, sz, i32]
    L.ref_x86_decode_mt.restype = C.c_double
    L.ref_x86_decode_mt.argtypes = [i32, i32, vp, vp, sz, i32, i32]
    return L


def ref_x86_code(L) -> Code:
    info = (C.c_int * 4)(); deg = (C.c_int * 8)(); rows = (C.c_int * 8)()
    L.ref_x86_info(info, deg, rows)
    pos = np.empty(info[2], np.uint32)
    L.ref_x86_table(pos.ctypes.data)
    return Code(info[0], info[1], list(deg[: info[3]]), list(rows[: info[3]]), pos)


def ref_x86_decode(L, algo: str, param: int, llr: np.ndarray, iters: int):
    llr = np.ascontiguousarray(llr, np.int8)
    F, n = llr.shape
    assert F % 16 == 0
    info = (C.c_int * 4)(); deg = (C.c_int * 8)(); rows = (C.c_int * 8)()
    L.ref_x86_info(info, deg, rows)
    hard = np.empty((F, n), np.uint8); post = np.empty((F, n), np.int8); msgs = np.empty((F, info[2]), np.int8)
    rc = L.ref_x86_decode(ALGO[algo], param, llr.ctypes.data, hard.ctypes.data, post.ctypes.data, msgs.ctypes.data, F, iters)
    if rc:
        raise RuntimeError(f"ref_x86_decode -> {rc}")
    return dict(hard=hard, post=post, msgs=msgs)


def ref_gpu(code_name: str):
    """The reference's own gpu_fixed kernels cross-compiled for sm_100a (oracle/_ref/libref_gpu_<code>.so); needs a GPU to run."""
    p = REF_DIR / f"libref_gpu_{code_name}.so"
    if not p.exists():
        return None
    L = C.CDLL(str(p))
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    L.ref_gpu_info.argtypes = [C.POINTER(C.c_int)]
    L.ref_gpu_table.argtypes = [vp]
    L.ref_gpu_decode.restype = i32
    L.ref_gpu_decode.argtypes = [i32, vp, vp, vp, vp, sz, i32, C.POINTER(C.c_float)]
    return L


def ref_gpu_decode(L, algo: str, llr: np.ndarray, iters: int, want_state=True):
    """Runs the reference kernel launch sequence (H2D, Interleaver_uint8, LDPC_Sched_Stage_1_*_SIMD, InvInterleaver_uint8, D2H).
    Returns dict(hard, post, msgs, kernel_ms, total_ms).  frames % 512 == 0."""
    llr = np.ascontiguousarray(llr, np.int8)
    F, n = llr.shape
    info = (C.c_int * 4)(); L.ref_gpu_info(info)
    assert n == info[0] and F % 512 == 0
    hard = np.empty((F, n), np.uint8)
    post = np.empty((F, n), np.int8) if want_state else None
    msgs = np.empty((F, info[2]), np.int8) if want_state else None
    ms = (C.c_float * 2)()
    rc = L.ref_gpu_decode(ALGO[algo], llr.ctypes.data, hard.ctypes.data, post.ctypes.data if want_state else None,
                          msgs.ctypes.data if want_state else None, F, iters, ms)
    if rc:
        raise RuntimeError(f"ref_gpu_decode -> {rc}")
    return dict(hard=hard, post=post, msgs=msgs, kernel_ms=float(ms[0]), total_ms=float(ms[1]))


def ref_arm(code_name: str):
    p = REF_DIR / f"libref_arm_{code_name}.so"
    if not p.exists():
        return None
    L = C.CDLL(str(p))
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    L.ref_arm_decode.restype = i32
    L.ref_arm_decode.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, vp, sz, i32]
    return L


def ref_arm_decode(L, code: Code, offset, sat_var, sat_msg, early, llr: np.ndarray, iters: int):
    llr = np.ascontiguousarray(llr, np.int8)
    F, n = llr.shape
    hard = np.empty((F, n), np.uint8); post = np.empty((F, n), np.int16); msgs = np.empty((F, code.m), np.int16); it = np.empty(F, np.uint8)
    rc = L.ref_arm_decode(offset, sat_var, sat_msg, int(early), llr.ctypes.data, hard.ctypes.data, post.ctypes.data, msgs.ctypes.data, it.ctypes.data, F, iters)
    if rc:
        raise RuntimeError(f"ref_arm_decode -> {rc}")
    return dict(hard=hard, post=post, msgs=msgs, iters=it)


# ---- synthetic inputs: all-zero codeword, BPSK 0 -> -1, AWGN, q = clamp((int)(8y), -31, 31) ------------
def awgn_llr(code: Code, frames: int, ebn0_db: float, seed: int) -> np.ndarray:
    rate = (code.n - code.n_checks) / code.n
    sigma = np.sqrt(10.0 ** (-(ebn0_db + 10.0 * np.log10(rate)) / 10.0) / 2.0)
    rng = np.random.Generator(np.random.Philox(seed))
    y = (-1.0 + sigma * rng.standard_normal((frames, code.n))).astype(np.float32)
    return oracle_quantize(y)


def stress_llr(code: Code, frames: int, seed: int, full_range=False) -> np.ndarray:
    """Uniform LLRs with random bias/spread per frame — drives posteriors onto the rails and exercises the x86 class>=1 abs quirk."""
    rng = np.random.Generator(np.random.Philox(seed))
    bias = rng.integers(-11, 1, size=(frames, 1))
    spread = rng.integers(10, 50, size=(frames, 1))
    q = bias + (rng.random((frames, code.n)) - 0.5) * 2 * spread
    lim = 127 if full_range else 31
    q = np.clip(np.trunc(q), -lim, lim).astype(np.int8)
    if full_range:
        q[rng.random(q.shape) < 0.02] = -128
        q[rng.random(q.shape) < 0.02] = 127
    return q
