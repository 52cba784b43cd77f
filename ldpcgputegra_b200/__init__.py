"""ldpcgputegra_b200 — host-side binding of the B200-native LDPC decoder.

The product is the C-ABI shared library ``libldpc_b200.so`` (``include/ldpc_b200.h``); this module is the thin
ctypes layer the tests, ``bench.py`` and Python users call it through.  Class and method names mirror the reference's
decoder boundary so that code written against it reads the same:

* ``CGPUDecoder(nb_frames, code, ...)`` / ``.decode(llr, iterations)``  — ref: ``code/gpu_fixed/decoder_template/CGPUDecoder.h:20-37``
* ``setOffset`` / ``setFactor`` / ``setVarRange`` / ``setMsgRange`` keyword arguments — ref: ``code/x86/CDecoder/OMS/CDecoder_OMS_fixed_SSE.h:26-39``,
  ``code/x86/CDecoder/template/CDecoder_fixed.h:40-41``
* ``CreateDecoder(type, arch, format, ...)`` — ref: ``code/x86/CDecoder/DecoderLibrary.h:44-134``

There is no CPU fallback here: if the CUDA library is missing or no GPU is visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
_LIB_PATH = Path(os.environ.get("LDPC_B200_LIB", _PKG / "libldpc_b200.so"))   # override = A/B experiments only
CODES_DIR = _PKG / "codes"

MAX_DEG_CLASSES = 8
ALGO = {"MS": 0, "OMS": 1, "NMS": 2, "2NMS": 3}
SEM = {"X86_SSE": 0, "UNIFORM": 1, "ARM_SCALAR": 2, "GPU_FIXED": 3}
SCHED = {"LAYERED": 0, "FLOODING": 1}
DTYPE = {"I8": 0, "I16": 1, "F32": 2}
NP_DTYPE = {0: np.int8, 1: np.int16, 2: np.float32}
OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_IO, ERR_NOMEM, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
INFO_KERNEL, INFO_LEVELS, INFO_SMEM_BYTES, INFO_FRAMES_PER_CTA, INFO_LAUNCHES, INFO_STREAM_SLOTS, INFO_DEVICE, INFO_FS_STAIR_ROWS, INFO_FS_VARIANT = range(9)


class LdpcError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"ldpc_b200 status {status}: {msg}")
        self.status = status


class CodeT(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_checks", C.c_int32), ("m", C.c_int32), ("nb_deg", C.c_int32),
                ("deg", C.c_int32 * MAX_DEG_CLASSES), ("rows", C.c_int32 * MAX_DEG_CLASSES),
                ("pos", C.POINTER(C.c_uint32))]


class ParamsT(C.Structure):
    _fields_ = [("algo", C.c_int32), ("schedule", C.c_int32), ("dtype", C.c_int32), ("semantics", C.c_int32),
                ("offset", C.c_int32), ("factor_q5", C.c_int32), ("factor1", C.c_float), ("factor2", C.c_float),
                ("sat_var", C.c_int32), ("sat_msg", C.c_int32), ("llr_scale", C.c_int32), ("sat_llr", C.c_int32),
                ("early_term", C.c_int32), ("out_format", C.c_int32), ("kernel", C.c_int32), ("reserved", C.c_int32 * 5)]


_lib = None


def build(verbose: bool = False) -> None:
    """Compile libldpc_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-j8", "-C", str(_PKG / "csrc")], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("building libldpc_b200.so failed")


def lib() -> C.CDLL:
    """Load the C-ABI library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise ImportError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (the decoder has no CPU fallback)")
    L = C.CDLL(str(_LIB_PATH))
    vp, sz, i32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
    sig = {
        "ldpc_b200_abi_version": (C.c_int, []),
        "ldpc_b200_device_count": (C.c_int, []),
        "ldpc_b200_status_string": (C.c_char_p, [i32]),
        "ldpc_b200_default_params": (None, [C.POINTER(ParamsT)]),
        "ldpc_b200_load_code_header": (i32, [C.POINTER(CodeT), C.c_char_p, C.c_char_p]),
        "ldpc_b200_load_code_table": (i32, [C.POINTER(CodeT), C.c_char_p]),
        "ldpc_b200_save_code_table": (i32, [C.POINTER(CodeT), C.c_char_p]),
        "ldpc_b200_check_code": (i32, [C.POINTER(CodeT)]),
        "ldpc_b200_free_code": (None, [C.POINTER(CodeT)]),
        "ldpc_b200_level_schedule": (i32, [C.POINTER(CodeT), C.POINTER(C.c_int32)]),
        "ldpc_b200_create": (i32, [C.POINTER(vp), C.POINTER(CodeT), C.POINTER(ParamsT), i32, sz]),
        "ldpc_b200_destroy": (None, [vp]),
        "ldpc_b200_last_error": (C.c_char_p, [vp]),
        "ldpc_b200_get_info": (i32, [vp, i32, C.POINTER(C.c_int64)]),
        "ldpc_b200_quantize": (i32, [vp, vp, vp, sz]),
        "ldpc_b200_decode": (i32, [vp, vp, vp, sz, i32, vp]),
        "ldpc_b200_decode_async": (i32, [vp, i32, vp, vp, sz, i32, vp]),
        "ldpc_b200_sync": (i32, [vp, i32]),
        "ldpc_b200_host_alloc": (i32, [C.POINTER(vp), sz]),
        "ldpc_b200_host_alloc_input": (i32, [C.POINTER(vp), sz]),
        "ldpc_b200_host_free": (i32, [vp]),
        "ldpc_b200_device_alloc": (i32, [vp, C.POINTER(vp), sz]),
        "ldpc_b200_device_free": (i32, [vp, vp]),
        "ldpc_b200_decode_device": (i32, [vp, vp, vp, sz, i32, vp, vp]),
        "ldpc_b200_stream": (vp, [vp, i32]),
        "ldpc_b200_set_debug": (i32, [vp, i32]),
        "ldpc_b200_debug_state": (i32, [vp, vp, vp, sz]),
        "ldpc_b200_awgn_device": (i32, [vp, vp, sz, C.c_float, u64, u64, vp]),
        "ldpc_b200_awgn": (i32, [vp, vp, sz, C.c_float, u64, u64]),
        "ldpc_b200_count_errors_device": (i32, [vp, vp, sz, C.POINTER(u64), vp]),
        "ldpc_b200_encoder_create": (i32, [C.POINTER(vp), C.POINTER(CodeT), i32]),
        "ldpc_b200_encoder_destroy": (None, [vp]),
        "ldpc_b200_encoder_last_error": (C.c_char_p, [vp]),
        "ldpc_b200_encoder_info": (i32, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "ldpc_b200_encode": (i32, [vp, vp, vp, sz]),
        "ldpc_b200_encode_device": (i32, [vp, vp, vp, sz, u64, u64, vp]),
        "ldpc_b200_awgn_codeword_device": (i32, [vp, vp, vp, sz, C.c_float, u64, u64, vp]),
        "ldpc_b200_count_errors_ref_device": (i32, [vp, vp, vp, sz, C.POINTER(u64), vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


EXPORTS = ["ldpc_b200_abi_version", "ldpc_b200_device_count", "ldpc_b200_status_string", "ldpc_b200_default_params",
           "ldpc_b200_load_code_header", "ldpc_b200_load_code_table", "ldpc_b200_save_code_table", "ldpc_b200_check_code",
           "ldpc_b200_free_code", "ldpc_b200_level_schedule", "ldpc_b200_create", "ldpc_b200_destroy", "ldpc_b200_last_error",
           "ldpc_b200_get_info", "ldpc_b200_quantize", "ldpc_b200_decode", "ldpc_b200_decode_async", "ldpc_b200_sync",
           "ldpc_b200_host_alloc", "ldpc_b200_host_alloc_input", "ldpc_b200_host_free", "ldpc_b200_device_alloc", "ldpc_b200_device_free", "ldpc_b200_decode_device", "ldpc_b200_stream", "ldpc_b200_set_debug", "ldpc_b200_debug_state",
           "ldpc_b200_awgn_device", "ldpc_b200_awgn", "ldpc_b200_count_errors_device",
           "ldpc_b200_encoder_create", "ldpc_b200_encoder_destroy", "ldpc_b200_encoder_last_error", "ldpc_b200_encoder_info", "ldpc_b200_encode",
           "ldpc_b200_encode_device", "ldpc_b200_awgn_codeword_device", "ldpc_b200_count_errors_ref_device"]


def _check(status: int, handle=None):
    if status != OK:
        L = lib()
        msg = (L.ldpc_b200_last_error(handle) or b"").decode() or L.ldpc_b200_status_string(status).decode()
        raise LdpcError(status, msg)


class Code:
    """A parity-check code table (the reference's constantes header as data)."""

    def __init__(self, n, n_checks, deg, rows, pos):
        self.n, self.n_checks = int(n), int(n_checks)
        self.deg, self.rows = [int(d) for d in deg], [int(r) for r in rows]
        self.pos = np.ascontiguousarray(pos, dtype=np.uint32)
        self.m = int(self.pos.size)

    @property
    def k_info(self) -> int:
        return self.n - self.n_checks

    def c_struct(self) -> CodeT:
        c = CodeT()
        c.n, c.n_checks, c.m, c.nb_deg = self.n, self.n_checks, self.m, len(self.deg)
        for i, (d, r) in enumerate(zip(self.deg, self.rows)):
            c.deg[i], c.rows[i] = d, r
        c.pos = self.pos.ctypes.data_as(C.POINTER(C.c_uint32))
        return c

    @staticmethod
    def _from_c(c: CodeT) -> "Code":
        pos = np.ctypeslib.as_array(c.pos, shape=(c.m,)).copy()
        code = Code(c.n, c.n_checks, list(c.deg[: c.nb_deg]), list(c.rows[: c.nb_deg]), pos)
        lib().ldpc_b200_free_code(C.byref(c))
        return code

    @staticmethod
    def from_header(header: str, table: str | None = None) -> "Code":
        """Parse a reference header (x86: one file; gpu_fixed: constantes_gpu.h + constantes_decoder.h)."""
        c = CodeT()
        _check(lib().ldpc_b200_load_code_header(C.byref(c), str(header).encode(), str(table).encode() if table else None))
        return Code._from_c(c)

    @staticmethod
    def load(name_or_path: str) -> "Code":
        """Load one of the bundled tables by name ('576x288') or a .ldpc file by path."""
        p = Path(name_or_path)
        if not p.exists():
            p = CODES_DIR / f"{name_or_path}.ldpc"
        c = CodeT()
        _check(lib().ldpc_b200_load_code_table(C.byref(c), str(p).encode()))
        return Code._from_c(c)

    def save(self, path: str) -> None:
        c = self.c_struct()
        _check(lib().ldpc_b200_save_code_table(C.byref(c), str(path).encode()))

    def level_schedule(self):
        lv = np.zeros(self.n_checks, dtype=np.int32)
        c = self.c_struct()
        r = lib().ldpc_b200_level_schedule(C.byref(c), lv.ctypes.data_as(C.POINTER(C.c_int32)))
        if r < 0:
            _check(r)
        return r, lv


def default_params(**kw) -> ParamsT:
    p = ParamsT()
    lib().ldpc_b200_default_params(C.byref(p))
    for k, v in kw.items():
        if k == "algo" and isinstance(v, str):
            v = ALGO[v]
        if k == "semantics" and isinstance(v, str):
            v = SEM[v]
        if k == "schedule" and isinstance(v, str):
            v = SCHED[v]
        if k == "dtype" and isinstance(v, str):
            v = DTYPE[v]
        if k == "chunk_waves":    # waves per pipeline chunk of decode() (experiment knob)
            p.reserved[2] = v
            continue
        if k == "fs_stages":      # ring depth of the staged frame-parallel kernel (A/B experiments)
            p.reserved[4] = (p.reserved[4] & ~255) | int(v)
            continue
        if k == "fs_nc":          # consumer threads per CTA of the staged kernel: 128 | 256 (A/B experiments)
            p.reserved[4] = (p.reserved[4] & ~(15 << 8)) | ({128: 1, 256: 2}[int(v)] << 8)
            continue
        if k == "fs_tma":         # staged kernel: message lines of a row as one 2-D tensor-map copy: 0 = library's choice, 1 = never, 2 = always (A/B experiments)
            p.reserved[4] = (p.reserved[4] & ~(3 << 12)) | (int(v) << 12)
            continue
        if k == "fs_g4":          # staged kernel: posterior lines through tile::gather4: 0 = library's choice, 1 = never, 2 = always (A/B experiments)
            p.reserved[4] = (p.reserved[4] & ~(3 << 14)) | (int(v) << 14)
            continue
        if k == "fs_nostair":     # staged kernel: no register-carried staircase runs (every row through the forwarding ring, as in round 1): A/B experiments
            p.reserved[4] = (p.reserved[4] & ~(1 << 18)) | (int(bool(v)) << 18)
            continue
        if k == "fs_pipe2":       # staged kernel: paired staircase rows (small-batch instantiation): 0 = library's choice, 1 = never, 2 = always (A/B experiments)
            p.reserved[4] = (p.reserved[4] & ~(3 << 19)) | (int(v) << 19)
            continue
        if k == "fs_cmp":         # staged kernel: compressed messages (four words per row instead of one per edge): 0 = library's choice, 1 = never, 2 = always
            p.reserved[4] = (p.reserved[4] & ~(3 << 16)) | (int(v) << 16)
            continue
        if k == "no_pair_fastest":   # keep lane -> (pair t / nrows, row t % nrows) in the static plan (A/B experiments)
            p.reserved[3] = 5 if v else 0
            continue
        if k == "small_steps":    # keep the <= 32-row steps of the on-chip plan even when the code allows 64-128-row steps (A/B experiments)
            p.reserved[3] = 3 if v else 0
            continue
        if k == "no_static":      # force the descriptor-driven step loop of the on-chip kernel (A/B experiments)
            p.reserved[3] = int(v)
            continue
        if k == "group":          # (G warps, P pairs) experiment knob of the on-chip kernel
            p.reserved[0], p.reserved[1] = v
            continue
        setattr(p, k, v)
    return p


class PinnedArray:
    """numpy view over cudaMallocHost memory (ref: CTrame's pinned buffers, code/gpu_fixed/trame/CTrame.cpp:38-41)."""

    def __init__(self, shape, dtype, write_combined: bool = False):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = C.c_void_p()
        alloc = lib().ldpc_b200_host_alloc_input if write_combined else lib().ldpc_b200_host_alloc       # write-combined: input buffers the host only writes
        _check(alloc(C.byref(self._p), max(self.nbytes, 1)))
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            lib().ldpc_b200_host_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CGPUDecoder:
    """decode(frames, iterations) on one GPU.  One object per (GPU, host thread), like the reference's CGPUDecoder."""

    def __init__(self, code: Code, nb_frames: int = 65536, device: int = 0, params: ParamsT | None = None, **kw):
        self.code = code
        self.params = params if params is not None else default_params(**kw)
        self._h = C.c_void_p()
        self.np_dtype = NP_DTYPE.get(self.params.dtype, np.int8)     # element type of LLRs / posteriors / messages at the boundary
        self._inflight = {}                                          # slot -> arrays an asynchronous decode still reads / writes
        c = code.c_struct()
        _check(lib().ldpc_b200_create(C.byref(self._h), C.byref(c), C.byref(self.params), device, nb_frames))

    def close(self):
        if self._h:
            lib().ldpc_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def initialize(self):  # ref: CGPUDecoder::initialize — empty in every reference subclass
        return None

    def info(self, what: int) -> int:
        v = C.c_int64()
        _check(lib().ldpc_b200_get_info(self._h, what, C.byref(v)), self._h)
        return v.value

    @property
    def hard_row_bytes(self) -> int:
        return (self.code.n + 7) // 8 if self.params.out_format == 1 else self.code.n

    def quantize(self, y: np.ndarray) -> np.ndarray:
        y = np.ascontiguousarray(y, dtype=np.float32)
        q = np.empty(y.shape, dtype=np.int8)
        _check(lib().ldpc_b200_quantize(self._h, y.ctypes.data, q.ctypes.data, y.size), self._h)
        return q

    def _check_io(self, llr: np.ndarray, out: np.ndarray | None):
        """Shapes, dtypes and layout the C side takes on trust (it only sees pointers)."""
        if llr.size % self.code.n or (llr.ndim == 2 and llr.shape[1] != self.code.n):
            raise LdpcError(ERR_INVALID, f"llr must be [frames, {self.code.n}], got {llr.shape}")
        frames = llr.size // self.code.n
        if out is not None:
            if not isinstance(out, np.ndarray) or out.dtype != np.uint8 or not out.flags.c_contiguous or not out.flags.writeable:
                raise LdpcError(ERR_INVALID, "out must be a writable C-contiguous uint8 array")
            if out.nbytes < frames * self.hard_row_bytes:
                raise LdpcError(ERR_INVALID, f"out holds {out.nbytes} bytes, {frames} frames need {frames * self.hard_row_bytes}")
        return frames

    def decode(self, llr: np.ndarray, iterations: int, out: np.ndarray | None = None, want_iters: bool = False):
        """llr: [frames, n] (host) of the handle's dtype (int8 / int16 / float32).  Returns hard decisions [frames, n] bytes in
        {0,1} (or packed), optionally iteration counts."""
        llr = np.ascontiguousarray(llr, dtype=self.np_dtype)
        frames = self._check_io(llr, out)
        if out is None:
            out = np.empty((frames, self.hard_row_bytes), dtype=np.uint8)
        it = np.empty(frames, dtype=np.uint8) if want_iters else None
        _check(lib().ldpc_b200_decode(self._h, llr.ctypes.data, out.ctypes.data, frames, iterations, it.ctypes.data if want_iters else None), self._h)
        return (out, it) if want_iters else out

    def decode_async(self, slot: int, llr: np.ndarray, out: np.ndarray, iterations: int):
        """Asynchronous: the copies run after this returns.  The (possibly converted) input and the output are kept alive on the
        object until sync(slot); pass pinned arrays (PinnedArray) for the copies to overlap."""
        llr = np.ascontiguousarray(llr, dtype=self.np_dtype)
        frames = self._check_io(llr, out)
        if out is None:
            raise LdpcError(ERR_INVALID, "decode_async needs an output array")
        _check(lib().ldpc_b200_decode_async(self._h, slot, llr.ctypes.data, out.ctypes.data, frames, iterations, None), self._h)
        self._inflight.setdefault(slot, []).append((llr, out))

    decode_stream = decode_async  # ref: CGPUDecoder::decode_stream

    def sync(self, slot: int = -1):
        _check(lib().ldpc_b200_sync(self._h, slot), self._h)
        if slot < 0:
            self._inflight.clear()
        else:
            self._inflight.pop(slot, None)

    def stream(self, slot: int = 0) -> int:
        """cudaStream_t of a stream slot as an integer (ldpc_b200_stream)."""
        return int(lib().ldpc_b200_stream(self._h, slot) or 0)

    def decode_device(self, d_llr: int, d_hard: int, frames: int, iterations: int, d_iters: int = 0, stream: int = 0):
        _check(lib().ldpc_b200_decode_device(self._h, d_llr, d_hard, frames, iterations, d_iters or None, stream or None), self._h)

    def set_debug(self, on: bool = True):
        _check(lib().ldpc_b200_set_debug(self._h, int(on)), self._h)

    def debug_state(self, frames: int):
        post = np.empty((frames, self.code.n), dtype=self.np_dtype)
        msgs = np.empty((frames, self.code.m), dtype=self.np_dtype)
        _check(lib().ldpc_b200_debug_state(self._h, post.ctypes.data, msgs.ctypes.data, frames), self._h)
        return post, msgs

    def awgn(self, frames: int, sigma: float, seed: int, first_frame: int = 0) -> np.ndarray:
        q = np.empty((frames, self.code.n), dtype=self.np_dtype)
        _check(lib().ldpc_b200_awgn(self._h, q.ctypes.data, frames, sigma, seed, first_frame), self._h)
        return q

    def awgn_device(self, d_llr: int, frames: int, sigma: float, seed: int, first_frame: int = 0, stream: int = 0):
        _check(lib().ldpc_b200_awgn_device(self._h, d_llr, frames, sigma, seed, first_frame, stream or None), self._h)

    def awgn_codeword_device(self, d_llr: int, d_codeword: int, frames: int, sigma: float, seed: int, first_frame: int = 0, stream: int = 0):
        _check(lib().ldpc_b200_awgn_codeword_device(self._h, d_llr, d_codeword, frames, sigma, seed, first_frame, stream or None), self._h)

    def count_errors_ref_device(self, d_hard: int, d_codeword: int, frames: int, stream: int = 0):
        out = (C.c_uint64 * 2)()
        _check(lib().ldpc_b200_count_errors_ref_device(self._h, d_hard, d_codeword, frames, out, stream or None), self._h)
        return int(out[0]), int(out[1])

    def count_errors_device(self, d_hard: int, frames: int, stream: int = 0):
        out = (C.c_uint64 * 2)()
        _check(lib().ldpc_b200_count_errors_device(self._h, d_hard, frames, out, stream or None), self._h)
        return int(out[0]), int(out[1])


class Encoder:
    """Systematic encoder derived from H (ref: the `-encoder` option / GenericEncoder, code/x86/CEncoder/GenericEncoder.cpp:38-78)."""

    def __init__(self, code: Code, device: int = 0):
        self.code = code
        self._e = C.c_void_p()
        c = code.c_struct()
        st = lib().ldpc_b200_encoder_create(C.byref(self._e), C.byref(c), device)
        if st != OK:
            raise LdpcError(st, (lib().ldpc_b200_encoder_last_error(None) or b"").decode())

    def close(self):
        if self._e:
            lib().ldpc_b200_encoder_destroy(self._e)
            self._e = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        a, b = C.c_int(), C.c_int()
        lib().ldpc_b200_encoder_info(self._e, C.byref(a), C.byref(b))
        return a.value, b.value

    def encode(self, info_bits: np.ndarray) -> np.ndarray:
        info_bits = np.ascontiguousarray(info_bits, dtype=np.uint8)
        frames = info_bits.shape[0]
        cw = np.empty((frames, self.code.n), dtype=np.uint8)
        st = lib().ldpc_b200_encode(self._e, info_bits.ctypes.data, cw.ctypes.data, frames)
        if st != OK:
            raise LdpcError(st, (lib().ldpc_b200_encoder_last_error(self._e) or b"").decode())
        return cw

    def encode_device(self, d_codeword: int, frames: int, seed: int = 0, first_frame: int = 0, d_info: int = 0, stream: int = 0):
        st = lib().ldpc_b200_encode_device(self._e, d_info or None, d_codeword, frames, seed, first_frame, stream or None)
        if st != OK:
            raise LdpcError(st, (lib().ldpc_b200_encoder_last_error(self._e) or b"").decode())


def CreateDecoder(type: str, arch: str, format: str, code: Code, nb_frames: int = 65536, device: int = 0,
                  oms_offset_fixed: int = 1, nms_factor_fixed: int = 29, vMin: int = -127, vMax: int = 127, mMin: int = -31, mMax: int = 31,
                  **kw) -> CGPUDecoder:
    """Factory with the reference's argument meaning (ref: code/x86/CDecoder/DecoderLibrary.h:44-134).

    arch 'sse' -> X86_SSE semantics, 'avx' -> UNIFORM, 'arm' -> ARM_SCALAR, 'gpu' -> GPU_FIXED.  Unknown combinations raise
    (the reference prints and exits)."""
    sem = {"sse": "X86_SSE", "avx": "UNIFORM", "arm": "ARM_SCALAR", "x86": "ARM_SCALAR", "gpu": "GPU_FIXED"}.get(arch)
    if sem is None or format != "fixed" or type not in ALGO:
        raise LdpcError(ERR_UNSUPPORTED, f"(EE) Requested LDPC decoder does not exist ({arch}:{type})")
    if -vMin != vMax and sem != "GPU_FIXED":
        raise LdpcError(ERR_INVALID, "asymmetric variable range")
    return CGPUDecoder(code, nb_frames, device, algo=type, semantics=sem, offset=oms_offset_fixed, factor_q5=nms_factor_fixed,
                       sat_var=vMax, sat_msg=mMax, **kw)


def sigma_for(ebn0_db: float, rate: float) -> float:
    """AWGN sigma of the reference's channel (ref: code/x86/CChanel/CChanelAWGN_MKL.cpp:97-110)."""
    return float(np.sqrt(10.0 ** (-(ebn0_db + 10.0 * np.log10(rate)) / 10.0) / 2.0))
