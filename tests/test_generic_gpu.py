"""GPU suite, generic engine (fp32 arithmetic; kernel 3 = one frame per thread with the state in HBM, kernel 5 = (row, frame)
tasks with the state in shared memory): int16 storage, float min-sum, the flooding schedule — and int8 layered in all four
reference semantics, which pins this engine to the reference-checked oracle.
Integer modes are bit-exact (hard decisions, posteriors, messages, iteration counts).  The float mode issues the oracle's
operations in the oracle's order, so it is compared bit-for-bit too; the stated tolerance of the float path is therefore 0 ulp
against oracle/ldpc_oracle.c (which is this project's own definition: no float decoder exists in the reference — parity unpinned)."""
import numpy as np
import pytest

import ldpcgputegra_b200 as pkg
from _helpers import Code, oracle_decode, oracle_decode_float, awgn_llr, stress_llr, ROOT

pytestmark = pytest.mark.gpu

COMBOS = [("X86_SSE", "OMS"), ("X86_SSE", "NMS"), ("UNIFORM", "OMS"), ("UNIFORM", "NMS"), ("ARM_SCALAR", "OMS"),
          ("GPU_FIXED", "MS"), ("GPU_FIXED", "OMS"), ("GPU_FIXED", "NMS"), ("GPU_FIXED", "2NMS")]


ENGINES = [6, 5, 3]


def gpu_decode(code, llr, iters, kernel=0, **kw):
    dec = pkg.CGPUDecoder(code, nb_frames=max(llr.shape[0], 1), device=0, kernel=kernel, **kw)
    dec.set_debug(True)
    hard, it = dec.decode(llr, iters, want_iters=True)
    post, msgs = dec.debug_state(llr.shape[0])
    k, prm = dec.info(pkg.INFO_KERNEL), dec.params
    dec.close()
    assert k == (kernel or k) and k in (3, 5, 6)
    return dict(hard=hard, post=post, msgs=msgs, iters=it, prm=prm)


def assert_same(g, o, what, iters=True):
    for key in ("hard", "post", "msgs") + (("iters",) if iters else ()):
        assert g[key].dtype == o[key].dtype and np.array_equal(g[key], o[key]), f"{what}: {key} differs in {(g[key] != o[key]).sum()} entries"


def float_llr(code, frames, ebn0, seed):
    """unquantised channel values y = -1 + sigma*n (norm_channel = false, ref: code/x86/main_p.cpp:124)"""
    rate = (code.n - code.n_checks) / code.n
    sigma = np.sqrt(10.0 ** (-(ebn0 + 10.0 * np.log10(rate)) / 10.0) / 2.0)
    rng = np.random.Generator(np.random.Philox(seed))
    return (-1.0 + sigma * rng.standard_normal((frames, code.n))).astype(np.float32)


@pytest.mark.parametrize("kernel", ENGINES)
@pytest.mark.parametrize("sem,algo", COMBOS)
def test_int8_layered_pins_generic_engine(code576, sem, algo, kernel):
    llr = np.concatenate([awgn_llr(code576, 200, 2.0, 141), stress_llr(code576, 150, 143), stress_llr(code576, 101, 144, full_range=True)])
    for iters in (1, 10):
        g = gpu_decode(code576, llr, iters, algo=algo, semantics=sem, kernel=kernel)
        assert_same(g, oracle_decode(code576, g["prm"], llr, iters), f"{sem}/{algo}/I{iters}/k{kernel}")
    if kernel in (5, 6):      # early termination of the fixed-point layered schedule through the on-chip engines (extrinsic-sign criterion)
        g = gpu_decode(code576, llr, 30, algo=algo, semantics=sem, kernel=kernel, early_term=1)
        assert_same(g, oracle_decode(code576, g["prm"], llr, 30), f"{sem}/{algo}/ET/k{kernel}")


@pytest.mark.parametrize("kernel", ENGINES)
@pytest.mark.parametrize("sem,algo", COMBOS)
def test_int8_flooding(code576, sem, algo, kernel):
    llr = np.concatenate([awgn_llr(code576, 200, 2.0, 151), stress_llr(code576, 133, 153, full_range=(sem == "GPU_FIXED"))])
    for iters, et in ((1, 0), (10, 0), (30, 1)):
        g = gpu_decode(code576, llr, iters, algo=algo, semantics=sem, schedule="FLOODING", early_term=et, kernel=kernel)
        assert_same(g, oracle_decode(code576, g["prm"], llr, iters), f"flooding {sem}/{algo}/I{iters}/et{et}/k{kernel}")
    assert g["iters"].min() < 30


@pytest.mark.parametrize("kernel", ENGINES)
def test_int16_pinned_against_the_reference_scalar_decoder(code576, kernel):
    """K7: the int16 storage path against the only reference code with wider-than-int8 state — the ARM tree's scalar decoder at
    rails far beyond int8 (fixture minted from it by tools/gen_golden.py: setVarRange(+-2047)/setMsgRange(+-511), +-32767/+-8191,
    +-300/+-300; posteriors reach the rails, messages 8188): hard decisions, posteriors, messages and iteration counts, both generic
    engines (kernel 3: HBM state, kernel 5: on-chip state)."""
    g7 = np.load(ROOT / "tests" / "golden" / "k7_576x288_armscalar_wide.npz")
    llr = g7["llr"].astype(np.int16)
    seen_wide = False
    for key in [k[:-5] for k in g7.files if k.endswith("_hard")]:
        _, off, sv, sm, imax, early = key.split("_")
        g = gpu_decode(code576, llr, int(imax), dtype="I16", semantics="ARM_SCALAR", algo="OMS", offset=int(off), sat_var=int(sv), sat_msg=int(sm),
                       early_term=int(early), kernel=kernel)
        assert np.array_equal(np.packbits(g["hard"], axis=1, bitorder="little"), g7[key + "_hard"]), key
        assert np.array_equal(g["post"], g7[key + "_post"]) and np.array_equal(g["msgs"], g7[key + "_msgs"]), key
        assert np.array_equal(g["iters"], g7[key + "_iters"]), key
        seen_wide = seen_wide or int(np.abs(g["post"].astype(np.int32)).max()) > 127
    assert seen_wide


@pytest.mark.parametrize("kernel", ENGINES)
@pytest.mark.parametrize("schedule", ["LAYERED", "FLOODING"])
@pytest.mark.parametrize("kw", [dict(semantics="ARM_SCALAR", algo="OMS", sat_var=32767, sat_msg=8191, offset=16),
                                dict(semantics="ARM_SCALAR", algo="OMS", sat_var=2047, sat_msg=511, offset=8),
                                dict(semantics="UNIFORM", algo="OMS", sat_var=32767, sat_msg=4095, offset=16),
                                dict(semantics="UNIFORM", algo="NMS", sat_var=8191, sat_msg=2047, factor_q5=29)])
def test_int16(code576, schedule, kw, kernel):
    q8 = np.concatenate([awgn_llr(code576, 200, 1.5, 161), stress_llr(code576, 120, 163)]).astype(np.int16)
    llr = (q8 * 16 + (np.arange(q8.size).reshape(q8.shape) % 13 - 6)).astype(np.int16)     # 16x finer grid, not multiples of 16
    llr[5] = np.clip(llr[5].astype(np.int32) * 40, -32768, 32767).astype(np.int16)         # drives the rails
    for iters, et in ((2, 0), (10, 0), (25, 1)):
        g = gpu_decode(code576, llr, iters, dtype="I16", schedule=schedule, early_term=et, kernel=kernel, **kw)
        assert_same(g, oracle_decode(code576, g["prm"], llr, iters), f"int16 {schedule} {kw} I{iters} et{et} k{kernel}")


@pytest.mark.parametrize("kernel", ENGINES)
@pytest.mark.parametrize("schedule", ["FLOODING", "LAYERED"])
@pytest.mark.parametrize("kw", [dict(algo="NMS", factor1=0.75), dict(algo="2NMS", factor1=0.75, factor2=0.875), dict(algo="MS"),
                                dict(algo="OMS", offset=1), dict(algo="NMS", factor1=0.8125)])
def test_float_min_sum(code576, schedule, kw, kernel):
    llr = np.concatenate([float_llr(code576, 150, 2.0, 171), float_llr(code576, 100, 0.5, 172), float_llr(code576, 51, 4.0, 173)])
    for iters, et in ((1, 0), (10, 0), (40, 1)):
        g = gpu_decode(code576, llr, iters, dtype="F32", schedule=schedule, early_term=et, kernel=kernel, **kw)
        o = oracle_decode_float(code576, g["prm"], llr, iters)
        # tolerance: 0 (same operations in the same order); hard decisions identical
        assert_same(g, o, f"float {schedule} {kw} I{iters} et{et}")
    assert g["iters"].min() < 40


@pytest.mark.parametrize("name", ["1944x972", "2048x384", "1200x600", "200x100"])
def test_generic_engine_other_codes(built, name):
    """degree 32 rows (2048x384: the run-time-degree path), column degree up to 15 (1200x600), N % 32 != 0 (200x100)"""
    code = Code.load(name)
    llr8 = awgn_llr(code, 70, 2.5, 181)
    y = float_llr(code, 70, 2.5, 182)
    for kernel in ENGINES:
        try:
            g = gpu_decode(code, llr8, 5, algo="OMS", semantics="X86_SSE", kernel=kernel)
        except pkg.LdpcError as e:        # 1200x600 (464 levels) and 2048x384 (degree 32): the layered plan of 32-row steps exceeds the warp-per-frame engine's 16-bit tables
            assert kernel == 6 and name in ("1200x600", "2048x384") and e.status == pkg.ERR_UNSUPPORTED
            g = None
        if g is not None:
            assert_same(g, oracle_decode(code, g["prm"], llr8, 5), f"{name} int8 layered k{kernel}")
        g = gpu_decode(code, llr8, 5, algo="NMS", semantics="UNIFORM", schedule="FLOODING", kernel=kernel)
        assert_same(g, oracle_decode(code, g["prm"], llr8, 5), f"{name} int8 flooding k{kernel}")
        g = gpu_decode(code, y, 6, dtype="F32", algo="NMS", schedule="FLOODING", early_term=1, kernel=kernel)
        assert_same(g, oracle_decode_float(code, g["prm"], y, 6), f"{name} float flooding k{kernel}")


def test_generic_ragged_packed_and_device_channel(code576):
    y = float_llr(code576, 45, 3.0, 191)
    for frames, kernel in ((0, 6), (1, 6), (33, 6), (45, 6), (0, 5), (1, 5), (31, 5), (45, 5), (45, 3), (1, 3)):
        dec = pkg.CGPUDecoder(code576, nb_frames=64, device=0, dtype="F32", algo="NMS", schedule="FLOODING", out_format=1, kernel=kernel)
        packed = dec.decode(y[:frames], 8)
        ref = oracle_decode_float(code576, dec.params, y[:frames], 8)["hard"]
        assert np.array_equal(np.unpackbits(packed, axis=1, bitorder="little")[:, :code576.n] if frames else packed.reshape(0, code576.n), ref)
        dec.close()
    # the on-device channel writes the handle's dtype; same (seed, frame) -> same noise whatever the dtype
    d8 = pkg.CGPUDecoder(code576, nb_frames=256, device=0)
    df = pkg.CGPUDecoder(code576, nb_frames=256, device=0, dtype="F32", algo="NMS", schedule="FLOODING")
    sigma = pkg.sigma_for(2.0, 0.5)
    q, yf = d8.awgn(256, sigma, seed=5), df.awgn(256, sigma, seed=5)
    assert yf.dtype == np.float32 and np.array_equal(np.clip(np.trunc(8.0 * yf), -31, 31).astype(np.int8), q)
    hard = df.decode(yf, 10)
    assert np.array_equal(hard, oracle_decode_float(code576, df.params, yf, 10)["hard"])
    d8.close(); df.close()


def test_unsupported_combinations_fail_loudly(code576):
    for kw in [dict(dtype="I16", semantics="X86_SSE"), dict(dtype="I16", semantics="GPU_FIXED"), dict(dtype="F32", kernel=2),
               dict(schedule="FLOODING", kernel=1), dict(dtype="I16", semantics="UNIFORM", sat_var=40000)]:
        with pytest.raises(pkg.LdpcError):
            pkg.CGPUDecoder(code576, nb_frames=64, device=0, **kw)


def test_float_flooding_full_batch_properties(code576):
    """BASELINE configs[2] at a full 65 536-frame batch (the on-chip generic engine picked by the library): strided sample against
    the CPU restatement (decisions and iteration counts), frames independent (permutation equivariance), early termination only
    ever stops frames whose syndrome is zero, FER where the fixed-point decoder puts it."""
    F = 65536
    dec = pkg.CGPUDecoder(code576, nb_frames=F, device=0, dtype="F32", algo="NMS", factor1=0.75, schedule="FLOODING", early_term=1)
    assert dec.info(pkg.INFO_KERNEL) == 6
    y = dec.awgn(F, pkg.sigma_for(2.5, 0.5), seed=2025)
    hard, it = dec.decode(y, 40, want_iters=True)
    sample = np.arange(0, F, 131)
    o = oracle_decode_float(code576, dec.params, y[sample], 40)
    assert np.array_equal(hard[sample], o["hard"]) and np.array_equal(it[sample], o["iters"])
    perm = np.random.default_rng(2).permutation(F)
    hard_p, it_p = dec.decode(y[perm], 40, want_iters=True)
    assert np.array_equal(hard_p, hard[perm]) and np.array_equal(it_p, it[perm])
    stopped = it < 40
    e = r = 0; synd = np.zeros(F, bool)                       # H c for every frame
    for d, cnt in zip(code576.deg, code576.rows):
        idx = code576.pos[e:e + d * cnt].reshape(cnt, d).astype(np.int64)
        synd |= ((hard[:, idx].sum(axis=2) & 1) != 0).any(axis=1)
        e += d * cnt
    assert stopped.mean() > 0.95 and not synd[stopped].any()
    assert hard[:, :code576.k_info].any(axis=1).mean() < 0.02
    dec.close()
