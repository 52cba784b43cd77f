"""CPU suite, part 1: the oracle against the reference's golden vectors (tests/golden, minted by tools/gen_golden.py from the
reference's own decoders) and — where oracle/_ref was built in this container — against the reference live."""
import hashlib

import numpy as np
import pytest

from _helpers import (Code, default_params, oracle_decode, oracle_decode_float, oracle_quantize, oracle_pack, awgn_llr, stress_llr,
                      ref_x86, ref_x86_decode, ref_arm, ref_arm_decode, ROOT)

GOLD = ROOT / "tests" / "golden"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def unpack(bits, n):
    return np.unpackbits(bits, axis=1, bitorder="little")[:, :n]


def test_k123_x86sse_576(code576):
    g = np.load(GOLD / "k123_576x288_x86sse.npz")
    llr = g["llr"]
    for algo, param in [("OMS", 1), ("OMS", 2), ("NMS", 29), ("NMS", 24)]:
        prm = default_params(algo=algo, semantics="X86_SSE", offset=param, factor_q5=param)
        for iters in [1, 2, 5, 10]:
            o = oracle_decode(code576, prm, llr, iters)
            key = f"{algo}_{param}_{iters}"
            assert np.array_equal(o["hard"], unpack(g[key + "_hard"], code576.n)), key
            if iters in (1, 10):
                assert np.array_equal(o["post"], g[key + "_post"]), key
                assert np.array_equal(o["msgs"], g[key + "_msgs"]), key
            else:
                assert [sha(o["post"]), sha(o["msgs"])] == list(g[key + "_sha"]), key


@pytest.mark.parametrize("name", ["1944x972", "2048x384", "2304x1152", "4000x2000", "64800x32400", "64800x7200", "64800x6480"])
def test_k4_other_codes(built, name):
    c = Code.load(name)
    g = np.load(GOLD / f"k4_{name}_x86sse.npz")
    for key in [k[:-5] for k in g.files if k.endswith("_hard")]:
        algo, param, iters = key.split("_")
        prm = default_params(algo=algo, semantics="X86_SSE", offset=int(param), factor_q5=int(param))
        o = oracle_decode(c, prm, g["llr"], int(iters))
        assert np.array_equal(o["hard"], unpack(g[key + "_hard"], c.n)), key
        assert [sha(o["post"]), sha(o["msgs"])] == list(g[key + "_sha"]), key


def test_k5_arm_scalar_early_termination(code576):
    g = np.load(GOLD / "k5_576x288_armscalar_et.npz")
    for (off, sv, sm) in [(1, 127, 31), (1, 63, 15)]:
        for imax in [10, 30]:
            key = f"ET_{off}_{sv}_{sm}_{imax}"
            prm = default_params(algo="OMS", semantics="ARM_SCALAR", offset=off, sat_var=sv, sat_msg=sm, early_term=1)
            o = oracle_decode(code576, prm, g["llr"], imax)
            assert np.array_equal(o["iters"], g[key + "_iters"]), key
            assert np.array_equal(o["hard"], unpack(g[key + "_hard"], code576.n)), key
            assert np.array_equal(o["post"], g[key + "_post"]) and np.array_equal(o["msgs"], g[key + "_msgs"]), key
    assert g["ET_1_127_31_30_iters"].min() < 30   # the fixture does exercise early stops


def test_k7_int16_wide_rails_against_the_arm_scalar_reference(code576):
    """int16 storage is pinned by the ARM tree's scalar decoder (short arrays, run-time rails) at rails beyond int8: fixture K7."""
    g = np.load(GOLD / "k7_576x288_armscalar_wide.npz")
    llr = g["llr"].astype(np.int16)
    keys = [k[:-5] for k in g.files if k.endswith("_hard")]
    assert len(keys) == 6
    for key in keys:
        _, off, sv, sm, imax, early = key.split("_")
        prm = default_params(algo="OMS", semantics="ARM_SCALAR", dtype=1, offset=int(off), sat_var=int(sv), sat_msg=int(sm), early_term=int(early))
        o = oracle_decode(code576, prm, llr, int(imax))
        assert np.array_equal(o["hard"], unpack(g[key + "_hard"], code576.n)), key
        assert np.array_equal(o["post"], g[key + "_post"]) and np.array_equal(o["msgs"], g[key + "_msgs"]) and np.array_equal(o["iters"], g[key + "_iters"]), key
    assert int(np.abs(g["W_3_32767_8191_30_0_post"].astype(np.int32)).max()) == 32767 and int(np.abs(g["W_3_32767_8191_30_0_msgs"].astype(np.int32)).max()) > 8000
    assert g["W_1_2047_511_20_1_iters"].min() < 20


@pytest.mark.parametrize("name", ["155x93", "2640x1320", "1920x960"])
def test_k8_arm_tree_tables(built, name):
    """The tables only the ARM tree carries, through its scalar decoder (fixture K8): oracle ARM_SCALAR == reference."""
    c = Code.load(name)
    g = np.load(GOLD / f"k8_{name}_armscalar.npz")
    for key in [k[:-5] for k in g.files if k.endswith("_hard")]:
        _, off, sv, sm, imax, early = key.split("_")
        prm = default_params(algo="OMS", semantics="ARM_SCALAR", offset=int(off), sat_var=int(sv), sat_msg=int(sm), early_term=int(early))
        o = oracle_decode(c, prm, g["llr"], int(imax))
        assert np.array_equal(o["hard"], unpack(g[key + "_hard"], c.n)) and np.array_equal(o["iters"], g[key + "_iters"]), key
        assert np.array_equal(o["post"], g[key + "_post"]) and np.array_equal(o["msgs"], g[key + "_msgs"]), key


def test_live_reference_arm_wide_rails(code576):
    L = ref_arm("576x288")
    if L is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    llr = np.clip(5 * awgn_llr(code576, 40, 2.5, 17).astype(np.int16), -128, 127).astype(np.int8)
    for off, sv, sm, early in ((2, 4095, 1023, True), (1, 32767, 32767, False)):
        prm = default_params(algo="OMS", semantics="ARM_SCALAR", dtype=1, offset=off, sat_var=sv, sat_msg=sm, early_term=int(early))
        r = ref_arm_decode(L, code576, off, sv, sm, early, llr, 25)
        o = oracle_decode(code576, prm, llr.astype(np.int16), 25)
        assert np.array_equal(r["iters"], o["iters"]) and np.array_equal(r["hard"], o["hard"])
        assert np.array_equal(r["post"], o["post"]) and np.array_equal(r["msgs"], o["msgs"])


def test_live_reference_x86(code576):
    L = ref_x86("576x288")
    if L is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    llr = np.concatenate([awgn_llr(code576, 64, 1.5, 11), stress_llr(code576, 64, 12, full_range=True)])
    for algo, param in [("OMS", 1), ("NMS", 29)]:
        prm = default_params(algo=algo, semantics="X86_SSE", offset=param, factor_q5=param)
        r = ref_x86_decode(L, algo, param, llr, 7)
        o = oracle_decode(code576, prm, llr, 7)
        for k in ("hard", "post", "msgs"):
            assert np.array_equal(r[k], o[k]), (algo, k)


def test_live_reference_arm(code576):
    L = ref_arm("576x288")
    if L is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    llr = awgn_llr(code576, 48, 2.5, 13)
    prm = default_params(algo="OMS", semantics="ARM_SCALAR", early_term=1)
    r = ref_arm_decode(L, code576, 1, 127, 31, True, llr, 20)
    o = oracle_decode(code576, prm, llr, 20)
    assert np.array_equal(r["iters"], o["iters"]) and np.array_equal(r["hard"], o["hard"])
    assert np.array_equal(r["post"], o["post"]) and np.array_equal(r["msgs"], o["msgs"])


def test_semantics_modes_differ_where_the_survey_says(code576):
    """the reference's own variants disagree on posteriors (-127 vs -128 rail) but rarely on hard decisions."""
    llr = awgn_llr(code576, 256, 3.0, 21)
    res = {s: oracle_decode(code576, default_params(algo="OMS", semantics=s), llr, 10) for s in ("X86_SSE", "UNIFORM", "GPU_FIXED")}
    assert (res["X86_SSE"]["post"] != res["GPU_FIXED"]["post"]).sum() > 1000
    assert (res["X86_SSE"]["hard"] != res["GPU_FIXED"]["hard"]).any(axis=1).mean() < 0.05
    assert (res["X86_SSE"]["hard"] != res["UNIFORM"]["hard"]).any(axis=1).mean() < 0.05


def test_quantiser_and_packing():
    y = np.array([-10.0, -3.876, -0.124, -0.0, 0.1249, 0.125, 3.874, 3.875, 3.99, 100.0], np.float32)
    assert oracle_quantize(y).tolist() == [-31, -31, 0, 0, 0, 1, 30, 31, 31, 31]
    hard = (np.arange(3 * 19).reshape(3, 19) % 3 == 0).astype(np.uint8)
    p = oracle_pack(hard, 19)
    assert p.shape == (3, 3) and np.array_equal(np.unpackbits(p, axis=1, bitorder="little")[:, :19], hard)


def test_all_zero_codeword_decodes_at_high_snr(code576):
    llr = awgn_llr(code576, 64, 5.0, 31)
    for sem, algo in [("X86_SSE", "OMS"), ("X86_SSE", "NMS"), ("GPU_FIXED", "MS"), ("GPU_FIXED", "2NMS"), ("ARM_SCALAR", "OMS")]:
        o = oracle_decode(code576, default_params(algo=algo, semantics=sem), llr, 10)
        assert not o["hard"].any(), (sem, algo)


def test_own_definitions_flooding_int16_float_behave(code576):
    """No reference decoder floods, stores int16 or computes in float: these oracle modes are this project's own
    definitions (parity unpinned).  What can be checked on the CPU: they decode, they agree with the pinned int8 layered decoder
    where they must, and early termination only ever shortens."""
    llr = awgn_llr(code576, 200, 3.0, 61)
    lay = oracle_decode(code576, default_params(algo="OMS", semantics="UNIFORM"), llr, 10)
    # int16 storage with int8 rails and int8 inputs is the int8 decoder
    w = oracle_decode(code576, default_params(algo="OMS", semantics="UNIFORM", dtype=1, sat_var=127), llr.astype(np.int16), 10)
    assert np.array_equal(w["post"], lay["post"].astype(np.int16)) and np.array_equal(w["hard"], lay["hard"])
    # flooding needs more iterations than layered but converges to the same codeword at 3 dB
    flo = oracle_decode(code576, default_params(algo="OMS", semantics="UNIFORM", schedule=1), llr, 30)
    assert (flo["hard"].any(axis=1) == lay["hard"].any(axis=1)).mean() > 0.97
    fl10 = oracle_decode(code576, default_params(algo="OMS", semantics="UNIFORM", schedule=1), llr, 10)
    assert fl10["hard"].any(axis=1).sum() >= flo["hard"].any(axis=1).sum()
    et = oracle_decode(code576, default_params(algo="OMS", semantics="UNIFORM", schedule=1, early_term=1), llr, 30)
    assert et["iters"].max() <= 30 and et["iters"].min() < 30
    conv = et["iters"] < 30
    assert not et["hard"][conv].any()                      # frames that stopped early satisfy every check: the all-zero codeword here
    # float: unquantised values of the same noise decode at least as well as the 6-bit quantised ones
    rng = np.random.Generator(np.random.Philox(61))
    y = (-1.0 + 0.7079 * rng.standard_normal((200, code576.n))).astype(np.float32)
    for sched in (0, 1):
        f = oracle_decode_float(code576, default_params(algo="NMS", dtype=2, schedule=sched, early_term=1), y, 40)
        assert f["hard"].any(axis=1).mean() < 0.05 and f["iters"].min() < 40
        assert not f["hard"][f["iters"] < 40].any()
    fo = oracle_decode_float(code576, default_params(algo="OMS", dtype=2, schedule=0), y, 10)
    assert fo["hard"].any(axis=1).mean() < 0.05
