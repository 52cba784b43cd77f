"""GPU suite: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs — bit-exact on hard
decisions, posteriors, messages and iteration counts — plus the committed golden fixtures and size-independent properties
at BASELINE.json's full batch size."""
import ctypes as C

import numpy as np
import pytest

import ldpcgputegra_b200 as pkg
from _helpers import (Code, default_params, oracle_decode, oracle_decode_mt, oracle_quantize, oracle_pack, awgn_llr, stress_llr, ROOT,
                      ref_gpu, ref_gpu_decode, ref_x86, ALGO)

pytestmark = pytest.mark.gpu
GOLD = ROOT / "tests" / "golden"

COMBOS = [("X86_SSE", "OMS"), ("X86_SSE", "NMS"), ("UNIFORM", "OMS"), ("UNIFORM", "NMS"), ("ARM_SCALAR", "OMS"),
          ("GPU_FIXED", "MS"), ("GPU_FIXED", "OMS"), ("GPU_FIXED", "NMS"), ("GPU_FIXED", "2NMS")]


def gpu_decode(code, llr, iters, want_iters=False, **kw):
    dec = pkg.CGPUDecoder(code, nb_frames=max(llr.shape[0], 1), device=0, **kw)
    dec.set_debug(True)
    r = dec.decode(llr, iters, want_iters=want_iters)
    hard, it = r if want_iters else (r, None)
    post, msgs = dec.debug_state(llr.shape[0])
    k = dec.info(pkg.INFO_KERNEL)
    prm = dec.params
    dec.close()
    return dict(hard=hard, post=post, msgs=msgs, iters=it, kernel=k, prm=prm)


def assert_same(g, o, what, msgs=True):
    assert np.array_equal(g["hard"], o["hard"]), f"{what}: hard decisions differ in {(g['hard'] != o['hard']).sum()} bits"
    assert np.array_equal(g["post"], o["post"]), f"{what}: posteriors differ in {(g['post'] != o['post']).sum()} entries"
    if msgs:
        assert np.array_equal(g["msgs"], o["msgs"]), f"{what}: messages differ in {(g['msgs'] != o['msgs']).sum()} entries"


@pytest.mark.parametrize("kernel", [2, 1])
@pytest.mark.parametrize("sem,algo", COMBOS)
def test_bit_exact_all_semantics_576(code576, kernel, sem, algo):
    full = sem == "GPU_FIXED"          # the GPU kernels take any int8 input; x86 inputs stay inside the quantiser range too
    llr = np.concatenate([awgn_llr(code576, 300, 2.0, 41), awgn_llr(code576, 100, 0.5, 42), stress_llr(code576, 200, 43),
                          stress_llr(code576, 101, 44, full_range=True)])
    for iters in (1, 2, 10):
        g = gpu_decode(code576, llr, iters, algo=algo, semantics=sem, kernel=kernel)
        assert g["kernel"] == kernel
        o = oracle_decode(code576, g["prm"], llr, iters)
        assert_same(g, o, f"{sem}/{algo}/k{kernel}/I{iters}")
    assert full or True


@pytest.mark.parametrize("kernel", [2, 1])
def test_parameter_setters(code576, kernel):
    llr = np.concatenate([awgn_llr(code576, 128, 1.5, 51), stress_llr(code576, 128, 52)])
    for kw in [dict(algo="OMS", semantics="X86_SSE", offset=2), dict(algo="OMS", semantics="X86_SSE", offset=0, sat_msg=15),
               dict(algo="NMS", semantics="X86_SSE", factor_q5=24), dict(algo="NMS", semantics="UNIFORM", factor_q5=31),
               dict(algo="OMS", semantics="ARM_SCALAR", sat_var=63, sat_msg=15), dict(algo="OMS", semantics="ARM_SCALAR", sat_var=100, offset=3)]:
        g = gpu_decode(code576, llr, 6, kernel=kernel, **kw)
        o = oracle_decode(code576, g["prm"], llr, 6)
        assert_same(g, o, str(kw))


@pytest.mark.parametrize("kernel", [2, 1])
def test_early_termination_iteration_counts(code576, kernel):
    llr = np.concatenate([awgn_llr(code576, 150, 1.0, 61), awgn_llr(code576, 150, 2.5, 62), awgn_llr(code576, 33, 4.0, 63)])
    for sem, algo in [("ARM_SCALAR", "OMS"), ("X86_SSE", "OMS"), ("GPU_FIXED", "2NMS")]:
        for imax in (10, 30):
            g = gpu_decode(code576, llr, imax, want_iters=True, algo=algo, semantics=sem, early_term=1, kernel=kernel)
            o = oracle_decode(code576, g["prm"], llr, imax)
            assert np.array_equal(g["iters"], o["iters"]), (sem, imax, np.flatnonzero(g["iters"] != o["iters"])[:8])
            assert_same(g, o, f"ET {sem} I{imax}")
            assert o["iters"].min() < imax and o["iters"].max() == imax


def test_golden_fixtures_through_the_abi(code576):
    g = np.load(GOLD / "k123_576x288_x86sse.npz")
    for algo, param in [("OMS", 1), ("OMS", 2), ("NMS", 29), ("NMS", 24)]:
        for iters in (1, 10):
            key = f"{algo}_{param}_{iters}"
            for kernel in (2, 1):
                r = gpu_decode(code576, g["llr"], iters, algo=algo, semantics="X86_SSE", offset=param, factor_q5=param, kernel=kernel)
                assert np.array_equal(np.packbits(r["hard"], axis=1, bitorder="little"), g[key + "_hard"]), key
                assert np.array_equal(r["post"], g[key + "_post"]) and np.array_equal(r["msgs"], g[key + "_msgs"]), key
    k5 = np.load(GOLD / "k5_576x288_armscalar_et.npz")
    for kernel in (2, 1):
        r = gpu_decode(code576, k5["llr"], 30, want_iters=True, algo="OMS", semantics="ARM_SCALAR", early_term=1, kernel=kernel)
        assert np.array_equal(r["iters"], k5["ET_1_127_31_30_iters"])
        assert np.array_equal(r["post"], k5["ET_1_127_31_30_post"]) and np.array_equal(r["msgs"], k5["ET_1_127_31_30_msgs"])


@pytest.mark.parametrize("name", ["1944x972", "2048x384", "2304x1152", "4000x2000", "1200x600", "200x100", "816x408"])
def test_other_codes(built, name):
    c = Code.load(name)
    llr = np.concatenate([awgn_llr(c, 40, 2.0, 71), stress_llr(c, 23, 72)])
    for sem, algo in [("X86_SSE", "OMS"), ("GPU_FIXED", "OMS")] + ([("X86_SSE", "NMS")] if len(c.deg) <= 2 else []):
        g = gpu_decode(c, llr, 5, algo=algo, semantics=sem)
        o = oracle_decode(c, g["prm"], llr, 5)
        assert_same(g, o, f"{name} {sem}/{algo} kernel {g['kernel']}")
        g1 = gpu_decode(c, llr, 5, algo=algo, semantics=sem, kernel=1)
        assert_same(g1, o, f"{name} {sem}/{algo} kernel 1")
        if c.n_checks >= 128 and max(c.deg) <= 10 and min(c.deg) >= 3:     # the bulk-copy-staged frame-parallel kernel
            for nc in (128, 256):
                g4 = gpu_decode(c, llr, 5, algo=algo, semantics=sem, kernel=4, fs_nc=nc)
                assert g4["kernel"] == 4
                assert_same(g4, o, f"{name} {sem}/{algo} kernel 4 NC{nc}")
            g4e = gpu_decode(c, llr, 12, algo=algo, semantics=sem, kernel=4, early_term=1, want_iters=True)
            oe = oracle_decode(c, g4e["prm"], llr, 12)
            assert_same(g4e, oe, f"{name} {sem}/{algo} kernel 4 early termination")
            assert np.array_equal(g4e["iters"], oe["iters"])
        elif c.n_checks >= 128 and max(c.deg) <= 32 and min(c.deg) >= 3:   # rows wider than 10: the staged kernel's two-pass row body, on request
            for kw in (dict(), dict(fs_tma=1, fs_g4=1), dict(fs_stages=2)):
                g4 = gpu_decode(c, llr, 5, algo=algo, semantics=sem, kernel=4, **kw)
                assert g4["kernel"] == 4
                assert_same(g4, o, f"{name} {sem}/{algo} kernel 4 (wide rows) {kw}")
            g4e = gpu_decode(c, llr, 12, algo=algo, semantics=sem, kernel=4, early_term=1, want_iters=True)
            oe = oracle_decode(c, g4e["prm"], llr, 12)
            assert_same(g4e, oe, f"{name} {sem}/{algo} kernel 4 (wide rows) early termination")
            assert np.array_equal(g4e["iters"], oe["iters"])
    gold = GOLD / f"k4_{name}_x86sse.npz"
    if gold.exists():
        gg = np.load(gold)
        r = gpu_decode(c, gg["llr"], 10, algo="OMS", semantics="X86_SSE", offset=1)
        assert np.array_equal(np.packbits(r["hard"], axis=1, bitorder="little"), gg["OMS_1_10_hard"])


@pytest.mark.parametrize("kernel", [0, 1, 4, 4256, 4001, 4002, 4003, 4004, 4005])
def test_dvbs2_long_code_frame_parallel(built, kernel):
    """DVB-S2 64800x32400 (a 32 399-deep chain in reference order) against the reference's own x86 decoder (golden fixture):
    plain frame-parallel kernel (1), bulk-copy-staged kernel (4), and whatever the library picks (0 -> 4)."""
    c = Code.load("64800x32400")
    gg = np.load(GOLD / "k4_64800x32400_x86sse.npz")
    # 4256: kernel 4 with 256-consumer CTAs; 4001: kernel 4 with one-dimensional bulk copies only (no tensor map, no gather4)
    # 4002 / 4003: kernel 4 on compressed messages (four words per row instead of one per edge), tensor-map / one-dimensional copies
    kw = dict(kernel=4, fs_nc=kernel - 4000) if kernel > 4100 else dict(kernel=4, fs_tma=1, fs_g4=1) if kernel == 4001 else \
         dict(kernel=4, fs_cmp=2) if kernel == 4002 else dict(kernel=4, fs_cmp=2, fs_tma=1, fs_g4=1, fs_nc=256) if kernel == 4003 else \
         dict(kernel=4, fs_pipe2=1) if kernel == 4004 else dict(kernel=4, fs_nostair=1) if kernel == 4005 else dict(kernel=kernel)     # 4004: staircase rows one by one; 4005: no staircase runs
    r = gpu_decode(c, gg["llr"], 10, algo="OMS", semantics="X86_SSE", **kw)
    assert r["kernel"] == (kw["kernel"] or 4)
    assert np.array_equal(np.packbits(r["hard"], axis=1, bitorder="little"), gg["OMS_1_10_hard"])
    import hashlib
    assert [hashlib.sha256(r["post"].tobytes()).hexdigest(), hashlib.sha256(r["msgs"].tobytes()).hexdigest()] == list(gg["OMS_1_10_sha"])


@pytest.mark.parametrize("name", ["64800x7200", "64800x6480"])
def test_dvbs2_high_rate_tables(built, name):
    """The reference tree's other two DVB-S2 tables (rate 8/9: rows of degree 27/26; rate 9/10: 30/29 — code/x86/Constantes/
    64800x7200.dvb-s2, 64800x6480.dvb-s2) against fixtures minted from the reference's own x86 decoder: OMS and NMS, hard decisions,
    SHA-256 of posteriors and messages, through whatever kernel the library picks and through the plain frame-parallel kernel."""
    import hashlib
    c = Code.load(name)
    gg = np.load(GOLD / f"k4_{name}_x86sse.npz")
    for algo, param in (("OMS", 1), ("NMS", 29)):
        key = f"{algo}_{param}_10"
        for kernel in (0, 1, 4):          # 4: the staged kernel's two-pass row body for rows wider than 10
            r = gpu_decode(c, gg["llr"], 10, algo=algo, semantics="X86_SSE", offset=param, factor_q5=param, kernel=kernel)
            assert np.array_equal(np.packbits(r["hard"], axis=1, bitorder="little"), gg[key + "_hard"]), (key, kernel)
            assert [hashlib.sha256(r["post"].tobytes()).hexdigest(), hashlib.sha256(r["msgs"].tobytes()).hexdigest()] == list(gg[key + "_sha"]), (key, kernel, r["kernel"])
    # GPU_FIXED semantics and early termination on the same table against the CPU restatement (a handful of frames: the oracle is scalar)
    llr = gg["llr"][4:10]
    for kw in (dict(algo="OMS", semantics="GPU_FIXED"), dict(algo="OMS", semantics="ARM_SCALAR", early_term=1), dict(algo="2NMS", semantics="GPU_FIXED", kernel=4),
               dict(algo="OMS", semantics="ARM_SCALAR", early_term=1, kernel=4), dict(algo="NMS", semantics="UNIFORM", early_term=1, kernel=4, fs_tma=1, fs_g4=1)):
        g = gpu_decode(c, llr, 12, want_iters=True, **kw)
        o = oracle_decode(c, g["prm"], llr, 12)
        assert_same(g, o, f"{name} {kw}")
        assert np.array_equal(g["iters"], o["iters"])


def test_dvbs2_early_termination_staged_vs_plain_vs_oracle(built):
    """DVB-S2 with the per-frame stop criterion: the staged kernel (a second pass over the ring per iteration), the plain
    frame-parallel kernel and the CPU restatement agree on hard decisions, posteriors, messages and iteration counts."""
    c = Code.load("64800x32400")
    llr = np.concatenate([np.load(GOLD / "k4_64800x32400_x86sse.npz")["llr"][:6], awgn_llr(c, 5, 2.6, 611), awgn_llr(c, 3, 0.5, 612)])
    res = {}
    for kernel in (4, 1):
        res[kernel] = gpu_decode(c, llr, 25, algo="OMS", semantics="ARM_SCALAR", kernel=kernel, early_term=1, want_iters=True)
        assert res[kernel]["kernel"] == kernel
    o = oracle_decode(c, res[4]["prm"], llr, 25)
    for kernel in (4, 1):
        assert_same(res[kernel], o, f"DVB-S2 early termination kernel {kernel}")
        assert np.array_equal(res[kernel]["iters"], o["iters"])
    assert o["iters"].min() < 25 and o["iters"].max() == 25          # some frames stop early, the 0.5 dB ones never do
    auto = pkg.CGPUDecoder(c, nb_frames=16, device=0, early_term=1)
    assert auto.info(pkg.INFO_KERNEL) == 4
    # the fast paths are really the ones that run: 32 398 of the 32 400 rows sit in one register-carried staircase run, and a batch this
    # small is decoded by the paired-row instantiation with 128 consumers per CTA
    assert auto.info(pkg.INFO_FS_STAIR_ROWS) == 32398
    auto.decode(llr[:16], 2)
    assert auto.info(pkg.INFO_FS_VARIANT) == 2 + 256 * 128
    auto.close()
    wide = pkg.CGPUDecoder(Code.load("64800x7200"), nb_frames=16, device=0)
    assert wide.info(pkg.INFO_KERNEL) == 4 and wide.info(pkg.INFO_FS_STAIR_ROWS) == 0
    wide.decode(np.zeros((16, 64800), dtype=np.int8), 1)
    assert wide.info(pkg.INFO_FS_VARIANT) == 1 + 16 + 256 * 128
    wide.close()


@pytest.mark.parametrize("sem,algo", COMBOS)
def test_staged_kernel_all_semantics(code576, sem, algo):
    """kernel 4 on a batch that spans several CTAs and a ragged tail, every (semantics, algorithm) pair, stage ring depths 2..15"""
    llr = np.concatenate([awgn_llr(code576, 700, 2.0, 241), stress_llr(code576, 333, 243, full_range=(sem == "GPU_FIXED"))])
    # tma / g4: 1 = one-dimensional bulk copies only, 2 = message lines through a 2-D tensor map / posterior lines through tile::gather4 (the defaults)
    for iters, stages, nc, tma, g4 in ((1, 0, 128, 2, 2), (10, 0, 128, 2, 2), (3, 2, 128, 2, 2), (3, 15, 128, 2, 2), (10, 0, 256, 2, 2), (3, 2, 256, 2, 2), (2, 9, 256, 2, 2),
                                       (10, 0, 128, 1, 1), (3, 2, 256, 1, 1), (3, 15, 128, 2, 1), (2, 9, 256, 1, 2), (10, 0, 128, 0, 0)):
        g = gpu_decode(code576, llr, iters, algo=algo, semantics=sem, kernel=4, fs_stages=stages, fs_nc=nc, fs_tma=tma, fs_g4=g4, want_iters=True)
        assert g["kernel"] == 4 and (g["iters"] == iters).all()
        assert_same(g, oracle_decode(code576, g["prm"], llr, iters), f"staged {sem}/{algo} I{iters} K{stages} NC{nc} tma{tma} g4{g4}")
    if (sem, algo) in (("X86_SSE", "OMS"), ("GPU_FIXED", "2NMS")):        # several CTAs + a ragged tail
        big = np.concatenate([llr, llr[::-1], llr[:700]])                  # 2766 frames -> 692 words per row -> 6 CTAs of 128 consumers
        for nc in (128, 256):
            g = gpu_decode(code576, big, 4, algo=algo, semantics=sem, kernel=4, fs_nc=nc)
            assert_same(g, oracle_decode(code576, default_params(algo=algo, semantics=sem), big, 4), f"staged {sem}/{algo} 6 CTAs NC{nc}")
    # per-frame early termination in the staged kernel (a second pass over the ring per iteration for the stop criterion): frames of
    # very different quality in one CTA, iteration counts, frozen state, several ring depths and both CTA widths
    mixed = np.concatenate([awgn_llr(code576, 300, 4.0, 245), awgn_llr(code576, 250, 0.0, 246), awgn_llr(code576, 477, 2.0, 247), stress_llr(code576, 40, 248, full_range=(sem == "GPU_FIXED"))])
    for iters, stages, nc, tma in ((10, 0, 128, 0), (20, 2, 128, 0), (7, 0, 256, 0), (2, 9, 256, 0), (3, 15, 128, 0), (1, 0, 128, 0), (10, 0, 128, 1), (7, 3, 256, 1)):
        g = gpu_decode(code576, mixed, iters, algo=algo, semantics=sem, kernel=4, fs_stages=stages, fs_nc=nc, fs_tma=tma, fs_g4=tma, early_term=1, want_iters=True)
        o = oracle_decode(code576, g["prm"], mixed, iters)
        assert g["kernel"] == 4
        assert_same(g, o, f"staged ET {sem}/{algo} I{iters} K{stages} NC{nc}")
        assert np.array_equal(g["iters"], o["iters"]), f"staged ET {sem}/{algo} I{iters}: iteration counts"
    # all frames of the batch converge early: the CTAs leave the loop long before 50 iterations and say so
    good = awgn_llr(code576, 1024, 6.0, 249)
    g = gpu_decode(code576, good, 50, algo=algo, semantics=sem, kernel=4, early_term=1, want_iters=True)
    o = oracle_decode(code576, g["prm"], good, 50)
    assert_same(g, o, f"staged ET {sem}/{algo} clean batch")
    assert np.array_equal(g["iters"], o["iters"]) and g["iters"].max() < 50


@pytest.mark.parametrize("sem,algo", COMBOS)
def test_staged_kernel_compressed_messages(code576, sem, algo):
    """kernel 4 with the check-to-variable messages COMPRESSED (per row and frame: the two magnitudes, which edges get the second one,
    the signs — kernel_fp.cuh: fp_row_math_c): hard decisions, posteriors AND the re-expanded messages equal the CPU restatement's for
    every (semantics, algorithm) pair, ring depths, CTA widths, both producer paths, saturating inputs, a ragged tail, early termination
    with frozen frames and iteration counts."""
    llr = np.concatenate([awgn_llr(code576, 700, 2.0, 241), stress_llr(code576, 333, 243, full_range=(sem == "GPU_FIXED"))])
    for iters, stages, nc, tma, g4 in ((1, 0, 128, 2, 2), (10, 0, 128, 2, 2), (3, 2, 128, 2, 2), (3, 15, 128, 2, 2), (10, 0, 256, 2, 2), (3, 2, 256, 1, 1), (2, 9, 256, 1, 2), (10, 0, 128, 1, 1)):
        g = gpu_decode(code576, llr, iters, algo=algo, semantics=sem, kernel=4, fs_cmp=2, fs_stages=stages, fs_nc=nc, fs_tma=tma, fs_g4=g4, want_iters=True)
        assert g["kernel"] == 4 and (g["iters"] == iters).all()
        assert_same(g, oracle_decode(code576, g["prm"], llr, iters), f"compressed {sem}/{algo} I{iters} K{stages} NC{nc} tma{tma} g4{g4}")
    mixed = np.concatenate([awgn_llr(code576, 300, 4.0, 245), awgn_llr(code576, 250, 0.0, 246), awgn_llr(code576, 477, 2.0, 247), stress_llr(code576, 40, 248, full_range=(sem == "GPU_FIXED"))])
    for iters, stages, nc, tma in ((10, 0, 128, 0), (20, 2, 128, 0), (7, 0, 256, 0), (2, 9, 256, 1), (1, 0, 128, 0)):
        g = gpu_decode(code576, mixed, iters, algo=algo, semantics=sem, kernel=4, fs_cmp=2, fs_stages=stages, fs_nc=nc, fs_tma=tma, fs_g4=tma, early_term=1, want_iters=True)
        o = oracle_decode(code576, g["prm"], mixed, iters)
        assert_same(g, o, f"compressed ET {sem}/{algo} I{iters} K{stages} NC{nc}")
        assert np.array_equal(g["iters"], o["iters"]), f"compressed ET {sem}/{algo} I{iters}: iteration counts"


def _ira_code(d0, r0, d1, r1, k_sys, seed):
    """A synthetic IRA code: two degree classes, every row = (d - 2) systematic columns + the parity staircase p[r-1], p[r] as its last
    two edges (DVB-S2's structure), the systematic columns laid out so that no column repeats within 17 rows — except in a handful of
    rows, which get a second hazard on purpose and cut the staircase runs into pieces of different lengths (one of them shorter than
    the 32 rows a run needs to become a segment of its own)."""
    rng = np.random.default_rng(seed)
    cuts = {100, 101, 300, r0 + 120, r0 + 380, r0 + 420, r0 + 440}
    rows = [(d0, r) for r in range(r0)] + [(d1, r0 + r) for r in range(r1)]
    stream, guard = [], 17 * max(d0, d1)                   # random permutations of the systematic columns, no repeat within `guard` entries
    while len(stream) < (r0 + r1) * max(d0, d1):
        perm = rng.permutation(k_sys).tolist()
        if not set(perm[:guard]) & set(stream[-guard:]):
            stream += perm
    pos, cursor = [], 0
    for d, r in rows:
        take = d - 2 if r else d - 1                         # row 0 has no p[-1]: one more systematic column
        cols = stream[cursor:cursor + take]
        cursor += take
        if r in cuts:
            cols[int(rng.integers(d - 2))] = int(prev[int(rng.integers(len(prev)))])     # a systematic column of the row before: second hazard
        pos += cols + ([k_sys + r - 1, k_sys + r] if r else [k_sys])
        prev = cols
    return Code(k_sys + r0 + r1, r0 + r1, [d0, d1], [r0, r1], np.asarray(pos, dtype=np.uint32))


@pytest.mark.parametrize("d0,d1", [(8, 7), (7, 6)])
def test_staged_kernel_staircase_runs(built, d0, d1):
    """The register-carried staircase runs of the staged kernel (kernel_fs.cuh: fs_row_stair, fs_row_stair2) on synthetic IRA codes whose
    runs are cut at random places, cross a degree-class boundary and wrap around the iteration: single rows, paired rows, neither,
    ring depths that switch the pairing off, compressed messages, every semantics family, early termination with frozen frames —
    hard decisions, posteriors, messages and iteration counts equal the CPU restatement's."""
    c = _ira_code(d0, 520, d1, 731, 1999, 7 * d0 + d1)
    llr = np.concatenate([awgn_llr(c, 260, 1.5, 31), stress_llr(c, 63, 32)])
    variants = [dict(), dict(fs_pipe2=1), dict(fs_pipe2=2), dict(fs_nostair=1), dict(fs_pipe2=2, fs_stages=4), dict(fs_pipe2=2, fs_stages=15), dict(fs_stages=3),
                dict(fs_cmp=2), dict(fs_nc=256), dict(fs_tma=1, fs_g4=1, fs_pipe2=2)]
    for sem, algo in (("X86_SSE", "OMS"), ("X86_SSE", "NMS"), ("GPU_FIXED", "2NMS"), ("ARM_SCALAR", "OMS")):
        prm = None
        for kw in variants:
            g = gpu_decode(c, llr, 6, algo=algo, semantics=sem, kernel=4, **kw)
            assert g["kernel"] == 4
            if prm is None:
                prm, o = g["prm"], oracle_decode(c, g["prm"], llr, 6)
            assert_same(g, o, f"staircase {d0}/{d1} {sem}/{algo} {kw}")
    mixed = np.concatenate([awgn_llr(c, 150, 4.0, 33), awgn_llr(c, 120, 0.5, 34), awgn_llr(c, 90, 2.0, 35)])
    oe = None
    for kw in (dict(), dict(fs_pipe2=1), dict(fs_nostair=1), dict(fs_cmp=2), dict(fs_pipe2=2, fs_stages=5)):
        g = gpu_decode(c, mixed, 15, algo="OMS", semantics="ARM_SCALAR", kernel=4, early_term=1, want_iters=True, **kw)
        if oe is None:
            oe = oracle_decode(c, g["prm"], mixed, 15)
        assert_same(g, oe, f"staircase ET {d0}/{d1} {kw}")
        assert np.array_equal(g["iters"], oe["iters"]), kw
    assert oe["iters"].min() < 15


@pytest.mark.parametrize("name,frames", [("576x288", 151552), ("4000x2000", 75776), ("64800x32400", 9472)])
def test_staged_kernel_stress_against_plain_kernel(built, name, frames):
    """Long-running cross-check of the staged kernel's ordering assumptions (generic-proxy stores -> proxy fence -> mbarrier ->
    bulk / tensor-map copies of the same lines, the forwarded word written into a stage slot the copy engine refills): many
    iterations, shallow and deep stage rings, every SM loaded with several CTAs, both producer paths — every posterior-derived
    decision and every iteration count must equal the plain frame-parallel kernel's, which has no asynchronous copies at all."""
    c = Code.load(name)
    d1 = pkg.CGPUDecoder(c, nb_frames=frames, kernel=1, semantics="ARM_SCALAR", early_term=1)
    llr = d1.awgn(frames, pkg.sigma_for(1.6, 0.5), seed=77)
    h1, it1 = d1.decode(llr, 40, want_iters=True)
    d1.close()
    variants = ((2, 128, 2, 1, 0), (3, 256, 2, 1, 0), (0, 128, 2, 1, 0), (2, 256, 1, 1, 0), (0, 256, 1, 1, 0), (2, 128, 2, 2, 0), (0, 256, 2, 2, 0), (3, 256, 1, 2, 0))
    if name == "64800x32400":       # DVB-S2: the staircase runs — single rows, pairs (ring depths 4 and 15), neither, compressed — for 40 iterations each
        variants = ((0, 128, 2, 1, 2), (4, 128, 2, 1, 2), (15, 128, 1, 1, 2), (0, 128, 2, 1, 1), (0, 256, 2, 1, 1), (3, 128, 2, 1, 0), (0, 128, 2, 2, 0))
    for stages, nc, tma, cmp, pipe in variants:
        d4 = pkg.CGPUDecoder(c, nb_frames=frames, kernel=4, semantics="ARM_SCALAR", early_term=1, fs_stages=stages, fs_nc=nc, fs_tma=tma, fs_g4=tma, fs_cmp=cmp, fs_pipe2=pipe,
                             fs_nostair=int(name == "64800x32400" and stages == 3), chunk_waves=1)
        h4, it4 = d4.decode(llr, 40, want_iters=True)
        d4.close()
        assert np.array_equal(h4, h1), f"{name} K{stages} NC{nc} tma{tma} cmp{cmp} pipe{pipe}: {(h4 != h1).any(axis=1).sum()} frames differ"
        assert np.array_equal(it4, it1), f"{name} K{stages} NC{nc} tma{tma} cmp{cmp} pipe{pipe}: iteration counts differ"
    assert it1.min() < 40 and it1.max() == 40


@pytest.mark.parametrize("kernel,name", [(4, "576x288"), (1, "576x288"), (4, "4000x2000")])
def test_frame_parallel_batch_quartered_over_stream_slots(built, kernel, name):
    """decode() cuts a frame-parallel batch of >= 8192 frames into four chunks on the four stream slots (the chunks' kernels run
    concurrently and their copies overlap): same bytes as one chunk (chunk_waves=1), oracle agreement on a sample, a ragged tail,
    iteration counts, and the packed format."""
    c = Code.load(name)
    F = 8192 + 4 * 517 + 3                                           # chunks of 2560 frames, the last one ragged
    llr = awgn_llr(c, F, 2.0, 901)
    dec = pkg.CGPUDecoder(c, nb_frames=F, device=0, kernel=kernel, early_term=int(kernel == 1), semantics="ARM_SCALAR" if kernel == 1 else "X86_SSE")
    hard, it = dec.decode(llr, 10, want_iters=True)
    one = pkg.CGPUDecoder(c, nb_frames=F, device=0, kernel=kernel, early_term=int(kernel == 1), semantics="ARM_SCALAR" if kernel == 1 else "X86_SSE", chunk_waves=1)
    hard1, it1 = one.decode(llr, 10, want_iters=True)
    assert dec.info(pkg.INFO_KERNEL) == kernel and np.array_equal(hard, hard1) and np.array_equal(it, it1)
    idx = np.r_[0:40, 2540:2580, F - 40:F]                           # across the first chunk boundary and the tail
    o = oracle_decode(c, dec.params, llr[idx], 10)
    assert np.array_equal(hard[idx], o["hard"]) and np.array_equal(it[idx], o["iters"])
    dec.close(); one.close()
    decp, decb = pkg.CGPUDecoder(c, nb_frames=F, device=0, kernel=kernel, out_format=1), pkg.CGPUDecoder(c, nb_frames=F, device=0, kernel=kernel)
    assert np.array_equal(decp.decode(llr, 10), oracle_pack(decb.decode(llr, 10), c.n))
    decp.close(); decb.close()


@pytest.mark.parametrize("kernel", [2, 1])
def test_ragged_sizes_and_packed_output(code576, kernel):
    big = awgn_llr(code576, 700, 1.5, 81)
    for frames in (1, 2, 3, 5, 31, 127, 129, 513, 700):
        llr = big[:frames]
        o = oracle_decode(code576, default_params(), llr, 4)
        for packed in (0, 1):
            dec = pkg.CGPUDecoder(code576, nb_frames=frames, kernel=kernel, out_format=packed)
            out = dec.decode(llr, 4)
            dec.close()
            want = oracle_pack(o["hard"], code576.n) if packed else o["hard"]
            assert out.shape == want.shape and np.array_equal(out, want), (frames, packed)
    dec = pkg.CGPUDecoder(code576, nb_frames=16, kernel=kernel)
    assert dec.decode(np.empty((0, code576.n), np.int8), 3).shape == (0, code576.n)     # empty batch
    out0 = dec.decode(big[:9], 0)                                                        # zero iterations = hard decision of the input
    assert np.array_equal(out0, (big[:9] > 0).astype(np.uint8))
    dec.close()


@pytest.mark.parametrize("name", ["155x93", "2640x1320", "1920x960"])
def test_arm_tree_tables_against_the_reference_scalar_decoder(built, name):
    """The tables only the ARM tree carries (code/ldpc_decoder_arm/Constantes: the (155,64) Tanner code with N % 16 != 0, 2640x1320,
    802.11e 1920x960) against fixture K8, minted from the ARM tree's scalar decoder: every int8 kernel that accepts the code, stop
    criterion on (iteration counts) and off."""
    c = Code.load(name)
    g8 = np.load(GOLD / f"k8_{name}_armscalar.npz")
    for key in [k[:-5] for k in g8.files if k.endswith("_hard")]:
        _, off, sv, sm, imax, early = key.split("_")
        for kernel in (0, 2, 1, 4) if c.n_checks >= 128 else (0, 2, 1):
            r = gpu_decode(c, g8["llr"], int(imax), want_iters=True, algo="OMS", semantics="ARM_SCALAR", offset=int(off), sat_var=int(sv), sat_msg=int(sm),
                           early_term=int(early), kernel=kernel)
            assert np.array_equal(np.packbits(r["hard"], axis=1, bitorder="little"), g8[key + "_hard"]), (key, kernel)
            assert np.array_equal(r["post"], g8[key + "_post"]) and np.array_equal(r["msgs"], g8[key + "_msgs"]) and np.array_equal(r["iters"], g8[key + "_iters"]), (key, kernel, r["kernel"])


def test_non_multiple_of_16_code_length(built):
    """N % 16 != 0 takes the scalar layout path in the reference (CDecoder_OMS_fixed_SSE.cpp:143-148); same here."""
    rng = np.random.default_rng(5)
    n, checks, d = 155, 93, 5
    pos = np.concatenate([rng.choice(n, d, replace=False) for _ in range(checks)]).astype(np.uint32)
    c = Code(n, checks, [d], [checks], pos)
    llr = stress_llr(c, 77, 91)
    for kernel in (2, 1):
        for packed in (0, 1):
            dec = pkg.CGPUDecoder(c, nb_frames=128, kernel=kernel, out_format=packed)
            out = dec.decode(llr, 5)
            o = oracle_decode(c, dec.params, llr, 5)
            dec.close()
            assert np.array_equal(out, oracle_pack(o["hard"], n) if packed else o["hard"]), (kernel, packed)


def test_full_batch_properties_64k(code576):
    """BASELINE config 1 at full size (65 536 frames): pipeline chunks agree with a single launch, a checksum over all
    outputs matches the oracle on a strided sample, frames are independent (permutation equivariance), and decoding is
    idempotent on converged frames."""
    F = 65536
    dec = pkg.CGPUDecoder(code576, nb_frames=F)
    llr = dec.awgn(F, pkg.sigma_for(2.0, 0.5), seed=2024)
    hard = dec.decode(llr, 10)
    sample = np.arange(0, F, 97)
    o = oracle_decode(code576, dec.params, llr[sample], 10, want_state=False)
    assert np.array_equal(hard[sample], o["hard"])
    perm = np.random.default_rng(1).permutation(F)
    hard_p = dec.decode(llr[perm], 10)
    assert np.array_equal(hard_p, hard[perm])
    fe = hard[:, :code576.k_info].any(axis=1)
    assert 0.02 < fe.mean() < 0.09          # FER ~ 0.05 at 2 dB
    # converged frames re-decode to themselves: feed +-31 LLRs of the decision
    conv = np.flatnonzero(~hard.any(axis=1))[:4096]
    again = dec.decode(np.where(hard[conv] > 0, 31, -31).astype(np.int8), 10)
    assert not again.any()
    dec.close()


def test_full_batch_every_frame_against_the_reference_decoder(code576):
    """BASELINE configs[1] at full size, EVERY frame: all 65 536 hard-decision rows of the GPU (on-chip kernel through the blocking
    host call, and the device-resident entry point) equal the reference's own x86 SSE decoder run on all host threads
    (oracle/_ref, CDecoder_OMS_fixed_SSE) — or the C restatement when the reference binary was not built here."""
    import os
    F = 65536
    dec = pkg.CGPUDecoder(code576, nb_frames=F)
    llr = dec.awgn(F, pkg.sigma_for(2.0, 0.5), seed=2024)
    hard = dec.decode(llr, 10)
    threads = len(os.sched_getaffinity(0))
    L = ref_x86("576x288")
    if L is not None:
        ref = np.empty((F, code576.n), np.uint8)
        assert L.ref_x86_decode_mt(ALGO["OMS"], 1, llr.ctypes.data, ref.ctypes.data, F, 10, threads) >= 0
    else:
        ref = oracle_decode_mt(code576, dec.params, llr, 10, threads)
    assert np.array_equal(hard, ref), f"{(hard != ref).any(axis=1).sum()} of {F} frames differ from the reference decoder"
    for algo, param in (("NMS", 29),):
        d2 = pkg.CGPUDecoder(code576, nb_frames=F, algo=algo, factor_q5=param)
        h2 = d2.decode(llr, 10)
        if L is not None:
            assert L.ref_x86_decode_mt(ALGO[algo], param, llr.ctypes.data, ref.ctypes.data, F, 10, threads) >= 0
        else:
            ref = oracle_decode_mt(code576, d2.params, llr, 10, threads)
        assert np.array_equal(h2, ref), f"{algo}: {(h2 != ref).any(axis=1).sum()} frames differ"
        d2.close()
    dec.close()


def test_python_layer_validates_shapes(code576):
    """The C side only sees pointers: the binding refuses arrays whose shape, dtype or layout would make it read or write out of bounds."""
    dec = pkg.CGPUDecoder(code576, nb_frames=64)
    llr = awgn_llr(code576, 8, 2.0, 1)
    for bad_llr in (llr[:, :288], np.zeros((3, 100), np.int8)):
        with pytest.raises(pkg.LdpcError):
            dec.decode(bad_llr, 2)
    for bad_out in (np.empty((4, 576), np.uint8), np.empty((8, 576), np.int8), np.empty((8, 1152), np.uint8)[:, ::2]):     # too small, wrong dtype, strided
        with pytest.raises(pkg.LdpcError):
            dec.decode(llr, 2, out=bad_out)
    out = np.empty((8, 576), np.uint8)
    assert dec.decode(llr, 2, out=out) is out
    dec.close()


def test_decode_device_on_two_streams_shares_slot_scratch_safely(built):
    """decode_device works in slot 0's scratch state (frame-parallel kernels): calls issued on two different streams without any
    synchronisation in between must still produce what they produce one after the other."""
    import torch
    c = Code.load("4000x2000")
    F = 8192
    dec = pkg.CGPUDecoder(c, nb_frames=F, kernel=4)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    d_llr = [torch.empty((F, c.n), dtype=torch.int8, device="cuda") for _ in range(2)]
    d_hard = [torch.empty((F, c.n), dtype=torch.uint8, device="cuda") for _ in range(4)]
    dec.awgn_device(d_llr[0].data_ptr(), F, pkg.sigma_for(2.0, 0.5), 11, 0, sa.cuda_stream)
    dec.awgn_device(d_llr[1].data_ptr(), F, pkg.sigma_for(1.0, 0.5), 12, 0, sa.cuda_stream)
    torch.cuda.synchronize()
    for k in range(2):                                   # reference: one call at a time
        dec.decode_device(d_llr[k].data_ptr(), d_hard[k].data_ptr(), F, 10, stream=sa.cuda_stream)
        torch.cuda.synchronize()
    for rep in range(3):                                 # back to back on two streams, no host synchronisation
        dec.decode_device(d_llr[0].data_ptr(), d_hard[2].data_ptr(), F, 10, stream=sa.cuda_stream)
        dec.decode_device(d_llr[1].data_ptr(), d_hard[3].data_ptr(), F, 10, stream=sb.cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(d_hard[0], d_hard[2]) and torch.equal(d_hard[1], d_hard[3])
    dec.close()


def test_quantiser_matches_reference_rule(code576):
    dec = pkg.CGPUDecoder(code576, nb_frames=16)
    rng = np.random.default_rng(3)
    y = np.concatenate([rng.normal(-1, 0.8, 100000), np.arange(-5, 5, 1 / 64.0), [0.0, -0.0, 3.875, -3.875, 1e9, -1e9]]).astype(np.float32)
    assert np.array_equal(dec.quantize(y), oracle_quantize(y))
    dec.close()


def test_device_resident_path_and_counters(code576):
    import torch
    F = 4096
    dec = pkg.CGPUDecoder(code576, nb_frames=F)
    d_llr = torch.empty((F, code576.n), dtype=torch.int8, device="cuda")
    d_hard = torch.empty((F, code576.n), dtype=torch.uint8, device="cuda")
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    st = ts.cuda_stream
    dec.awgn_device(d_llr.data_ptr(), F, pkg.sigma_for(2.0, 0.5), 77, 0, st)
    dec.decode_device(d_llr.data_ptr(), d_hard.data_ptr(), F, 10, stream=st)
    be, fe = dec.count_errors_device(d_hard.data_ptr(), F, st)
    hard = d_hard.cpu().numpy()
    host = dec.awgn(F, pkg.sigma_for(2.0, 0.5), 77, 0)
    assert np.array_equal(host, d_llr.cpu().numpy())                       # counter-based generator is reproducible
    o = oracle_decode(code576, dec.params, host[:512], 10, want_state=False)
    assert np.array_equal(hard[:512], o["hard"])
    assert be == int(hard[:, :code576.k_info].sum()) and fe == int(hard[:, :code576.k_info].any(axis=1).sum())
    dec.close()


def test_async_slots_overlap_and_agree(code576):
    F = 8192
    dec = pkg.CGPUDecoder(code576, nb_frames=F)
    slots = dec.info(pkg.INFO_STREAM_SLOTS)
    src = [pkg.PinnedArray((F, code576.n), np.int8) for _ in range(slots)]
    dst = [pkg.PinnedArray((F, code576.n), np.uint8) for _ in range(slots)]
    for s in range(slots):
        src[s].array[:] = awgn_llr(code576, F, 2.0, 100 + s)
    for s in range(slots):
        dec.decode_stream(s, src[s].array, dst[s].array, 10)
    dec.sync()
    for s in range(slots):
        o = oracle_decode(code576, dec.params, src[s].array[:256], 10, want_state=False)
        assert np.array_equal(dst[s].array[:256], o["hard"])
    dec.close()


@pytest.mark.parametrize("name", ["576x288", "1200x600", "2304x1152"])
def test_reference_gpu_kernels(built, name):
    """K6 of SURVEY 8c: the reference's OWN gpu_fixed kernels (LDPC_Sched_Stage_1_{MS,OMS,NMS,2NMS}_SIMD + (Inv)Interleaver_uint8,
    unmodified, cross-compiled for sm_100a into oracle/_ref) run on this GPU; the oracle's GPU_FIXED mode and this library's
    GPU_FIXED kernels must reproduce their hard decisions, posteriors and messages bit for bit — which pins that mode."""
    L = ref_gpu(name)
    if L is None:
        pytest.skip("oracle/_ref/libref_gpu_*.so not built (needs /root/reference at build time)")
    c = Code.load(name)
    llr = np.concatenate([awgn_llr(c, 512, 2.0, 301), awgn_llr(c, 256, 0.5, 302), stress_llr(c, 256, 303, full_range=True)])
    for algo in ("MS", "OMS", "NMS"):
        for iters in (1, 3, 10):
            r = ref_gpu_decode(L, algo, llr, iters)
            prm = default_params(algo=algo, semantics="GPU_FIXED")
            o = oracle_decode(c, prm, llr, iters)
            for k in ("hard", "post", "msgs"):
                assert np.array_equal(r[k], o[k]), f"{name} {algo} I{iters}: oracle GPU_FIXED differs from the reference kernel in {k} ({(r[k] != o[k]).sum()})"
        g = gpu_decode(c, llr, 10, algo=algo, semantics="GPU_FIXED")
        for k in ("hard", "post", "msgs"):
            assert np.array_equal(r[k], g[k]), f"{name} {algo}: this library differs from the reference kernel in {k}"
    # 2NMS is compiled with EARLY_TERM 1 (CUDA_2NMS_SIMD.cu:17): lanes of a warp that satisfy the test `break` past __syncthreads
    # calls the other lanes still execute (:286-291) — undefined behaviour that DEADLOCKS on sm_100 (observed: the kernel never
    # returns for 2 or more iterations on AWGN frames).  It can therefore only be run for one iteration, where the loop is not
    # entered; the 0.75 / 0.875 arithmetic it shares with NMS is pinned above.
    r1 = ref_gpu_decode(L, "2NMS", llr, 1)
    g1 = gpu_decode(c, llr, 1, algo="2NMS", semantics="GPU_FIXED")
    o1 = oracle_decode(c, g1["prm"], llr, 1)
    for k in ("hard", "post", "msgs"):
        assert np.array_equal(r1[k], g1[k]) and np.array_equal(r1[k], o1[k]), f"{name} 2NMS I1 {k}"


def test_dvbs2_full_batch_properties(built):
    """BASELINE configs[4] on a batch that fills the GPU's frame-parallel kernel (16 384 frames of 64 800 bits, 4.8 GB of state):
    strided sample against the oracle, staged kernel == plain kernel on every frame, frames independent."""
    c = Code.load("64800x32400")
    F = 16384
    dec = pkg.CGPUDecoder(c, nb_frames=F)
    assert dec.info(pkg.INFO_KERNEL) == 4
    llr = dec.awgn(F, pkg.sigma_for(1.0, 0.5), seed=31)
    hard = dec.decode(llr, 10)
    sample = np.arange(0, F, 1489)
    assert np.array_equal(hard[sample], oracle_decode(c, dec.params, llr[sample], 10, want_state=False)["hard"])
    d1 = pkg.CGPUDecoder(c, nb_frames=F, kernel=1)
    assert np.array_equal(d1.decode(llr, 10), hard)
    d1.close()
    perm = np.random.default_rng(3).permutation(F)
    assert np.array_equal(dec.decode(llr[perm], 10), hard[perm])
    assert hard.any()                                               # 1 dB is below this code's waterfall at 10 iterations: real work, not all-zero output
    dec.close()
    # the same with the per-frame stop criterion near the waterfall (frames stop at very different iterations inside one CTA):
    # staged == plain on every frame, iteration counts included
    e4, e1 = pkg.CGPUDecoder(c, nb_frames=4096, early_term=1), pkg.CGPUDecoder(c, nb_frames=4096, early_term=1, kernel=1)
    assert e4.info(pkg.INFO_KERNEL) == 4
    llr2 = np.concatenate([llr[:2048], e4.awgn(2048, pkg.sigma_for(1.8, 0.5), seed=32)])
    h4, it4 = e4.decode(llr2, 30, want_iters=True)
    h1, it1 = e1.decode(llr2, 30, want_iters=True)
    assert np.array_equal(h4, h1) and np.array_equal(it4, it1)
    assert it4.max() == 30 and (it4 < 30).sum() > 500
    e4.close(); e1.close()


def test_argument_errors_return_codes_not_exits(code576):
    """The reference exit(0)s on bad input (custom_cuda.cu:5-17); the library returns a status and keeps the handle usable."""
    dec = pkg.CGPUDecoder(code576, nb_frames=64, device=0, early_term=1)
    llr = awgn_llr(code576, 8, 2.0, 1)
    with pytest.raises(pkg.LdpcError) as e:
        dec.decode(llr, 300, want_iters=True)
    assert e.value.status == pkg.ERR_INVALID and "255" in str(e.value)
    with pytest.raises(pkg.LdpcError):
        dec.decode(llr, -1)
    hard, it = dec.decode(llr, 255, want_iters=True)              # still alive, and 255 iterations with early stop is fine
    assert hard.shape == (8, 576) and it.max() <= 255
    dec.close()


def test_four_decoder_objects_on_four_host_threads(code576):
    """The reference harness runs 4 decoder objects in 4 OpenMP sections against one GPU (code/gpu_fixed/test.cpp:241-281,347-393):
    one handle per host thread, no shared mutable state.  Four threads decode different batches at once, several times over, with
    different configurations; every result must equal the single-threaded one."""
    import threading
    cfgs = [dict(algo="OMS", semantics="X86_SSE"), dict(algo="NMS", semantics="GPU_FIXED"), dict(algo="OMS", semantics="ARM_SCALAR", early_term=1),
            dict(algo="OMS", semantics="UNIFORM", kernel=4)]
    batches = [awgn_llr(code576, 3000 + 17 * i, 1.0 + 0.5 * i, 400 + i) for i in range(4)]
    expect = []
    for kw, llr in zip(cfgs, batches):
        d = pkg.CGPUDecoder(code576, nb_frames=llr.shape[0], device=0, **kw); expect.append(d.decode(llr, 8)); d.close()
    results, errors = [None] * 4, []

    def work(i):
        try:
            d = pkg.CGPUDecoder(code576, nb_frames=batches[i].shape[0], device=0, **cfgs[i])
            for _ in range(5):
                results[i] = d.decode(batches[i], 8)
            d.close()
        except Exception as e:      # noqa: BLE001
            errors.append((i, repr(e)))

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]; [t.join() for t in th]
    assert not errors, errors
    for i in range(4):
        assert np.array_equal(results[i], expect[i]), f"thread {i} ({cfgs[i]})"


@pytest.mark.parametrize("name", ["2304x1152", "1248x624", "1944x972"])
def test_large_step_plans_with_early_termination_and_packing(built, name):
    """The 64-128-row steps of the on-chip plan (one frame pair over 2-4 warps): early termination with iteration counts, packed
    output, odd frame counts, and the 32-row-step plan of the same code as a cross-check."""
    c = Code.load(name)
    llr = np.concatenate([awgn_llr(c, 61, 2.5, 501), awgn_llr(c, 30, 1.0, 502), stress_llr(c, 12, 503)])
    for kw in (dict(algo="OMS", semantics="ARM_SCALAR", early_term=1), dict(algo="OMS", semantics="X86_SSE", early_term=1), dict(algo="NMS", semantics="GPU_FIXED")):
        g = gpu_decode(c, llr, 20, want_iters=True, **kw)
        assert g["kernel"] == 2
        o = oracle_decode(c, g["prm"], llr, 20)
        assert_same(g, o, f"{name} {kw}")
        assert np.array_equal(g["iters"], o["iters"])
        g2 = gpu_decode(c, llr, 20, want_iters=True, small_steps=1, **kw)
        assert_same(g2, o, f"{name} {kw} small steps")
    dec = pkg.CGPUDecoder(c, nb_frames=128, out_format=1)
    assert np.array_equal(dec.decode(llr, 6), oracle_pack(oracle_decode(c, dec.params, llr, 6)["hard"], c.n))
    dec.close()
