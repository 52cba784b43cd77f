"""GPU suite: the H-derived systematic encoder and decoding of NON-ZERO codewords.
The reference simulates the all-zero codeword unless `-encoder` is given (code/x86/main_p.cpp:232-233); these tests check that
nothing in the decoders depends on that: H c = 0 for every encoded frame, the information part is systematic, decoders recover
random codewords, and the sign-symmetric decoders commute exactly with flipping the channel values by a codeword."""
import numpy as np
import pytest
import torch

import ldpcgputegra_b200 as pkg
from _helpers import Code, oracle_decode

pytestmark = pytest.mark.gpu


def syndrome(code, cw):
    """H c over GF(2), rows in table order; cw [F, n] bytes"""
    out = np.zeros((cw.shape[0], code.n_checks), np.uint8)
    e = r = 0
    for d, cnt in zip(code.deg, code.rows):
        idx = code.pos[e:e + d * cnt].reshape(cnt, d).astype(np.int64)
        out[:, r:r + cnt] = cw[:, idx].sum(axis=2) & 1
        e += d * cnt; r += cnt
    return out


@pytest.mark.parametrize("name", ["576x288", "2304x1152", "1944x972", "1200x600", "1248x624", "64800x32400"])
def test_encoder_produces_codewords(built, name):
    code = Code.load(name)
    enc = pkg.Encoder(code)
    phases, dense = enc.info()
    rng = np.random.Generator(np.random.Philox(11))
    F = 70 if code.n < 10000 else 33
    info = rng.integers(0, 2, size=(F, code.k_info), dtype=np.uint8)
    info[0] = 0; info[1] = 1
    cw = enc.encode(info)
    assert cw.shape == (F, code.n) and set(np.unique(cw)) <= {0, 1}
    assert np.array_equal(cw[:, :code.k_info], info), "information bits must be the first n - n_checks positions"
    assert not syndrome(code, cw).any(), f"{name}: H c != 0 ({phases} phases, {dense} dense unknowns)"
    assert not cw[0].any()                                            # linearity: the zero word
    assert np.array_equal(enc.encode(info[2:3] ^ info[3:4]), cw[2:3] ^ cw[3:4])
    if name == "64800x32400":
        assert dense == 0, "an IRA (staircase) code is solved by peeling alone, like the reference's accumulate + staircase encoder"
    # random information bits generated on the device: reproducible per (seed, first_frame), different across frames
    d_cw = torch.empty((64, code.n), dtype=torch.uint8, device="cuda")
    enc.encode_device(d_cw.data_ptr(), 64, seed=5, first_frame=64)
    a = d_cw.cpu().numpy()
    enc.encode_device(d_cw.data_ptr(), 32, seed=5, first_frame=96)
    torch.cuda.synchronize()
    assert np.array_equal(d_cw.cpu().numpy()[:32], a[32:]) and not syndrome(code, a).any()
    assert 0.4 < a[:, :code.k_info].mean() < 0.6 and len({r.tobytes() for r in a}) == 64
    enc.close()


def test_unsupported_tables_fail_loudly(built):
    for name in ["4000x2000", "2048x384"]:          # the last n_checks columns of these tables are not an information-set complement
        with pytest.raises(pkg.LdpcError) as e:
            pkg.Encoder(Code.load(name))
        assert e.value.status == pkg.ERR_UNSUPPORTED


@pytest.mark.parametrize("kw", [dict(algo="OMS", semantics="UNIFORM"), dict(algo="NMS", semantics="UNIFORM"),
                                dict(dtype="F32", algo="NMS", schedule="FLOODING"), dict(dtype="I16", semantics="ARM_SCALAR", sat_var=2047, sat_msg=511)])
def test_codeword_symmetry_is_exact(code576, kw):
    """decode(flip_c(LLR)) == decode(LLR) xor c for the sign-symmetric semantics (UNIFORM rails +-127, ARM scalar, float):
    flipping the channel values where the codeword has a 1 must flip exactly those decisions, posteriors and messages."""
    dec = pkg.CGPUDecoder(code576, nb_frames=512, device=0, **kw)
    enc = pkg.Encoder(code576)
    rng = np.random.Generator(np.random.Philox(21))
    cw = enc.encode(rng.integers(0, 2, size=(512, code576.k_info), dtype=np.uint8))
    llr0 = dec.awgn(512, pkg.sigma_for(1.5, 0.5), seed=9)                  # all-zero codeword
    if llr0.dtype != np.float32 and kw.get("semantics") == "ARM_SCALAR":
        llr0 = (llr0.astype(np.int16) * 8).astype(np.int16)
    sign = np.where(cw == 1, -1, 1).astype(llr0.dtype)
    dec.set_debug(True)
    h0 = dec.decode(llr0, 10).copy(); p0, m0 = dec.debug_state(512)
    hc = dec.decode(llr0 * sign, 10); pc, mc = dec.debug_state(512)
    zero = p0 == 0                                                         # bit = (posterior > 0): a zero posterior decides 0 on both sides
    assert np.array_equal((hc ^ cw)[~zero], h0[~zero])
    assert np.array_equal(pc, p0 * sign)
    edge_sign = sign[:, code576.pos.astype(np.int64)]
    assert np.array_equal(mc, m0 * edge_sign)
    dec.close(); enc.close()


def test_device_chain_with_real_codewords(code576):
    """encode -> BPSK + AWGN -> decode -> count against the transmitted word, all on the device.  The all-zero chain at the same
    Eb/N0 must give the same FER within sampling error (the noise samples coincide but land on different signs, so the two
    counts are two draws of the same statistic, not equal numbers)."""
    F = 16384
    dec = pkg.CGPUDecoder(code576, nb_frames=F, device=0)                  # reference semantics (X86_SSE OMS)
    enc = pkg.Encoder(code576)
    d_cw = torch.empty((F, 576), dtype=torch.uint8, device="cuda")
    d_llr = torch.empty((F, 576), dtype=torch.int8, device="cuda")
    d_hard = torch.empty((F, 576), dtype=torch.uint8, device="cuda")
    sigma = pkg.sigma_for(2.0, 0.5)
    enc.encode_device(d_cw.data_ptr(), F, seed=3, first_frame=0)
    torch.cuda.synchronize()
    dec.awgn_codeword_device(d_llr.data_ptr(), d_cw.data_ptr(), F, sigma, seed=4)
    dec.decode_device(d_llr.data_ptr(), d_hard.data_ptr(), F, 10)
    be, fe = dec.count_errors_ref_device(d_hard.data_ptr(), d_cw.data_ptr(), F)
    torch.cuda.synchronize()
    hard, cw, llr = d_hard.cpu().numpy(), d_cw.cpu().numpy(), d_llr.cpu().numpy()
    diff = (hard ^ cw)[:, :288]
    assert (be, fe) == (int(diff.sum()), int(diff.any(axis=1).sum()))
    assert np.array_equal(hard[:256], oracle_decode(code576, dec.params, llr[:256], 10)["hard"])   # still bit-exact against the oracle
    dec.awgn_device(d_llr.data_ptr(), F, sigma, seed=4)                    # the all-zero chain on the same noise
    dec.decode_device(d_llr.data_ptr(), d_hard.data_ptr(), F, 10)
    be0, fe0 = dec.count_errors_device(d_hard.data_ptr(), F)
    assert 0.04 < fe / F < 0.062 and 0.04 < fe0 / F < 0.062 and abs(fe - fe0) <= 4.0 * np.sqrt(fe + fe0), (fe, fe0)
    assert be0 > 0 and 8 < be / fe < 16
    dec.close(); enc.close()
