"""GPU suite: BASELINE configs[2] (float normalised min-sum, flooding, syndrome early termination, Eb/N0 sweep 0-4 dB) as curve
parity against the CPU restatement on identical channel values, and the reference-compatible C simulator (harness/ldpc_sim)."""
import re
import subprocess

import numpy as np
import pytest

import ldpcgputegra_b200 as pkg
from _helpers import ROOT, oracle_decode, oracle_decode_float

pytestmark = pytest.mark.gpu


def test_float_flooding_curve_parity_vs_cpu(code576):
    """Every point of the sweep: the GPU's own channel values are decoded by the GPU and by the CPU restatement; the BER/FER
    points must be IDENTICAL (same hard decisions, same iteration counts), and the curve must fall with Eb/N0."""
    dec = pkg.CGPUDecoder(code576, nb_frames=2048, device=0, dtype="F32", algo="NMS", factor1=0.75, schedule="FLOODING", early_term=1)
    fers = []
    for i, ebn0 in enumerate([0.0, 1.0, 2.0, 3.0, 4.0]):
        y = dec.awgn(1536, pkg.sigma_for(ebn0, 0.5), seed=77, first_frame=i * 4096)
        hard, it = dec.decode(y, 50, want_iters=True)
        o = oracle_decode_float(code576, dec.params, y, 50)
        assert np.array_equal(hard, o["hard"]) and np.array_equal(it, o["iters"]), f"Eb/N0 = {ebn0} dB"
        info = hard[:, :code576.k_info]
        fers.append(float(info.any(axis=1).mean()))
    dec.close()
    assert fers[0] > 0.9 and fers[4] < 0.01 and all(a >= b for a, b in zip(fers, fers[1:])), fers


def test_fixed_and_float_curves_agree_statistically(code576):
    """K8 of SURVEY 8c: int8 layered OMS (reference semantics, pinned) and float layered NMS see the same noise realisations;
    at 2 dB / 10 iterations both FERs sit near the survey's probe value 0.05 and within a few binomial sigmas of each other."""
    F = 32768
    d8 = pkg.CGPUDecoder(code576, nb_frames=F, device=0)
    df = pkg.CGPUDecoder(code576, nb_frames=F, device=0, dtype="F32", algo="NMS", factor1=0.75)
    sigma = pkg.sigma_for(2.0, 0.5)
    fe8 = d8.decode(d8.awgn(F, sigma, seed=5), 10)[:, :code576.k_info].any(axis=1).mean()
    fef = df.decode(df.awgn(F, sigma, seed=5), 10)[:, :code576.k_info].any(axis=1).mean()
    d8.close(); df.close()
    assert 0.04 < fe8 < 0.062 and 0.02 < fef < 0.062, (fe8, fef)     # float sees unquantised values: at least as good


def run_sim(*args):
    exe = ROOT / "harness" / "ldpc_sim"
    r = subprocess.run([str(exe), *map(str, args)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    pts = {}
    for m in re.finditer(r"SNR = ([\d.]+) \| BER =\s+([\d.e+-]+) \| FER =\s+([\d.e+-]+) .*?MATRICES =\s*(\d+)\| FE = (\d+) \| BE = (\d+)", r.stdout):
        pts[float(m.group(1))] = dict(ber=float(m.group(2)), fer=float(m.group(3)), frames=int(m.group(4)), fe=int(m.group(5)), be=int(m.group(6)))
    return r.stdout, pts


def test_c_simulator_reference_cli(built, code576):
    """ldpc_sim with the reference's own options: fixed-point OMS offset 1, 10 iterations, 1-3 dB.  FER must match the survey's
    probe of the reference decoder (0.671 / 0.050 / 5e-5 at 1 / 2 / 3 dB) within sampling error, and the 2 dB point must equal
    what the Python layer gets for the same seed and frame range (same channel counters)."""
    out, pts = run_sim("-fixed", "-sse", "-OMS", 1, "-iter", 10, "-min", 1, "-max", 3.01, "-pas", 1, "-fer", 200, "-frames", 65536, "-max-frames", 262144)
    assert "(II) Code LDPC (N, K)     : (576,288)" in out and "OFFSET MIN-SUM" in out and "row-parallel, on-chip state" in out
    assert set(pts) == {1.0, 2.0, 3.0}
    assert 0.62 < pts[1.0]["fer"] < 0.72 and 0.043 < pts[2.0]["fer"] < 0.058 and pts[3.0]["fer"] < 4e-4
    assert pts[3.0]["frames"] == 262144 and pts[1.0]["frames"] == 65536
    # same numbers through the Python mirror: point index 1 (2 dB) uses seed + 1, frames 0..65535 first
    dec = pkg.CGPUDecoder(code576, nb_frames=65536, device=0)
    hard = dec.decode(dec.awgn(65536, pkg.sigma_for(2.0, 0.5), seed=1 + 1, first_frame=0), 10)
    dec.close()
    info = hard[:, :code576.k_info]
    assert pts[2.0]["frames"] == 65536 and pts[2.0]["fe"] == int(info.any(axis=1).sum()) and pts[2.0]["be"] == int(info.sum())


def test_c_simulator_float_flooding_and_code_header(built, tmp_path):
    out, pts = run_sim("-float", "-NMS", 0.75, "-flooding", "-early", "-iter", 40, "-min", 3, "-max", 3, "-fer", 20, "-frames", 8192, "-max-frames", 65536)
    assert "FLOODING + SYNDROME STOP" in out and "generic engine (fp32 arithmetic, on-chip state" in out and 3.0 in pts and pts[3.0]["fer"] < 0.01
    # H-matrix load from a header in the reference's own format
    code = pkg.Code.load("200x100")
    hdr = tmp_path / "constantes_sse.h"
    hdr.write_text("#define _N %d\n#define _K %d\n#define _M %d\n#define NB_DEGRES %d\n" % (code.n, code.n_checks, code.m, len(code.deg))
                   + "".join(f"#define DEG_{i+1} {d}\n#define DEG_{i+1}_COMPUTATIONS {r}\n" for i, (d, r) in enumerate(zip(code.deg, code.rows)))
                   + "const unsigned short PosNoeudsVariable[_M] = {" + ", ".join(map(str, code.pos.tolist())) + "};\n")
    out, pts = run_sim("-fixed", "-gpu", "-OMS", 1, "-iter", 5, "-min", 2, "-max", 2, "-header", hdr, "-frames", 4096, "-max-frames", 4096)
    assert "(200,100)" in out and pts[2.0]["frames"] == 4096


def test_c_simulator_real_encoder(built):
    """-encoder: random codewords instead of the all-zero word; the FER point must sit where the all-zero one does"""
    out, pts = run_sim("-fixed", "-avx", "-OMS", 1, "-iter", 10, "-min", 2, "-max", 2, "-fer", 1000, "-frames", 65536, "-max-frames", 65536, "-encoder")
    out0, pts0 = run_sim("-fixed", "-avx", "-OMS", 1, "-iter", 10, "-min", 2, "-max", 2, "-fer", 1000, "-frames", 65536, "-max-frames", 65536)
    assert "systematic, derived from H" in out and "all-zero codeword" in out0
    fe, fe0 = pts[2.0]["fe"], pts0[2.0]["fe"]
    assert 0.04 < pts[2.0]["fer"] < 0.062 and abs(fe - fe0) <= 4.0 * (fe + fe0) ** 0.5, (fe, fe0)


def test_c_simulator_real_encoder_many_batches(built, code576):
    """-encoder over MANY small batches: encoder, channel, decoder and counters of consecutive batches run on one stream (the encoder
    has no handle: the harness passes it the decoder's slot-0 stream).  A batch whose channel read a half-written or a previous
    batch's codeword would count about half of its bits as errors; the FER must stay at the all-zero-codeword value instead."""
    out, pts = run_sim("-fixed", "-sse", "-OMS", 1, "-iter", 10, "-min", 2.5, "-max", 2.5, "-fer", 100000, "-frames", 512, "-max-frames", 131072, "-encoder")
    out0, pts0 = run_sim("-fixed", "-sse", "-OMS", 1, "-iter", 10, "-min", 2.5, "-max", 2.5, "-fer", 100000, "-frames", 512, "-max-frames", 131072)
    assert pts[2.5]["frames"] == 131072 and pts0[2.5]["frames"] == 131072             # 256 batches each
    fe, fe0 = pts[2.5]["fe"], pts0[2.5]["fe"]
    assert fe0 > 50 and abs(fe - fe0) <= 5.0 * (fe + fe0) ** 0.5, (fe, fe0)
    assert pts[2.5]["ber"] < 3 * pts0[2.5]["ber"] + 1e-6


def test_encoder_null_stream_is_safe_without_synchronisation(built, code576):
    """ldpc_b200_encode_device(..., stream NULL) followed at once by awgn_codeword_device(..., NULL): NULL means the legacy default
    stream for the handle-less encoder and the handle's non-blocking slot-0 stream for the channel, which nothing orders — the
    encoder therefore finishes before it returns.  Checked by decoding noiseless codewords of many consecutive batches."""
    import torch
    enc = pkg.Encoder(code576, device=0)
    F = 4096
    dec = pkg.CGPUDecoder(code576, nb_frames=F, device=0)
    d_cw = torch.empty((F, code576.n), dtype=torch.uint8, device="cuda")
    d_llr = torch.empty((F, code576.n), dtype=torch.int8, device="cuda")
    d_hard = torch.empty((F, code576.n), dtype=torch.uint8, device="cuda")
    for b in range(24):
        enc.encode_device(d_cw.data_ptr(), F, seed=5, first_frame=b * F)              # stream 0 = NULL
        dec.awgn_codeword_device(d_llr.data_ptr(), d_cw.data_ptr(), F, 0.05, seed=9, first_frame=b * F)   # NULL = slot-0 stream
        dec.decode_device(d_llr.data_ptr(), d_hard.data_ptr(), F, 3)
        be, fe = dec.count_errors_ref_device(d_hard.data_ptr(), d_cw.data_ptr(), F)
        assert (be, fe) == (0, 0), f"batch {b}: {be} bit errors at sigma 0.05"
        assert int(d_cw.sum().item()) > F * code576.n // 4                              # random codewords, not the all-zero word
    dec.close(); enc.close()


@pytest.mark.parametrize("name,algo", [("576x288", "-OMS"), ("576x288", "-NMS"), ("2304x1152", "-OMS")])
def test_the_reference_simulator_itself_linked_against_this_library(built, name, algo):
    """The reference's OWN simulator — code/gpu_fixed/main.cpp and every non-decoder source, unmodified, compiled where they lie by
    oracle/Makefile — linked against libldpc_b200.so through oracle/ref_main_shim.cu in place of the reference's decoder .cu files.
    One Eb/N0 point: the `SNR = ... | BER = ... | FER = ...` line it prints must agree with harness/ldpc_sim -gpu on the same point
    within sampling error (the two draw different noise: cuRAND XORWOW seed 1234 there, counter-based Philox here)."""
    exe = ROOT / "oracle" / "_ref" / f"ref_gpu_main_{name}"
    if not exe.exists():
        pytest.skip("oracle/_ref/ref_gpu_main_* not built (needs /root/reference at build time)")
    ebn0 = 2.0
    r = subprocess.run([str(exe), algo, "-min", str(ebn0), "-max", str(ebn0), "-iter", "10", "-fer", "600", "-n", "16384"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "libldpc_b200" in r.stdout and f"(II) Code LDPC (N, K)     : ({name.split('x')[0]}," in r.stdout
    m = re.search(r"SNR = ([\d.]+) \| BER =\s+([\d.e+-]+) \| FER =\s+([\d.e+-]+) .*?MATRICES =\s*(\d+)\| FE = (\d+) \| BE = (\d+)", r.stdout)
    assert m, r.stdout[-1500:]
    frames, fe, be = int(m.group(4)), int(m.group(5)), int(m.group(6))
    assert frames >= 65536 and fe >= 600 and abs(float(m.group(1)) - ebn0) < 1e-6
    out, pts = run_sim("-fixed", "-gpu", algo, "1" if algo == "-OMS" else "0.75", "-iter", 10, "-min", ebn0, "-max", ebn0, "-fer", 600, "-code", name, "-frames", 65536, "-max-frames", 4 * 65536)
    p = pts[ebn0]
    fer_a, fer_b = fe / frames, p["fe"] / p["frames"]
    sigma = (fer_a * (1 - fer_a) / frames + fer_b * (1 - fer_b) / p["frames"]) ** 0.5
    assert abs(fer_a - fer_b) < 5 * sigma + 1e-4, (fer_a, fer_b, sigma)
    ber_a, ber_b = float(m.group(2)), p["ber"]          # both count over the information part (ref: ber_analyzer/CErrorAnalyzer.cpp:119-159)
    assert be > 0 and 0.4 < (ber_a + 1e-12) / (ber_b + 1e-12) < 2.5, (ber_a, ber_b)
