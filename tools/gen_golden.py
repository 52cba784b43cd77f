#!/usr/bin/env python3
"""Mint known-answer fixtures from the REFERENCE's own decoders (oracle/_ref, built from /root/reference by oracle/Makefile).

The reference ships no golden vectors, so the fixtures under tests/golden/ are outputs of its decoders run here:
  K1/K2/K3  576x288  x86 SSE binary: OMS offset 1/2, NMS factor 29/24, I in {1,2,5,10}, AWGN @2 dB + saturating stress inputs
  K4        16 frames each of 1944x972, 2048x384, 2304x1152, 4000x2000, 64800x32400, I=10 (posteriors/messages as SHA-256 for the big ones)
  K5        ARM-tree scalar decoder with the stop criterion: per-frame iteration counts, I_max in {10,30}, 1..3 dB
  K8        the ARM tree's own tables (155x93, 2640x1320, 1920x960) through its scalar decoder, stop criterion on and off
  K7        ARM-tree scalar decoder at rails beyond int8 (+-2047/+-511, +-32767/+-8191, +-300/+-300): pins the int16 storage path
Run in the container that has /root/reference; the GPU box and CI only read the .npz files.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))
from _helpers import Code, awgn_llr, stress_llr, ref_x86, ref_x86_decode, ref_arm, ref_arm_decode  # noqa: E402

OUT = ROOT / "tests" / "golden"


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def mint_k4(names):
    for name in names:
        c = Code.load(name)
        Lc = ref_x86(name)
        assert Lc is not None, f"oracle/_ref/libref_x86_{name}.so missing (make -C oracle ref)"
        if c.k_info / c.n > 0.8:    # the rate-8/9 and 9/10 DVB-S2 tables: half of the AWGN frames above their waterfall, so that the decoder converges on some
            llr = np.concatenate([awgn_llr(c, 6, 2.0, 200 + c.n % 97), awgn_llr(c, 6, 5.0, 201 + c.n_checks % 97), stress_llr(c, 4, 300 + c.n % 89)])
        else:
            llr = np.concatenate([awgn_llr(c, 12, 2.0, 200 + c.n % 97), stress_llr(c, 4, 300 + c.n % 89)])
        out = {"llr": llr}
        algos = [("OMS", 1)] + ([("NMS", 29)] if len(c.deg) <= 2 else [])
        for algo, param in algos:
            r = ref_x86_decode(Lc, algo, param, llr, 10)
            key = f"{algo}_{param}_10"
            out[key + "_hard"] = np.packbits(r["hard"], axis=1, bitorder="little")
            out[key + "_sha"] = np.array([sha(r["post"]), sha(r["msgs"])])
        np.savez_compressed(OUT / f"k4_{name}_x86sse.npz", **out)


def mint_k7():
    """K7: int16 storage.  The only reference code with wider-than-int8 state is the ARM tree's scalar decoder (short arrays, run-time
    rails: code/ldpc_decoder_arm/CDecoder/template/CDecoder_fixed_x86.h:29-30, OMS/CDecoder_OMS_fixed_x86.cpp:30-32,61-200).  Rails far
    beyond int8 — setVarRange(+-2047) / setMsgRange(+-511) and +-32767 / +-8191 — full-range int8 inputs (its decode() takes signed
    char), 20 and 30 iterations so that posteriors and messages leave [-127, 127]; with and without the stop criterion."""
    code = Code.load("576x288")
    La = ref_arm("576x288")
    assert La is not None, "oracle/_ref/libref_arm_576x288.so missing (make -C oracle ref)"
    llr = np.concatenate([4 * awgn_llr(code, 24, 2.0, 701).astype(np.int16), 4 * awgn_llr(code, 16, 3.5, 702).astype(np.int16),
                          stress_llr(code, 16, 703, full_range=True).astype(np.int16), 3 * awgn_llr(code, 8, 0.5, 704).astype(np.int16)])
    llr = np.clip(llr, -128, 127).astype(np.int8)
    out = {"llr": llr}
    for (off, sv, sm) in [(1, 2047, 511), (3, 32767, 8191), (0, 300, 300)]:
        for imax, early in [(20, True), (30, False)]:
            r = ref_arm_decode(La, code, off, sv, sm, early, llr, imax)
            key = f"W_{off}_{sv}_{sm}_{imax}_{int(early)}"
            out[key + "_hard"] = np.packbits(r["hard"], axis=1, bitorder="little")
            out[key + "_iters"] = r["iters"]
            out[key + "_post"] = r["post"]; out[key + "_msgs"] = r["msgs"]
            print(key, "max |posterior|", int(np.abs(r["post"].astype(np.int32)).max()), "max |message|", int(np.abs(r["msgs"].astype(np.int32)).max()),
                  "iterations", int(r["iters"].min()), "..", int(r["iters"].max()))
    np.savez_compressed(OUT / "k7_576x288_armscalar_wide.npz", **out)


def mint_k8():
    """K8: the tables only the ARM tree carries (155x93 — N % 16 != 0 —, 2640x1320, 802.11e 1920x960) through the ARM tree's scalar decoder
    with its stop criterion: hard decisions, posteriors, messages, per-frame iteration counts (code/ldpc_decoder_arm/Constantes/*)."""
    for name in ("155x93", "2640x1320", "1920x960"):
        c = Code.load(name)
        La = ref_arm(name)
        assert La is not None, f"oracle/_ref/libref_arm_{name}.so missing (make -C oracle ref)"
        llr = np.concatenate([awgn_llr(c, 24, 2.5, 801), awgn_llr(c, 16, 4.0, 802), stress_llr(c, 8, 803)])
        out = {"llr": llr}
        for (off, sv, sm, imax, early) in [(1, 127, 31, 20, True), (1, 127, 31, 10, False)]:
            r = ref_arm_decode(La, c, off, sv, sm, early, llr, imax)
            key = f"A_{off}_{sv}_{sm}_{imax}_{int(early)}"
            out[key + "_hard"] = np.packbits(r["hard"], axis=1, bitorder="little")
            out[key + "_iters"] = r["iters"]
            out[key + "_post"] = r["post"].astype(np.int8); out[key + "_msgs"] = r["msgs"].astype(np.int8)
            print(name, key, "iterations", int(r["iters"].min()), "..", int(r["iters"].max()))
        np.savez_compressed(OUT / f"k8_{name}_armscalar.npz", **out)


def main():
    OUT.mkdir(exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "k8":
        mint_k8()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "k7":
        mint_k7()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "k4":      # mint only the named K4 fixtures (leaves the committed ones untouched)
        mint_k4(sys.argv[2:])
        return
    code = Code.load("576x288")
    L = ref_x86("576x288")
    assert L is not None, "build oracle/_ref first (make -C oracle)"
    llr = np.concatenate([awgn_llr(code, 48, 2.0, 101), stress_llr(code, 32, 102), stress_llr(code, 16, 103, full_range=True)])
    out = {"llr": llr}
    for algo, param in [("OMS", 1), ("OMS", 2), ("NMS", 29), ("NMS", 24)]:
        for iters in [1, 2, 5, 10]:
            r = ref_x86_decode(L, algo, param, llr, iters)
            key = f"{algo}_{param}_{iters}"
            out[key + "_hard"] = np.packbits(r["hard"], axis=1, bitorder="little")
            if iters in (1, 10):
                out[key + "_post"] = r["post"]; out[key + "_msgs"] = r["msgs"]
            else:
                out[key + "_sha"] = np.array([sha(r["post"]), sha(r["msgs"])])
    np.savez_compressed(OUT / "k123_576x288_x86sse.npz", **out)

    mint_k4(["1944x972", "2048x384", "2304x1152", "4000x2000", "64800x32400", "64800x7200", "64800x6480"])

    La = ref_arm("576x288")
    llr = np.concatenate([awgn_llr(code, 24, 1.0, 401), awgn_llr(code, 24, 2.0, 402), awgn_llr(code, 24, 3.0, 403), stress_llr(code, 8, 404)])
    out = {"llr": llr}
    for (off, sv, sm) in [(1, 127, 31), (1, 63, 15)]:
        for imax in [10, 30]:
            r = ref_arm_decode(La, code, off, sv, sm, True, llr, imax)
            key = f"ET_{off}_{sv}_{sm}_{imax}"
            out[key + "_hard"] = np.packbits(r["hard"], axis=1, bitorder="little")
            out[key + "_iters"] = r["iters"]
            out[key + "_post"] = r["post"].astype(np.int8); out[key + "_msgs"] = r["msgs"].astype(np.int8)
    np.savez_compressed(OUT / "k5_576x288_armscalar_et.npz", **out)
    mint_k7()
    mint_k8()
    for p in sorted(OUT.glob("*.npz")):
        print(p.name, p.stat().st_size)


if __name__ == "__main__":
    main()
