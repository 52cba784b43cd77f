"""The header-only C++ adapters (ldpcgputegra_b200/adapters): they must compile against the C ABI alone, link, and — on a GPU —
decode exactly what the Python layer decodes; inside the reference tree they must derive from the reference's own classes."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import ldpcgputegra_b200 as pkg
from _helpers import ROOT, awgn_llr

ADAPT = ROOT / "ldpcgputegra_b200" / "adapters"
LIBDIR = ROOT / "ldpcgputegra_b200"


def build_harness(tmp_path):
    exe = tmp_path / "adapter_harness"
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-I", str(ADAPT), "-I", str(ROOT / "include"), str(ROOT / "tests" / "cxx" / "adapter_harness.cpp"),
           "-o", str(exe), "-L", str(LIBDIR), "-lldpc_b200", f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_adapters_compile_and_fail_like_the_reference_without_a_gpu(built, tmp_path):
    exe = build_harness(tmp_path)
    if pkg.lib().ldpc_b200_device_count() > 0:
        pytest.skip("a GPU is visible here (covered by the gpu test)")
    llr = np.zeros((16, 576), np.int8); (tmp_path / "llr.bin").write_bytes(llr.tobytes())
    r = subprocess.run([str(exe), str(pkg.CODES_DIR / "576x288.ldpc"), "gpu", str(tmp_path / "llr.bin"), str(tmp_path / "out.bin"), "16", "5"], capture_output=True, text=True)
    # the reference prints "(EE) ..." and exit(0)s on a CUDA failure (custom_cuda.cu:5-17); so do the adapters
    assert r.returncode == 0 and "(EE) ldpc_b200 create failed" in r.stdout and not (tmp_path / "out.bin").exists()


def test_adapters_derive_from_the_reference_classes(built, tmp_path):
    ref = Path("/root/reference/code/gpu_fixed")
    if not ref.exists():
        pytest.skip("reference tree not present on this machine")
    src = tmp_path / "derive.cu"
    src.write_text('#include "decoder_template/CGPUDecoder.h"\n#define LDPC_B200_DERIVE_FROM_REFERENCE\n#include "CGPU_Decoder_B200.h"\n'
                   'extern const unsigned int PosNoeudsVariable[_M];\n'
                   'CGPUDecoder* make(size_t t) { static uint32_t tab[_M]; static ldpc_code_t c = ldpc_b200_adapters::code_from_reference_macros(PosNoeudsVariable, tab);\n'
                   '  return new CGPU_Decoder_B200(t, _N, _K, _M, c, "OMS"); }\n')
    r = subprocess.run(["/usr/local/cuda/bin/nvcc", "-c", "-w", "-I", str(ref), "-I", str(ADAPT), "-I", str(ROOT / "include"), str(src), "-o", str(tmp_path / "derive.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    # x86 flavour: CDecoder_B200 : CDecoder_fixed, fed from the compiled-in constantes_sse.h table (ref: code/x86/CDecoder/DecoderLibrary.h:64-69)
    x86 = Path("/root/reference/code/x86")
    src = tmp_path / "derive_x86.cpp"
    src.write_text('#include <string>\nusing namespace std;\n#include "Constantes/constantes_sse.h"\n#include "CDecoder/template/CDecoder_fixed.h"\n'
                   '#define LDPC_B200_DERIVE_FROM_REFERENCE\n#include "CGPU_Decoder_B200.h"\n'
                   'CDecoder* make() { static uint32_t tab[_M]; static ldpc_code_t c = ldpc_b200_adapters::code_from_reference_macros(PosNoeudsVariable, tab);\n'
                   '  CDecoder_B200* d = new CDecoder_B200(c, "OMS", 16); d->setOffset(1); d->setVarRange(-127, 127); d->setMsgRange(-31, 31); return d; }\n')
    r = subprocess.run(["g++", "-std=c++14", "-c", "-w", "-msse4.1", "-I", str(x86), "-I", str(x86 / "CDecoder" / "template"), "-I", str(ADAPT), "-I", str(ROOT / "include"),
                        str(src), "-o", str(tmp_path / "derive_x86.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("flavour,sem", [("gpu", "GPU_FIXED"), ("x86", "X86_SSE")])
def test_adapters_decode_like_the_python_layer(built, tmp_path, code576, flavour, sem):
    exe = build_harness(tmp_path)
    llr = awgn_llr(code576, 512, 1.5, 91)
    (tmp_path / "llr.bin").write_bytes(llr.tobytes())
    r = subprocess.run([str(exe), str(pkg.CODES_DIR / "576x288.ldpc"), flavour, str(tmp_path / "llr.bin"), str(tmp_path / "out.bin"), "512", "7"], capture_output=True, text=True)
    assert r.returncode == 0 and "decoded 512 frames" in r.stdout, r.stdout + r.stderr
    out = np.frombuffer((tmp_path / "out.bin").read_bytes(), np.uint8).reshape(512, 576)
    dec = pkg.CGPUDecoder(code576, nb_frames=512, device=0, algo="OMS", semantics=sem)
    assert np.array_equal(out, dec.decode(llr, 7))
    dec.close()
