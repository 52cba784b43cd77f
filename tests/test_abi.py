"""CPU suite, part 2: the drop-in boundary without a GPU — the library loads, exports every symbol include/ldpc_b200.h declares,
the H-matrix loader / level schedule work, and compute entry points fail loudly instead of falling back to a CPU path."""
import ctypes as C
from pathlib import Path
import re

import numpy as np
import pytest

import ldpcgputegra_b200 as pkg
from _helpers import ROOT


def test_exports_match_header(built):
    hdr = (ROOT / "include" / "ldpc_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(ldpc_b200_\w+)\s*\(", hdr)))
    L = pkg.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert sorted(pkg.EXPORTS) == declared
    assert L.ldpc_b200_abi_version() == 1


def test_struct_layout_matches_header(built):
    assert C.sizeof(pkg.ParamsT) == 4 * 20
    assert C.sizeof(pkg.CodeT) == 4 * 4 + 4 * 8 * 2 + 8
    p = pkg.default_params()
    assert (p.algo, p.semantics, p.offset, p.factor_q5, p.sat_var, p.sat_msg, p.llr_scale, p.sat_llr) == (1, 0, 1, 29, 127, 31, 8, 31)


def test_bundled_tables_and_level_schedule(built):
    expect = {"576x288": (576, 288, 1824, 10), "1200x600": (1200, 600, 4818, 464), "1944x972": (1944, 972, 6966, 12),
              "2304x1152": (2304, 1152, 7296, 10), "4000x2000": (4000, 2000, 12000, 26), "64800x32400": (64800, 32400, 226799, 32399)}
    for name, (n, k, m, lv) in expect.items():
        c = pkg.Code.load(name)
        levels, level_of_row = c.level_schedule()
        assert (c.n, c.n_checks, c.m, levels) == (n, k, m, lv), name
        if n <= 4000:   # rows of one level share no variable
            e = 0; row = 0; used = {}
            for d, r in zip(c.deg, c.rows):
                for _ in range(r):
                    for v in c.pos[e:e + d]:
                        key = (int(level_of_row[row]), int(v))
                        assert key not in used, (name, row)
                        used[key] = row
                    e += d; row += 1


def test_header_parser_roundtrip(built, tmp_path):
    c = pkg.Code.load("576x288")
    rows = []
    e = 0
    for d, r in zip(c.deg, c.rows):
        for _ in range(r):
            rows.append(f"/* msg = {len(rows):6d}, deg = {d:2d} */ " + ", ".join(f"{v:6d}" for v in c.pos[e:e + d]) + ", ")
            e += d
    text = f"""#ifndef CONSTANTES
#define CONSTANTES
#define NB_DEGRES            {len(c.deg)}
#define _N                   {c.n} // Nombre de Variables
#define _K                   {c.n_checks} // Nombre de Checks
#define _M                   {c.m} // Nombre de Messages
#define NmoinsK     (_N-_K)
""" + "".join(f"#define DEG_{i + 1}                {d}\n#define DEG_{i + 1}_COMPUTATIONS   {r}\n" for i, (d, r) in enumerate(zip(c.deg, c.rows))) + \
        "#endif\nconst unsigned short PosNoeudsVariable[%d] ={\n" % c.m + "\n".join(rows).rstrip(", ") + "\n};\n"
    p = tmp_path / "constantes_sse.h"
    p.write_text(text)
    c2 = pkg.Code.from_header(p)
    assert (c2.n, c2.n_checks, c2.deg, c2.rows) == (c.n, c.n_checks, c.deg, c.rows) and np.array_equal(c2.pos, c.pos)
    out = tmp_path / "t.ldpc"
    c2.save(out)
    c3 = pkg.Code.load(str(out))
    assert np.array_equal(c3.pos, c.pos)
    # malformed inputs are refused, not guessed
    (tmp_path / "bad.h").write_text(text.replace(f"#define _M                   {c.m}", f"#define _M                   {c.m + 1}"))
    with pytest.raises(pkg.LdpcError):
        pkg.Code.from_header(tmp_path / "bad.h")
    bad = pkg.Code(c.n, c.n_checks, c.deg, c.rows, np.where(np.arange(c.m) == 5, c.n, c.pos))
    assert pkg.lib().ldpc_b200_check_code(C.byref(bad.c_struct())) == pkg.ERR_INVALID


def test_no_cpu_fallback(built):
    """Without a GPU, create() must fail with LDPC_ERR_NO_DEVICE — never decode on the CPU."""
    if pkg.lib().ldpc_b200_device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(pkg.LdpcError) as e:
        pkg.CGPUDecoder(pkg.Code.load("576x288"))
    assert e.value.status == pkg.ERR_NO_DEVICE
    p = C.c_void_p()
    assert pkg.lib().ldpc_b200_host_alloc(C.byref(p), 16) == pkg.ERR_NO_DEVICE


def test_invalid_configurations_are_rejected(built):
    c = pkg.Code.load("576x288")
    cs = c.c_struct()
    h = C.c_void_p()
    for kw in [dict(semantics="X86_SSE", algo="MS"), dict(sat_var=100), dict(dtype=7), dict(schedule=3), dict(factor_q5=999), dict(early_term=7),
               dict(dtype=1, semantics="X86_SSE"), dict(dtype=2, kernel=2), dict(schedule=1, kernel=1), dict(kernel=9)]:
        prm = pkg.default_params(**kw)
        rc = pkg.lib().ldpc_b200_create(C.byref(h), C.byref(cs), C.byref(prm), 0, 1024)
        assert rc in (pkg.ERR_INVALID, pkg.ERR_UNSUPPORTED), kw
        assert pkg.lib().ldpc_b200_last_error(None)


def test_product_path_never_imports_the_oracle():
    for p in list((ROOT / "ldpcgputegra_b200").rglob("*.py")) + list((ROOT / "ldpcgputegra_b200" / "csrc").glob("*")) + [ROOT / "include" / "ldpc_b200.h"]:
        if p.is_file() and p.suffix in (".py", ".cu", ".cuh", ".cpp", ".h"):
            text = p.read_text()
            for pat in (r'#include\s*[<"][^>"]*oracle', r'^\s*(from|import)\s+\S*(oracle|_helpers)', r'liboracle', r'oracle_decode', r'dlopen'):
                assert not re.search(pat, text, flags=re.M), (p, pat)


def test_encoder_analysis_runs_on_the_host(built):
    """The encoder's table analysis (peeling order, GF(2) inverse) is host code and runs before any device is touched: tables whose
    last n_checks columns are singular are refused as UNSUPPORTED with or without a GPU; good tables get as far as the device."""
    with pytest.raises(pkg.LdpcError) as e:
        pkg.Encoder(pkg.Code.load("2048x384"))
    assert e.value.status == pkg.ERR_UNSUPPORTED and "singular" in str(e.value)
    if pkg.lib().ldpc_b200_device_count() == 0:
        for name in ("576x288", "64800x32400"):
            with pytest.raises(pkg.LdpcError) as e:
                pkg.Encoder(pkg.Code.load(name))
            assert e.value.status == pkg.ERR_NO_DEVICE


REF = Path("/root/reference/code")


@pytest.mark.skipif(not REF.exists(), reason="reference tree not present on this machine")
def test_header_parser_on_the_real_reference_headers(built):
    """The H-matrix loader run over the reference's own headers where they lie — every x86-tree table (one file) and every gpu_fixed
    table (macros + index array in two files) — against the bundled .ldpc tables minted from them and the level counts of SURVEY App. C."""
    x86 = {"576x288": "576x288", "1944x972": "1944x972", "2048x384": "2048x384", "2304x1152": "2304x1152", "4000x2000": "4000x2000",
           "64800x32400": "64800x32400.dvb-s2", "64800x7200": "64800x7200.dvb-s2", "64800x6480": "64800x6480.dvb-s2"}
    seen = 0
    for name, d in x86.items():
        c = pkg.Code.from_header(REF / "x86/Constantes" / d / "constantes_sse.h")
        b = pkg.Code.load(name)
        assert (c.n, c.n_checks, c.m, c.deg, c.rows) == (b.n, b.n_checks, b.m, b.deg, b.rows) and np.array_equal(c.pos, b.pos), name
        assert sum(dg * r for dg, r in zip(c.deg, c.rows)) == c.m and sum(c.rows) == c.n_checks and int(c.pos.max()) < c.n
        seen += 1
    for d in sorted(p for p in (REF / "gpu_fixed/matrix").iterdir() if p.is_dir()):
        if not (d / "constantes_decoder.h").exists():
            continue
        g = pkg.Code.from_header(d / "constantes_gpu.h", d / "constantes_decoder.h")
        b = pkg.Code.load(d.name)
        assert (g.n, g.n_checks, g.deg, g.rows) == (b.n, b.n_checks, b.deg, b.rows) and np.array_equal(g.pos, b.pos), d.name
        seen += 1
    for name, d in {"155x93": "155x93", "2640x1320": "2640x1320", "1920x960": "802.11e.1920x960"}.items():      # the ARM tree's own tables
        c = pkg.Code.from_header(REF / "ldpc_decoder_arm/Constantes" / d / "constantes_sse.h")
        b = pkg.Code.load(name)
        assert (c.n, c.n_checks, c.deg, c.rows) == (b.n, b.n_checks, b.deg, b.rows) and np.array_equal(c.pos, b.pos), name
        seen += 1
    assert seen >= 23
    for name, levels in (("576x288", 10), ("2304x1152", 10), ("1200x600", 464), ("64800x32400", 32399)):
        assert pkg.Code.load(name).level_schedule()[0] == levels, name


def test_reference_arm_of_the_bench_runs_without_the_product_library(built, tmp_path):
    """`bench.py --impl reference` times the reference's CPU decoder and must not even map libldpc_b200.so: run it with the library
    path pointed at nothing (any use of the product would raise) and look at the process's own list of mapped objects."""
    import json, os, subprocess, sys
    env = dict(os.environ, LDPC_B200_LIB=str(tmp_path / "absent.so"), LDPC_BENCH_REPORT_MAPS="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--frames", "2048"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0
    assert not any("libldpc_b200" in m for m in line["mapped_objects"]), line["mapped_objects"]
    assert any("oracle" in m for m in line["mapped_objects"])
