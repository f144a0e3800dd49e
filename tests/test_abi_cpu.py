"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/schwinger_b200.h declares, refuses to work without a GPU (no CPU fallback), and its
host-only configuration-file routines are byte-exact against the reference's SaveConf output."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_golden

import schwingermodel_b200 as sb
from schwingermodel_b200 import _abi


def test_library_is_built_and_exports_every_declared_symbol():
    names = sb.declared_symbols()
    assert len(names) >= 40 and len(set(names)) == len(names)
    lib = sb.load()
    for n in names:
        assert hasattr(lib, n), f"libschwinger_b200.so lacks {n}"
    # the Python binding types exactly the declared set
    assert set(_abi._SIGS) | {"sm_last_error"} == set(names)
    out = subprocess.run(["nm", "-D", "--defined-only", sb.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported


def test_built_for_sm_100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", sb.LIB_PATH], capture_output=True, text=True).stdout
    archs = {l.split(".")[-2] for l in out.splitlines() if "sm_" in l}
    assert archs == {"sm_100a"}, out


def test_one_pass_kernels_stage_rows_with_bulk_copies_and_nothing_uses_tensor_cores():
    """SASS of the shipped library: every k_dd_tma instantiation stages its rows with TMA bulk copies completing on
    mbarriers (UBLKCP + SYNCS), moves fields as 16-byte vectors, and no kernel holds a tensor-core instruction -- the
    path has no dense contraction."""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump is not installed")
    sass = subprocess.run(["cuobjdump", "-sass", sb.LIB_PATH], capture_output=True, text=True).stdout
    funcs, name = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            funcs[name] = []
        elif name is not None:
            funcs[name].append(line)
    tma = {n: "\n".join(b) for n, b in funcs.items() if "k_dd_tma" in n}
    assert len(tma) >= 3, sorted(funcs)          # PLAIN (3 and 4 stages), DOT, CG
    for n, body in tma.items():
        assert "UBLKCP" in body and "SYNCS" in body, n
        stores = [w for line in body.splitlines() for w in line.split() if w.startswith("STG.")]
        assert "LDS.128" in body and stores and all(w.endswith(".128") for w in stores if ".NA." in w), (n, set(stores))
        assert "LDGSTS" not in body, n           # no per-thread cp.async left in the TMA kernels
    everything = "\n".join("\n".join(b) for b in funcs.values())
    for mnemonic in ("UTCMMA", "UTCHMMA", "HMMA", "DMMA", "IMMA", "QGMMA", "HGMMA"):
        assert mnemonic not in everything, mnemonic


def test_no_gpu_no_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sb.SchwingerError) as e:
        sb.Lattice(8, 8)
    assert e.value.code == _abi.SM_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_argument_errors_are_reported_not_thrown():
    lib = sb.load()
    ctx = _abi.ctx_p()
    assert lib.sm_create(1, 8, 0, C.byref(ctx)) == _abi.SM_ERR_ARG
    assert b"2x2" in lib.sm_last_error()
    assert lib.sm_create(8, 8, 0, None) == _abi.SM_ERR_ARG
    assert lib.sm_local_dims(None, None) == _abi.SM_ERR_ARG
    assert lib.sm_create_dist(8, 8, 3, 1, 0, 0, b"x" * 128, C.byref(ctx)) == _abi.SM_ERR_ARG   # 8 % 3 != 0
    assert b"divisible" in lib.sm_last_error()


def test_config_file_bytes(tmp_path):
    g = load_golden(8, 8)
    f = tmp_path / "a.ctxt"
    sb.SaveConf(g["U"], 8, 8, str(f))
    want = open(os.path.join(GOLDEN, "ref_8x8.ctxt"), "rb").read()
    assert f.read_bytes() == want and len(want) == 2 * 64 * 28
    assert np.array_equal(sb.readBinary(8, 8, os.path.join(GOLDEN, "ref_8x8.ctxt")), g["U"])
    with pytest.raises(sb.SchwingerError):
        sb.readBinary(8, 8, str(tmp_path / "missing.ctxt"))


def test_file_name_tags_and_jackknife():
    s = np.load(os.path.join(GOLDEN, "ref_64x64_scalars.npz"))
    assert [sb.format_tag(2.0), sb.format_tag(-0.18), sb.format_tag(0.0)] == [str(x) for x in s["strings"]]
    from oracle.port import Port
    rng = np.random.default_rng(0)
    d = rng.normal(size=47)
    assert abs(sb.jackknife_error(d, 20) - Port(2, 2).jackknife(d, 20)) < 1e-15


def test_public_header_is_plain_c(tmp_path):
    """include/schwinger_b200.h is the C ABI: it must compile as C99 (no C++-isms, no torch / CUDA types) and as C++."""
    import subprocess
    src = tmp_path / "use_header.c"
    src.write_text('#include "schwinger_b200.h"\nint main(void) { sm_traj_result r; sm_hmc_params p; (void)r; (void)p; return SM_OK; }\n')
    inc = os.path.join(ROOT, "include")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only"],
                ["g++", "-std=c++17", "-Wall", "-fsyntax-only", "-x", "c++"]):
        r = subprocess.run(cmd + ["-I", inc, str(src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_ctypes_structs_match_the_header(tmp_path):
    """sm_traj_result / sm_hmc_params as the ctypes mirror declares them have the size and field offsets the C compiler
    gives the header's structs (a silent mismatch would corrupt every trajectory result)."""
    import subprocess
    fields = {"sm_traj_result": [f[0] for f in _abi.TrajResult._fields_], "sm_hmc_params": [f[0] for f in _abi.HmcParams._fields_]}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "schwinger_b200.h"', "int main(void) {"]
    for st, names in fields.items():
        lines.append(f'  printf("{st} %zu", sizeof({st}));')
        for n in names:
            lines.append(f'  printf(" %zu", offsetof({st}, {n}));')
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    src, exe = tmp_path / "layout.c", tmp_path / "layout"
    src.write_text("\n".join(lines))
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    for line, (st, cls) in zip(out, [("sm_traj_result", _abi.TrajResult), ("sm_hmc_params", _abi.HmcParams)]):
        got = [int(v) for v in line.split()[1:]]
        want = [C.sizeof(cls)] + [getattr(cls, f[0]).offset for f in cls._fields_]
        assert got == want, (st, got, want)
