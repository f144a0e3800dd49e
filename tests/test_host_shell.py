"""The C++ host shell (host/): reference-named entry points over the C ABI.

GPU tests drive host/bin/api_probe_* (which calls D_phi, conjugate_gradient, GaugeConf::...,
SaveConf through the reference's C++ interface) and the SM_NSxNT executable with the reference's
stdin protocol, and compare with the oracle.  The CPU test checks that the shell builds and
refuses to run without a GPU."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_golden, relerr

HOST = os.path.join(ROOT, "host")
BIN = os.path.join(HOST, "bin")


def _build(ns, nt):
    subprocess.run(["make", "-s", "-C", HOST, f"NS={ns}", f"NT={nt}"], check=True, capture_output=True)


def test_host_shell_builds_and_fails_loudly_without_gpu(tmp_path):
    import torch
    _build(8, 8)
    exe = os.path.join(BIN, "SM_8x8")
    assert os.path.exists(exe) and os.path.exists(os.path.join(BIN, "api_probe_8x8"))
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe], input="1\n1\n0\n4\n1\n2\n1\n20\n0\n0\n", capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    # the reference's prompts, in the reference's order, on stderr (src/main.cpp:30-58)
    order = ["ranks_x:", "ranks_t:", "m0:", "Molecular dynamics steps:", "Trajectory length:", "beta:",
             "Thermalization:", "Measurements:", "Step (sweeps between measurements):", "Save configurations yes/no"]
    pos = [r.stderr.index(s) for s in order]
    assert pos == sorted(pos)


def test_cmake_keeps_the_reference_configure_interface(tmp_path):
    """`cmake -DNS=.. -DNT=..` names the executable SM_${NS}x${NT} and fixes the lattice at configure time, as the reference's
    CMakeLists.txt:17-23 does; the GPU library is a CUDA target for sm_100a only.  (Configure step only: the build itself is
    what schwingermodel_b200/csrc/build.sh and host/Makefile do in-tree.)"""
    import shutil
    if shutil.which("cmake") is None:
        pytest.skip("cmake is not installed")
    r = subprocess.run(["cmake", "-S", ROOT, "-B", str(tmp_path), "-DNS=16", "-DNT=24"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    cache = open(os.path.join(tmp_path, "CMakeCache.txt")).read()
    assert "NS:STRING=16" in cache and "NT:STRING=24" in cache and "CMAKE_CUDA_ARCHITECTURES" in open(
        os.path.join(ROOT, "CMakeLists.txt")).read()
    targets = subprocess.run(["cmake", "--build", str(tmp_path), "--target", "help"], capture_output=True, text=True).stdout
    assert "SM_16x24" in targets and "schwinger_b200" in targets


@pytest.mark.gpu
@pytest.mark.parametrize("nx,nt", [(8, 8), (16, 24)])
def test_reference_cxx_interface_vs_oracle(tmp_path, nx, nt):
    from oracle.port import Port
    _build(nx, nt)
    g = load_golden(nx, nt)
    P = Port(nx, nt)
    V = nx * nt
    U, phi, m0, beta = g["U"], g["phi"], float(g["m0"]), float(g["beta"])
    conf = tmp_path / "in.ctxt"
    P.save_conf(U, str(conf))
    (tmp_path / "phi.bin").write_bytes(np.ascontiguousarray(phi).tobytes())
    out, saved = tmp_path / "out.bin", tmp_path / "saved.ctxt"
    r = subprocess.run([os.path.join(BIN, f"api_probe_{nx}x{nt}"), str(conf), str(tmp_path / "phi.bin"), repr(m0),
                        repr(beta), str(out), str(saved)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    buf = out.read_bytes()
    off = 0

    def take(dtype, n):
        nonlocal off
        a = np.frombuffer(buf, dtype=dtype, count=n, offset=off)
        off += a.nbytes
        return a

    assert np.array_equal(take(np.int32, 2 * V), g["tab_RightPB"])
    assert np.array_equal(take(np.int32, 2 * V), g["tab_LeftPB"])
    assert np.array_equal(take(np.complex128, 2 * V), g["tab_SignR"])
    assert np.array_equal(take(np.complex128, 2 * V), g["tab_SignL"])
    assert relerr(take(np.complex128, 2 * V).reshape(2, V), g["D"]) <= 1e-13
    assert relerr(take(np.complex128, 2 * V).reshape(2, V), g["Ddag"]) <= 1e-13
    dd = take(np.complex128, 2 * V).reshape(2, V)
    assert relerr(dd, g["DDdag"]) <= 1e-13
    z = take(np.complex128, 1)[0]
    assert abs(z - P.dot(phi, g["DDdag"])) <= 1e-11 * abs(z)
    ok, its = take(np.float64, 2)
    assert ok == 1 and abs(its + 2 - int(g["cg_apps"])) <= 1
    x = take(np.complex128, 2 * V).reshape(2, V)
    assert relerr(x, g["cg_x"]) <= 1e-9
    assert relerr(take(np.float64, 2 * V).reshape(2, V), g["fforce"]) <= 1e-9
    assert relerr(take(np.complex128, 2 * V).reshape(2, V), g["staple"]) <= 1e-14
    assert relerr(take(np.complex128, V), g["plaq"]) <= 1e-14
    sp, sg = take(np.float64, 2)
    assert abs(sp - g["plaq_sums"][0]) <= 1e-11 * V and abs(sg - g["plaq_sums"][1]) <= 1e-11 * V
    assert off == len(buf)
    assert saved.read_bytes() == conf.read_bytes()      # readBinary -> copy -> SaveConf is byte-exact


@pytest.mark.gpu
def test_executable_protocol_and_outputs(tmp_path):
    """printf params | SM_16x24: prompts on stderr, banner/results on stdout, SimData + .ctxt files."""
    _build(16, 24)
    params = "1\n1\n-0.05\n12\n0.6\n2\n40\n20\n0\n1\n"    # eps = 0.05: dH ~ 1 from the hot start
    env = dict(os.environ, SM_SEED="5", HOSTNAME="testhost")
    r = subprocess.run([os.path.join(BIN, "SM_16x24")], input=params, capture_output=True, text=True, cwd=tmp_path,
                       env=env, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "ranks_x:" in r.stderr and "Save configurations yes/no (1 or 0):" in r.stderr
    for s in ["* Nx = 16, Nt = 24", "* m0 = -0.05, kappa = ", "* Total number of MPI ranks = 1",
              "* Each rank has 384 lattice sites", "Thermalization done", "Average plaquette value / volume: Ep = ",
              "Acceptance rate: ", "Execution time = "]:
        assert s in r.stdout, s
    sim = (tmp_path / "2D_U1_16x24_m0-0.050000000000000003_SimData.txt").read_text().splitlines()
    assert sim[0] == "#Date and time" and sim[2] == "#Host" and sim[3] == "testhost"
    assert sim[4] == "#Nx      #Nt" and sim[5] == f"{16:>10}{24:>10}"
    assert sim[-8] == "#Ep                           #dEp" and sim[-2] == "#Execution time"
    ep = float(sim[-7].split()[0])
    acc = float(sim[-3])
    assert 0.0 < acc <= 1.0
    assert 0.3 < ep < 0.95
    confs = sorted(p.name for p in tmp_path.glob("2D_U1_Ns16_Nt24_b20000_m-00500_*.ctxt"))
    assert len(confs) == 20
    raw = (tmp_path / confs[-1]).read_bytes()
    assert len(raw) == 2 * 384 * 28
    rec = np.frombuffer(raw, dtype=np.dtype([("x", "<i4"), ("t", "<i4"), ("mu", "<i4"), ("re", "<f8"), ("im", "<f8")]))
    assert rec["x"][-1] == 15 and rec["t"][-1] == 23 and rec["mu"][-1] == 1
    assert np.abs(np.hypot(rec["re"], rec["im"]) - 1).max() < 1e-12     # links stay on the unit circle


_NUM = r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?"


def _masked(line):
    """numbers -> '#', runs of blanks -> one blank"""
    import re
    return re.sub(r"\s+", " ", re.sub(_NUM, "#", line)).strip()


@pytest.mark.gpu
def test_executable_transcript_equals_reference_executable(tmp_path):
    """Same stdin into host/bin/SM_16x24 and into the UNMODIFIED reference's own executable (oracle/_ref/SM_16x24, built
    from /root/reference/src/main.cpp by oracle/ref_build/build_ref.sh --exe; its transcript of the same input is also
    committed under tests/golden/ref_exe_16x24/): stderr (banner + prompts) byte for byte; stdout and _SimData.txt line
    by line -- exactly, except where the line holds a clock, a timing or a Monte-Carlo average (the reference seeds its
    generators from time(0) and random_device, src/main.cpp:17, src/hmc.cpp:7), where the text with numbers masked and
    the line's width must agree."""
    _build(16, 24)
    params = "1\n1\n-0.05\n12\n0.6\n2\n40\n20\n0\n0\n"
    env = dict(os.environ, SM_SEED="5", HOSTNAME="testhost")
    sim_name = "2D_U1_16x24_m0-0.050000000000000003_SimData.txt"
    mine = tmp_path / "mine"
    mine.mkdir()
    r = subprocess.run([os.path.join(BIN, "SM_16x24")], input=params, capture_output=True, text=True, cwd=mine, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stderr
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "SM_16x24")
    gold = os.path.join(GOLDEN, "ref_exe_16x24")
    refs = [(open(os.path.join(gold, "stdout.txt")).read(), open(os.path.join(gold, "stderr.txt")).read(),
             open(os.path.join(gold, "SimData.txt")).read(), "committed transcript")]
    if os.path.exists(ref_exe):                       # the live reference executable on this box's CPU
        theirs = tmp_path / "ref"
        theirs.mkdir()
        q = subprocess.run([ref_exe], input=params, capture_output=True, text=True, cwd=theirs, env=env, timeout=900)
        assert q.returncode == 0, q.stderr
        refs.append((q.stdout, q.stderr, (theirs / sim_name).read_text(), "live reference executable"))
    my_sim = (mine / sim_name).read_text()
    volatile_out = ("* Start time:", "Average plaquette value", "Average gauge action", "Acceptance rate:", "Execution time =")
    for ref_out, ref_err, ref_sim, what in refs:
        assert r.stderr == ref_err, what
        a, b = r.stdout.splitlines(), ref_out.splitlines()
        assert len(a) == len(b), (what, a, b)
        for x, y in zip(a, b):
            if x.startswith(volatile_out):
                assert _masked(x) == _masked(y), (what, x, y)
            else:
                assert x == y, (what, x, y)
        a, b = my_sim.splitlines(), ref_sim.splitlines()
        assert len(a) == len(b), what
        for i, (x, y) in enumerate(zip(a, b)):
            header = b[i - 1] if i else ""
            if header in ("#Date and time", "#Ep                           #dEp", "#gS                           #dgS",
                          "#Acceptance rate", "#Execution time"):
                assert _masked(x) == _masked(y) and (len(x) == len(y) or header == "#Date and time"), (what, x, y)
            else:
                assert x == y, (what, i, x, y)


@pytest.mark.gpu
def test_executable_forks_one_rank_per_gpu(tmp_path):
    """ranks_x = 2: the binary forks a second rank itself (no mpirun), NCCL between the two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _build(64, 64)
    params = "2\n1\n0\n10\n1\n2\n60\n20\n0\n1\n"      # BASELINE config 1 parameters, split over two GPUs
    env = dict(os.environ, SM_SEED="9", HOSTNAME="testhost")
    r = subprocess.run([os.path.join(BIN, "SM_64x64")], input=params, capture_output=True, text=True, cwd=tmp_path,
                       env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "* Total number of MPI ranks = 2" in r.stdout and "* Each rank has 2048 lattice sites" in r.stdout
    sim = (tmp_path / "2D_U1_64x64_m00_SimData.txt").read_text().splitlines()
    assert sim[7].split() == ["2", "1", "2"]
    ep, acc = float(sim[-7].split()[0]), float(sim[-3])
    assert 0.4 < ep < 0.9 and 0.0 < acc <= 1.0
    confs = sorted(tmp_path.glob("2D_U1_Ns64_Nt64_b20000_m00000_*.ctxt"))
    assert len(confs) == 20 and confs[0].stat().st_size == 2 * 4096 * 28
    rec = np.frombuffer(confs[-1].read_bytes(),
                        dtype=np.dtype([("x", "<i4"), ("t", "<i4"), ("mu", "<i4"), ("re", "<f8"), ("im", "<f8")]))
    assert np.abs(np.hypot(rec["re"], rec["im"]) - 1).max() < 1e-12
    # the gathered global field is consistent: its plaquette equals the measured one of the last configuration
    U = (rec["re"] + 1j * rec["im"]).reshape(64 * 64, 2).T
    import schwingermodel_b200 as sb
    lat = sb.Lattice(64, 64)
    _, sp, _ = lat.Compute_Plaquette01(np.ascontiguousarray(U), 2.0, want_field=False)
    lat.close()
    assert 0.3 < sp / 4096 < 0.95
