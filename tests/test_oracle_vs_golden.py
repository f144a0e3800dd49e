"""The C restatement (oracle/schwinger_oracle.c) against golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  Bit-exact: same operation order, same libm/libgcc."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle.port import Port, gaussian_fields

CASES = [(8, 8), (16, 24), (32, 32)]


@pytest.fixture(scope="module", params=CASES, ids=lambda c: f"{c[0]}x{c[1]}")
def case(request):
    nx, nt = request.param
    g = load_golden(nx, nt)
    return Port(nx, nt), g


def test_hot_start_bit_exact(case):
    P, g = case
    assert np.array_equal(P.hot_start(12345), g["U"])


def test_hot_start_fingerprint():
    # srand(12345) -> U0[0] with glibc rand()
    U = Port(64, 64).hot_start(12345)
    assert U[0, 0] == complex(0.43488050387792643, 0.90048817168626971)


def test_tables_bit_exact(case):
    P, g = case
    t = P.tables()
    for k, v in t.items():
        assert np.array_equal(v, g["tab_" + k]), k
    t22 = P.tables(2, 2, 3)
    flat = np.concatenate([v.view(np.float64).ravel() if v.dtype == np.complex128 else v.astype(np.float64)
                           for v in t22.values()])
    assert np.array_equal(flat, g["tab22_rank3"])


def test_table_spot_values():
    # 32x32 single rank
    t = Port(32, 32).tables()
    assert t["SignR"][2 * 31] == -1 and t["SignL"][0] == -1
    assert t["RightPB"][2 * 31] == 0 and t["LeftPB"][0] == 31 and t["LeftPB"][1] == 992
    assert t["x_1_t1"][0] == 993 and t["x1_t_1"][0] == 63


def test_operators_bit_exact(case):
    P, g = case
    U, phi, m0 = g["U"], g["phi"], float(g["m0"])
    assert np.array_equal(P.D(U, phi, m0), g["D"])
    assert np.array_equal(P.D(U, phi, m0, True), g["Ddag"])
    assert np.array_equal(P.DDdag(U, phi, m0), g["DDdag"])
    z = P.dot(phi, g["chi"])
    assert (z.real, z.imag) == tuple(g["dot"])


def test_cg_bit_exact(case):
    P, g = case
    x, ok, apps, _ = P.cg(g["U"], g["phi"], float(g["m0"]))
    assert ok == int(g["cg_ok"]) == 1
    assert apps == int(g["cg_apps"])
    assert np.array_equal(x, g["cg_x"])


def test_forces_and_gauge_bit_exact(case):
    P, g = case
    U, m0, beta = g["U"], float(g["m0"]), float(g["beta"])
    x = g["cg_x"]
    assert np.array_equal(P.fermion_force(U, x, P.D(U, x, m0, True)), g["fforce"])
    assert np.array_equal(P.staple(U), g["staple"])
    pl, sp, sg = P.plaquette(U, beta)
    assert np.array_equal(pl, g["plaq"])
    assert (sp, sg) == tuple(g["plaq_sums"])
    F, ok = P.force(U, g["phi"], beta, m0)
    assert ok == 1 and np.array_equal(F, g["force"])


def test_hamiltonian_leapfrog_trajectory_bit_exact(case):
    P, g = case
    U, pi, phi, chi = g["U"], g["pi"], g["phi"], g["chi"]
    m0, beta, md, tau = float(g["m0"]), float(g["beta"]), int(g["md"]), float(g["tau"])
    assert P.action(U, phi, beta, m0) == float(g["action"])
    assert P.hamiltonian(U, pi, phi, beta, m0) == float(g["hamiltonian"])
    Ul, pl = P.leapfrog(U, pi, phi, md, tau, beta, m0)
    assert np.array_equal(Ul, g["lf_U"]) and np.array_equal(pl, g["lf_pi"])
    tr = P.trajectory(U, pi, chi, md, tau, beta, m0)
    assert np.array_equal(tr["phi"], g["tr_phi"])
    assert np.array_equal(tr["U"], g["tr_U"]) and np.array_equal(tr["pi"], g["tr_pi"])
    assert (tr["H_old"], tr["H_new"]) == tuple(g["tr_H"])
    assert (tr["sp"], tr["sg"]) == tuple(g["tr_aux"])


def test_64x64_fingerprints():
    s = np.load(os.path.join(GOLDEN, "ref_64x64_scalars.npz"))
    from oracle.port import gaussian_fields
    P = Port(64, 64)
    U = P.hot_start(12345)
    chi, pi = gaussian_fields(64, 64, 777)
    phi, _ = gaussian_fields(64, 64, 778)
    x, ok, apps, _ = P.cg(U, phi, 0.0)
    assert ok == 1 and apps == int(s["cg_apps"])
    assert np.linalg.norm(x) == float(s["cg_x_norm"])
    tr = P.trajectory(U, pi, chi, 10, 1.0, 2.0, 0.0)
    assert (tr["H_old"], tr["H_new"]) == tuple(s["tr_H"])


def test_config_file_bytes(tmp_path):
    g = load_golden(8, 8)
    P = Port(8, 8)
    f = tmp_path / "a.ctxt"
    P.save_conf(g["U"], str(f))
    want = open(os.path.join(GOLDEN, "ref_8x8.ctxt"), "rb").read()
    assert f.read_bytes() == want and len(want) == 2 * 64 * 28
    assert np.array_equal(P.read_binary(os.path.join(GOLDEN, "ref_8x8.ctxt")), g["U"])


# ---- known-answer physics checks the reference satisfies -----------------

def test_adjointness_and_gamma5(case):
    P, g = case
    U, phi, chi, m0 = g["U"], g["phi"], g["chi"], float(g["m0"])
    lhs = P.dot(P.D(U, phi, m0), chi)
    rhs = P.dot(phi, P.D(U, chi, m0, True))
    assert abs(lhs - rhs) < 1e-11 * abs(lhs)
    s3 = np.array([1.0, -1.0])[:, None]
    assert np.abs(s3 * P.D(U, s3 * phi, m0) - P.D(U, phi, m0, True)).max() < 1e-13


def test_free_field_symbol():
    # U == 1, plane wave with antiperiodic-compatible p_t: D has the 2x2 symbol of SURVEY section 4
    nx, nt, m0 = 8, 12, 0.3
    P = Port(nx, nt)
    U = np.ones((2, nx * nt), complex)
    kx, kt = 3, 2
    px, pt = 2 * np.pi * kx / nx, (2 * kt + 1) * np.pi / nt
    x, t = np.divmod(np.arange(nx * nt), nt)
    wave = np.exp(1j * (px * x + pt * t))
    v = np.array([0.3 - 0.2j, 1.1 + 0.7j])
    psi = v[:, None] * wave[None, :]
    s0 = np.array([[0, 1], [1, 0]], complex)
    s1 = np.array([[0, -1j], [1j, 0]], complex)
    sym = (m0 + 2 - np.cos(pt) - np.cos(px)) * np.eye(2) + 1j * (s0 * np.sin(pt) + s1 * np.sin(px))
    want = (sym @ v)[:, None] * wave[None, :]
    assert np.abs(P.D(U, psi, m0) - want).max() < 1e-13


def test_force_is_minus_dS(case):
    P, g = case
    U, phi, m0, beta = g["U"].copy(), g["phi"], float(g["m0"]), float(g["beta"])
    F, _ = P.force(U, phi, beta, m0, tol=1e-13)
    w = 1e-5
    for mu, n in [(0, 3), (1, 5), (0, U.shape[1] - 1)]:
        Up, Um = U.copy(), U.copy()
        Up[mu, n] *= np.exp(1j * w)
        Um[mu, n] *= np.exp(-1j * w)
        dS = (P.action(Up, phi, beta, m0, tol=1e-13) - P.action(Um, phi, beta, m0, tol=1e-13)) / (2 * w)
        assert abs(F[mu, n] + dS) < 2e-5 * max(1.0, abs(dS))


def test_leapfrog_reversible():
    g = load_golden(8, 8)
    P = Port(8, 8)
    U, pi, phi = g["U"], g["pi"], g["phi"]
    m0, beta = float(g["m0"]), float(g["beta"])
    U1, p1 = P.leapfrog(U, pi, phi, 6, 0.5, beta, m0, tol=1e-13)
    U2, p2 = P.leapfrog(U1, -p1, phi, 6, 0.5, beta, m0, tol=1e-13)
    assert np.abs(U2 - U).max() < 1e-11 and np.abs(p2 + pi).max() < 1e-10
    assert np.abs(np.abs(U1) - 1).max() < 1e-14


def test_gauge_covariance():
    """Under a local U(1) transformation g(n): U_mu(n) -> g(n) U_mu(n) conj(g(n + mu)), psi -> g psi the operators are
    covariant, D[U'] (g psi) = g D[U] psi (likewise D^dagger, D D^dagger; the antiperiodic sign of the fermions in t
    does not see g), the plaquette, the gauge action, the Hamiltonian and the force (a gauge-invariant real field) do
    not change, and the CG solution transforms like psi -- properties of the reference's operator (eq. 34-38 of its notes)
    that any index, sign or conjugation slip in a restatement breaks."""
    nx, nt, m0, beta = 6, 10, -0.04, 2.0
    P = Port(nx, nt)
    rng = np.random.default_rng(5)
    U = P.hot_start(77)
    psi, _ = gaussian_fields(nx, nt, 78)
    _, pi = gaussian_fields(nx, nt, 79)
    g = np.exp(1j * rng.uniform(0, 2 * np.pi, (nx, nt)))
    Ug = np.empty_like(U)
    Ug[0] = (g * U[0].reshape(nx, nt) * np.conj(np.roll(g, -1, axis=1))).ravel()     # mu = 0: time, n + mu = (x, t+1)
    Ug[1] = (g * U[1].reshape(nx, nt) * np.conj(np.roll(g, -1, axis=0))).ravel()     # mu = 1: space, (x+1, t)
    gf = g.ravel()[None, :]
    for dag in (False, True):
        assert np.abs(P.D(Ug, gf * psi, m0, dag) - gf * P.D(U, psi, m0, dag)).max() < 1e-13
    assert np.abs(P.DDdag(Ug, gf * psi, m0) - gf * P.DDdag(U, psi, m0)).max() < 1e-13
    p0, p1 = P.plaquette(U, beta), P.plaquette(Ug, beta)
    for a, b in zip(p0, p1):
        assert np.abs(np.asarray(a) - np.asarray(b)).max() < 1e-12
    x0, ok0, apps0, _ = P.cg(U, psi, m0, tol=1e-13)
    x1, ok1, apps1, _ = P.cg(Ug, gf * psi, m0, tol=1e-13)
    assert ok0 == ok1 == 1 and abs(apps0 - apps1) <= 1 and np.abs(x1 - gf * x0).max() < 1e-10 * np.abs(x0).max()
    H0 = P.hamiltonian(U, pi, psi, beta, m0, tol=1e-13)
    H1 = P.hamiltonian(Ug, pi, gf * psi, beta, m0, tol=1e-13)
    assert abs(H0 - H1) < 1e-9 * abs(H0)
    F0, _ = P.force(U, psi, beta, m0, tol=1e-13)
    F1, _ = P.force(Ug, gf * psi, beta, m0, tol=1e-13)
    assert np.abs(F0 - F1).max() < 1e-8 * max(1.0, np.abs(F0).max())
