"""GPU parity: libschwinger_b200.so (through the C ABI, host-mirror class Lattice) against
  (1) the golden vectors the UNMODIFIED reference produced (tests/golden/*.npz), and
  (2) the C oracle (oracle/schwinger_oracle.c, itself pinned bit-exactly to the reference)
on the same seeded inputs.  Tolerances are BASELINE.json's: tables bit-exact, one D / D^dagger
application <= 1e-13 relative, CG same flag / iterations within +-1 / solution to the stated
residual, dH <= 1e-8.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, relerr

pytestmark = pytest.mark.gpu

CASES = [(8, 8), (16, 24), (32, 32)]
TOL_D = 1e-13          # one stencil application, relative (north_star)
TOL_X = 1e-9           # CG solution, relative
TOL_DH = 1e-8          # per trajectory (north_star)


@pytest.fixture(scope="module")
def sb():
    import schwingermodel_b200 as m
    return m


@pytest.fixture(scope="module", params=CASES, ids=lambda c: f"{c[0]}x{c[1]}")
def case(request, sb):
    nx, nt = request.param
    lat = sb.Lattice(nx, nt)
    yield lat, load_golden(nx, nt)
    lat.close()


def test_tables_bit_exact(case):
    lat, g = case
    t = lat.periodic_boundary()
    for k, v in t.items():
        assert np.array_equal(v, g["tab_" + k]), k
    t22 = lat.periodic_boundary(2, 2, 3)
    flat = np.concatenate([v.view(np.float64).ravel() if v.dtype == np.complex128 else v.astype(np.float64)
                           for v in t22.values()])
    assert np.array_equal(flat, g["tab22_rank3"])


def test_tables_every_rank_vs_oracle(sb):
    from oracle.port import Port
    lat, P = sb.Lattice(16, 24), Port(16, 24)
    for rx, rt in [(2, 1), (1, 2), (2, 2), (4, 3), (1, 8)]:
        for rank in range(rx * rt):
            a, b = lat.periodic_boundary(rx, rt, rank), P.tables(rx, rt, rank)
            for k in a:
                assert np.array_equal(a[k], b[k]), (rx, rt, rank, k)
    lat.close()


def test_D_Ddag_DDdag_vs_reference(case):
    lat, g = case
    U, phi, m0 = g["U"], g["phi"], float(g["m0"])
    assert relerr(lat.D_phi(U, phi, m0), g["D"]) <= TOL_D
    assert relerr(lat.D_dagger_phi(U, phi, m0), g["Ddag"]) <= TOL_D
    assert relerr(lat.D_D_dagger_phi(U, phi, m0), g["DDdag"]) <= TOL_D


def test_dot_vs_reference(case):
    lat, g = case
    z = lat.dot(g["phi"], g["chi"])
    want = complex(*g["dot"])
    assert abs(z - want) <= 1e-13 * np.abs(g["phi"]).size


def test_cg_vs_reference(case):
    lat, g = case
    x, ok, its = lat.conjugate_gradient(g["U"], g["phi"], float(g["m0"]))
    assert ok == int(g["cg_ok"]) == 1
    assert abs((its + 2) - int(g["cg_apps"])) <= 1       # DD^dagger applications = k + 2
    assert relerr(x, g["cg_x"]) <= TOL_X
    # true residual to the stated tolerance
    r = g["phi"] - lat.D_D_dagger_phi(g["U"], x, float(g["m0"]))
    assert np.linalg.norm(r) <= 2e-10 * np.linalg.norm(g["phi"])


def test_forces_and_gauge_vs_reference(case):
    lat, g = case
    U, m0, beta = g["U"], float(g["m0"]), float(g["beta"])
    x = g["cg_x"]
    chi = lat.D_dagger_phi(U, x, m0)
    assert relerr(lat.phi_dag_partialD_phi(U, x, chi), g["fforce"]) <= 1e-12
    assert relerr(lat.Compute_Staple(U), g["staple"]) <= 1e-14
    P, sp, sg = lat.Compute_Plaquette01(U, beta)
    assert relerr(P, g["plaq"]) <= 1e-14
    assert abs(sp - g["plaq_sums"][0]) <= 1e-12 * U.shape[1]
    assert abs(sg - g["plaq_sums"][1]) <= 1e-12 * U.shape[1]
    lat.hmc_configure(beta, m0, int(g["md"]), float(g["tau"]))
    lat.hmc_set_gauge(U)
    F, ok = lat.hmc_force(g["phi"])
    assert ok == 1 and relerr(F, g["force"]) <= 1e-8


def test_hamiltonian_leapfrog_trajectory_vs_reference(case):
    lat, g = case
    U, pi, phi, chi = g["U"], g["pi"], g["phi"], g["chi"]
    m0, beta, md, tau = float(g["m0"]), float(g["beta"]), int(g["md"]), float(g["tau"])
    lat.hmc_configure(beta, m0, md, tau)
    lat.hmc_set_gauge(U)
    H = lat.hmc_hamiltonian(pi, phi)
    assert abs(H - float(g["hamiltonian"])) <= TOL_DH
    Ul, pl, ok = lat.hmc_leapfrog(pi, phi)
    assert ok == 1
    assert np.abs(Ul - g["lf_U"]).max() <= 1e-9 and np.abs(pl - g["lf_pi"]).max() <= 1e-8
    lat.hmc_inject(pi, chi)
    r = lat.hmc_trajectory()
    assert r.cg_all_converged == 1 and r.cg_solves == (md - 1) + 2
    assert relerr(lat.hmc_get_phi(), g["tr_phi"]) <= TOL_D
    assert np.abs(lat.hmc_get_gauge(True) - g["tr_U"]).max() <= 1e-9
    assert np.abs(lat.hmc_get_momenta(True) - g["tr_pi"]).max() <= 1e-8
    H_old, H_new = g["tr_H"]
    assert abs(r.H_old - H_old) <= TOL_DH and abs(r.H_new - H_new) <= TOL_DH
    assert abs(r.dH - (H_new - H_old)) <= TOL_DH
    assert abs(r.sum_re_plaq_new - g["tr_aux"][0]) <= 1e-9 and abs(r.gauge_action_new - g["tr_aux"][1]) <= 1e-9
    # accept swaps the proposal in; reject keeps U
    lat.hmc_accept(False)
    assert np.array_equal(lat.hmc_get_gauge(False), U)
    lat.hmc_accept(True)
    assert np.abs(lat.hmc_get_gauge(False) - g["tr_U"]).max() <= 1e-9


def test_64x64_config1_vs_reference_fingerprints(sb):
    """BASELINE config 1 size: CG on the hot start and one full trajectory (MD=10, tau=1, beta=2, m0=0)."""
    from oracle.port import Port, gaussian_fields
    s = np.load(os.path.join(GOLDEN, "ref_64x64_scalars.npz"))
    lat = sb.Lattice(64, 64)
    U = Port(64, 64).hot_start(12345)
    chi, pi = gaussian_fields(64, 64, 777)
    phi, _ = gaussian_fields(64, 64, 778)
    x, ok, its = lat.conjugate_gradient(U, phi, 0.0)
    assert ok == 1 and abs(its + 2 - int(s["cg_apps"])) <= 1
    assert abs(np.linalg.norm(x) - float(s["cg_x_norm"])) <= 1e-9 * float(s["cg_x_norm"])
    lat.hmc_configure(2.0, 0.0, 10, 1.0)
    lat.hmc_set_gauge(U)
    lat.hmc_inject(pi, chi)
    r = lat.hmc_trajectory()
    H_old, H_new = s["tr_H"]
    assert abs(r.H_old - H_old) <= TOL_DH and abs(r.H_new - H_new) <= TOL_DH
    assert abs(r.dH - (H_new - H_old)) <= TOL_DH
    assert abs(r.sum_re_plaq_new - float(s["tr_sp"])) <= 1e-8
    Un = lat.hmc_get_gauge(True)
    assert abs(Un.sum() - complex(*s["tr_U_sum"])) <= 1e-8
    lat.close()


@pytest.mark.parametrize("nx,nt,m0", [(4, 4, 0.2), (2, 2, 0.5), (6, 40, 0.0), (40, 6, -0.1), (64, 256, 0.0), (130, 70, 0.1),
                                      (256, 256, 0.0)])
def test_shapes_vs_oracle(sb, nx, nt, m0):
    """Tiny, ragged (partial tiles in both directions), non-square and config-2-sized lattices."""
    from oracle.port import Port, gaussian_fields
    P, lat = Port(nx, nt), sb.Lattice(nx, nt)
    U = P.hot_start(4321)
    phi, pi = gaussian_fields(nx, nt, 5)
    assert relerr(lat.D_phi(U, phi, m0), P.D(U, phi, m0)) <= TOL_D
    assert relerr(lat.D_dagger_phi(U, phi, m0), P.D(U, phi, m0, True)) <= TOL_D
    assert relerr(lat.D_D_dagger_phi(U, phi, m0), P.DDdag(U, phi, m0)) <= TOL_D
    z, w = lat.dot(phi, U), P.dot(phi, U)
    assert abs(z - w) <= 1e-12 * nx * nt
    x = P.cg(U, phi, m0)[0] if nx * nt <= 70000 else phi
    chi = P.D(U, x, m0, True)
    assert relerr(lat.phi_dag_partialD_phi(U, x, chi), P.fermion_force(U, x, chi)) <= 1e-12
    assert relerr(lat.Compute_Staple(U), P.staple(U)) <= 1e-14
    Pg, sp, sg = lat.Compute_Plaquette01(U, 2.0)
    Po, spo, sgo = P.plaquette(U, 2.0)
    assert relerr(Pg, Po) <= 1e-14 and abs(sp - spo) <= 1e-12 * nx * nt and abs(sg - sgo) <= 1e-12 * nx * nt
    lat.close()


def test_config2_cg_256_vs_oracle(sb):
    """BASELINE config 2: one CG solve on the 256x256 hot start, m0 = 0."""
    from oracle.port import Port, gaussian_fields
    P, lat = Port(256, 256), sb.Lattice(256, 256)
    U = P.hot_start(12345)
    phi, _ = gaussian_fields(256, 256, 777)
    xo, oko, apps, _ = P.cg(U, phi, 0.0)
    x, ok, its = lat.conjugate_gradient(U, phi, 0.0)
    assert ok == oko == 1 and abs(its + 2 - apps) <= 1
    assert relerr(x, xo) <= TOL_X
    lat.close()


def test_cg_not_converged_returns_zero(sb):
    from oracle.port import Port, gaussian_fields
    P, lat = Port(16, 16), sb.Lattice(16, 16)
    U = P.hot_start(1)
    phi, _ = gaussian_fields(16, 16, 2)
    for max_iter in (1, 5, 8, 9, 17):
        lat.set_cg(1e-10, max_iter)
        x, ok, its = lat.conjugate_gradient(U, phi, 0.0)
        xo, oko, apps, _ = P.cg(U, phi, 0.0, 1e-10, max_iter)
        assert ok == oko == 0 and its == max_iter and apps == max_iter + 1
        assert relerr(x, xo) <= 1e-10
    lat.set_cg(1e-3, 100)   # loose tolerance: stops early, same count as the oracle
    x, ok, its = lat.conjugate_gradient(U, phi, 0.0)
    xo, oko, apps, _ = P.cg(U, phi, 0.0, 1e-3, 100)
    assert ok == oko == 1 and its + 2 == apps
    lat.close()


def test_device_resident_path_matches_host_path(sb):
    from oracle.port import Port, gaussian_fields
    nx, nt, m0 = 32, 48, -0.02
    P, lat = Port(nx, nt), sb.Lattice(nx, nt)
    U = P.hot_start(9)
    phi, _ = gaussian_fields(nx, nt, 10)
    dU, dphi, dout, dx = (lat.new_field(True, U), lat.new_field(True, phi), lat.new_field(), lat.new_field())
    lat.dev_D(dU, dphi, dout, m0)
    assert np.array_equal(dout.download(), lat.D_phi(U, phi, m0))
    lat.dev_DDdag(dU, dphi, dout, m0)
    assert np.array_equal(dout.download(), lat.D_D_dagger_phi(U, phi, m0))
    ok, its = lat.dev_cg(dU, dphi, dx, m0)
    x, ok2, its2 = lat.conjugate_gradient(U, phi, m0)
    assert (ok, its) == (ok2, its2) and np.array_equal(dx.download(), x)
    assert lat.dev_dot(dx, dphi) == lat.dot(x, phi)
    ms = lat.dev_DDdag_loop(dU, dphi, dout, m0, 5)
    assert ms > 0 and np.array_equal(dout.download(), lat.D_D_dagger_phi(U, phi, m0))
    assert lat.launch_count() > 0
    for f in (dU, dphi, dout, dx):
        f.free()
    lat.close()


# ---- size-independent properties (also run at larger sizes) ------------------------------------

@pytest.mark.parametrize("nx,nt", [(32, 32), (512, 512), (1024, 2048)])
def test_adjointness_gamma5_linearity(sb, nx, nt):
    from oracle.port import gaussian_fields
    lat = sb.Lattice(nx, nt)
    rng = np.random.default_rng(3)
    U = np.exp(2j * np.pi * rng.random((2, nx * nt)))
    a, _ = gaussian_fields(nx, nt, 11)
    b, _ = gaussian_fields(nx, nt, 12)
    m0 = -0.05
    Da, Ddb = lat.D_phi(U, a, m0), lat.D_dagger_phi(U, b, m0)
    lhs, rhs = lat.dot(Da, b), lat.dot(a, Ddb)
    assert abs(lhs - rhs) <= 1e-12 * abs(lhs) + 1e-9
    s3 = np.array([1.0, -1.0])[:, None]
    assert relerr(s3 * lat.D_phi(U, s3 * a, m0), lat.D_dagger_phi(U, a, m0)) <= TOL_D
    lin = lat.D_phi(U, 2.0 * a + (0.5 - 1j) * b, m0)
    assert relerr(lin, 2.0 * Da + (0.5 - 1j) * lat.D_phi(U, b, m0)) <= 5e-13
    # CG really inverts D D^dagger
    lat.set_cg(1e-10, 10000)
    x, ok, its = lat.conjugate_gradient(U, a, 0.1)
    assert ok == 1
    r = a - lat.D_D_dagger_phi(U, x, 0.1)
    assert np.linalg.norm(r) <= 2e-10 * np.linalg.norm(a)
    lat.close()


def _cheap_fields(nx, nt, block=64):
    """Full-size inputs in a few seconds: a random block of `block` rows repeated along x, every row turned by its own
    random phase (so no two rows are equal and nothing is periodic in x).  Links have unit modulus, the source is
    Gaussian."""
    assert nx % block == 0
    rng = np.random.default_rng(99)
    reps = nx // block
    turn = np.exp(2j * np.pi * rng.random(nx)).reshape(reps, block, 1)
    fields = []
    for gaussian in (False, True):
        f = np.empty((2, nx * nt), np.complex128)
        for mu in range(2):
            if gaussian:
                base = rng.standard_normal((block, nt)) + 1j * rng.standard_normal((block, nt))
            else:
                base = np.exp(2j * np.pi * rng.random((block, nt)))
            np.multiply(base[None, :, :], turn, out=f[mu].reshape(reps, block, nt))
        fields.append(f)
    return fields[0], fields[1]


def test_full_size_8192_properties(sb):
    """BASELINE configs[3]'s lattice itself (8192 x 8192, 2 GiB per field), device-resident:
      * the one-pass D D^dagger equals the two stencil passes everywhere (every strip and chunk boundary of the
        temporally blocked kernel), and both equal the oracle on bands of rows -- a band of the lattice with two extra
        rows on each side is an exact sub-problem for its interior rows (D D^dagger reaches two rows; t is kept whole,
        so the antiperiodic seam is the lattice's own);
      * <D a, b> = <a, D^dagger b>;
      * the CG solution's TRUE residual meets the reference's stopping rule."""
    import psutil
    if psutil.virtual_memory().available < 24 * 2 ** 30:
        pytest.skip("needs ~16 GiB of host memory")
    from oracle.port import Port
    n, m0 = 8192, 0.0
    V = n * n
    U, phi = _cheap_fields(n, n)
    one = sb.Lattice(n, n)
    assert one.one_pass_dd()
    dU, dphi, dout = one.new_field(True, U), one.new_field(True, phi), one.new_field(True)
    one.dev_DDdag(dU, dphi, dout, m0)
    got = dout.download()
    # two stencil passes on the same fields (dev_D never takes the one-pass kernel)
    dtmp = one.new_field(True)
    one.dev_D(dU, dphi, dtmp, m0, True)
    one.dev_D(dU, dtmp, dout, m0, False)
    two = dout.download()
    scale = np.abs(two[:, :4 * n]).max()
    assert np.abs(got - two).max() <= 1e-14 * scale * 10
    del two
    # bands against the oracle: rows [x0, x0 + R) with two extra rows on each side, wrapped in x
    R = 60
    for x0 in (0, 1000, 4093, n - R):
        rows = np.arange(x0 - 2, x0 + R + 2) % n
        idx = (rows[:, None] * n + np.arange(n)[None, :]).ravel()
        P = Port(R + 4, n)
        want = P.DDdag(np.ascontiguousarray(U[:, idx]), np.ascontiguousarray(phi[:, idx]), m0)
        inner = slice(2 * n, (R + 2) * n)
        assert relerr(got[:, idx][:, inner], want[:, inner]) <= 2 * TOL_D, x0
    del got
    # adjointness at full size: a = phi, b = a second field (the links serve: unit modulus, uncorrelated with phi)
    db = dU
    one.dev_D(dU, dphi, dtmp, m0, False)              # D a
    lhs = one.dev_dot(dtmp, db)
    bound = np.sqrt(one.dev_dot(dtmp, dtmp).real * one.dev_dot(db, db).real)    # |<D a, b>| <= |D a| |b|
    one.dev_D(dU, db, dtmp, m0, True)                 # D^dagger b
    rhs = one.dev_dot(dphi, dtmp)
    assert abs(lhs - rhs) <= 1e-12 * bound, (lhs, rhs, bound)
    # CG: converged flag and true residual
    dx = dtmp
    ok, its = one.dev_cg(dU, dphi, dx, m0)
    assert ok == 1 and 50 < its < 400
    one.dev_DDdag(dU, dx, dout, m0)
    res = phi - dout.download()
    assert np.linalg.norm(res.ravel()) <= 3e-10 * np.linalg.norm(phi.ravel())
    one.close()


def test_full_size_8192_reference_fingerprint(sb):
    """BASELINE configs[3]'s lattice against numbers the UNMODIFIED reference produced at that size
    (tests/golden/ref_8192x8192_fingerprint.npz, made by tests/golden/make_golden.py 8192 over 8 forked ranks):
    hot start srand(12345), Gaussian source, m0 = 0 -> D D^dagger phi and the CG solution at 256 sites (random ones,
    lattice corners, the antiperiodic seam, strip and chunk edges of the one-pass kernel), their norms and the CG's
    application count (src/dirac_operator.cpp:477-480, src/conjugate_gradient.cpp:4-67)."""
    import psutil
    if psutil.virtual_memory().available < 24 * 2 ** 30:
        pytest.skip("needs ~16 GiB of host memory")
    from oracle.port import Port, gaussian_fields
    fp = np.load(os.path.join(GOLDEN, "ref_8192x8192_fingerprint.npz"))
    n, m0, sites = 8192, float(fp["m0"]), fp["sites"]
    U = Port(n, n).hot_start(int(fp["hot_start_seed"]))
    phi = gaussian_fields(n, n, int(fp["source_seed"]))[0]
    assert np.array_equal(U[:, sites], fp["U_s"]) and np.array_equal(phi[:, sites], fp["phi_s"])   # same inputs
    lat = sb.Lattice(n, n)
    assert lat.one_pass_dd()
    dU, dphi, dout = lat.new_field(True, U), lat.new_field(True, phi), lat.new_field(True)
    del U
    lat.dev_DDdag(dU, dphi, dout, m0)
    got = dout.download()
    assert relerr(got[:, sites], fp["dd_s"]) <= TOL_D
    assert abs(np.linalg.norm(got.ravel()) - float(fp["dd_norm"])) <= 1e-12 * float(fp["dd_norm"])
    del got
    ok, its = lat.dev_cg(dU, dphi, dout, m0)
    assert ok == 1 and abs(its + 2 - int(fp["cg_apps"])) <= 1, (its, int(fp["cg_apps"]))
    x = dout.download()
    assert relerr(x[:, sites], fp["cg_x_s"]) <= TOL_X
    assert abs(np.linalg.norm(x.ravel()) - float(fp["cg_x_norm"])) <= 1e-9 * float(fp["cg_x_norm"])
    assert abs(x.sum() - complex(*fp["cg_x_sum"])) <= 1e-9 * float(fp["cg_x_norm"]) * np.sqrt(x.size)
    lat.close()


def test_free_field_symbol(sb):
    nx, nt, m0 = 8, 12, 0.3
    lat = sb.Lattice(nx, nt)
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
    assert np.abs(lat.D_phi(U, psi, m0) - want).max() < 1e-13
    lat.close()


def test_force_is_minus_dS_and_leapfrog_reversible(sb):
    g = load_golden(8, 8)
    lat = sb.Lattice(8, 8)
    U, pi, phi = g["U"].copy(), g["pi"], g["phi"]
    m0, beta = float(g["m0"]), float(g["beta"])
    lat.set_cg(1e-13, 10000)
    lat.hmc_configure(beta, m0, 6, 0.5)
    lat.hmc_set_gauge(U)
    F, _ = lat.hmc_force(phi)
    zero = np.zeros_like(pi)
    w = 1e-5
    for mu, n in [(0, 3), (1, 5), (0, 63)]:
        Up, Um = U.copy(), U.copy()
        Up[mu, n] *= np.exp(1j * w)
        Um[mu, n] *= np.exp(-1j * w)
        lat.hmc_set_gauge(Up)
        sp = lat.hmc_hamiltonian(zero, phi)
        lat.hmc_set_gauge(Um)
        sm = lat.hmc_hamiltonian(zero, phi)
        dS = (sp - sm) / (2 * w)
        assert abs(F[mu, n] + dS) < 2e-5 * max(1.0, abs(dS))
    lat.hmc_set_gauge(U)
    U1, p1, _ = lat.hmc_leapfrog(pi, phi)
    lat.hmc_set_gauge(U1)
    U2, p2, _ = lat.hmc_leapfrog(-p1, phi)
    assert np.abs(U2 - U).max() < 1e-11 and np.abs(p2 + pi).max() < 1e-10
    assert np.abs(np.abs(U1) - 1).max() < 1e-14
    lat.close()


def test_device_rng_moments_and_decomposition_independence(sb):
    """HMC::RandomPI / RandomCHI on the device (src/hmc.cpp:5-28): pi ~ N(0,1) per link, Re/Im chi ~ N(0, 1/sqrt 2) per
    spin component.  Mean, variance, 4th moment and tails of both, independence of the two values that come out of
    one Box-Muller pair (pi0(n), pi1(n); Re, Im of a chi component) and of neighbouring sites, replay determinism."""
    n = 512
    V = n * n
    lat = sb.Lattice(n, n)
    lat.hmc_configure(2.0, 0.0, 4, 0.1)
    lat.hmc_refresh(42, 0)
    pi, chi = lat.hmc_get_momenta(False), lat.hmc_get_chi()
    N = 2 * V                                           # samples per real field
    se = 1.0 / np.sqrt(N)

    def gaussian_checks(v, var, what):
        v = v.ravel()
        sd = np.sqrt(var)
        assert abs(v.mean()) < 5 * sd * se, what
        assert abs(v.var() - var) < 5 * var * np.sqrt(2.0) * se, what
        assert abs((v ** 4).mean() / var ** 2 - 3.0) < 5 * np.sqrt(96.0) * se, what          # kurtosis 3
        assert abs((v ** 3).mean()) < 5 * sd ** 3 * np.sqrt(15.0) * se, what                   # no skew
        for k, p in ((2.0, 0.04550026), (3.0, 2.699796e-3), (4.0, 6.334248e-5)):               # two-sided tails
            frac = np.mean(np.abs(v) > k * sd)
            assert abs(frac - p) < 5 * np.sqrt(p / v.size) + 1e-7, (what, k, frac)
        assert np.abs(v).max() < 7.0 * sd, what

    gaussian_checks(pi, 1.0, "pi")
    gaussian_checks(chi.real, 0.5, "Re chi")
    gaussian_checks(chi.imag, 0.5, "Im chi")

    def corr(a, b):
        a, b = a.ravel() - a.mean(), b.ravel() - b.mean()
        return float((a * b).mean() / np.sqrt((a * a).mean() * (b * b).mean()))

    lim = 5.0 / np.sqrt(V)
    assert abs(corr(pi[0], pi[1])) < lim                       # the two outputs of one Philox block / Box-Muller pair
    assert abs(corr(pi[0] ** 2, pi[1] ** 2)) < lim             # ... also in their radii
    assert abs(corr(chi[0].real, chi[0].imag)) < lim and abs(corr(chi[1].real, chi[1].imag)) < lim
    assert abs(corr(chi[0].real, chi[1].real)) < lim and abs(corr(chi[0].real, pi[0])) < lim
    assert abs(corr(pi[0][:-1], pi[0][1:])) < lim              # neighbouring sites (consecutive counters)
    assert abs(corr(pi[0][:-n], pi[0][n:])) < lim
    lat.hmc_refresh(42, 1)
    pi2 = lat.hmc_get_momenta(False)
    assert not np.array_equal(pi, pi2) and abs(corr(pi, pi2)) < lim     # next trajectory: fresh, uncorrelated
    lat.hmc_refresh(43, 0)
    assert abs(corr(pi, lat.hmc_get_momenta(False))) < lim              # another seed
    lat.hmc_refresh(42, 0)
    assert np.array_equal(pi, lat.hmc_get_momenta(False)) and np.array_equal(chi, lat.hmc_get_chi())
    lat.close()


def test_config_file_bytes_and_roundtrip(sb, tmp_path):
    g = load_golden(8, 8)
    f = tmp_path / "a.ctxt"
    sb.SaveConf(g["U"], 8, 8, str(f))
    want = open(os.path.join(GOLDEN, "ref_8x8.ctxt"), "rb").read()
    assert f.read_bytes() == want
    assert np.array_equal(sb.readBinary(8, 8, os.path.join(GOLDEN, "ref_8x8.ctxt")), g["U"])
    with pytest.raises(sb.SchwingerError):
        sb.readBinary(16, 16, str(f))      # short file


def test_hmc_chain_statistics_64(sb):
    """Physics smoke of SURVEY section 4: 64x64, beta=2, m0=0, MD=10, tau=1 -> <P> ~ 0.7187(11), acceptance ~0.65."""
    from oracle.port import Port
    lat = sb.Lattice(64, 64)
    U = Port(64, 64).hot_start(12345)
    h = sb.HMC(lat, U, 10, 1.0, 60, 20, 0, 2.0, 0.0, seed=7)
    Ep, dEp = h.HMC_algorithm()
    assert abs(Ep - 0.7187) < 0.02
    assert 0.3 < h.getacceptance_rate() <= 1.0
    assert all(ok for (_, _, _, ok, _) in h.history)
    lat.close()


def test_one_pass_and_two_pass_DDdag_agree(sb):
    """The temporally blocked D D^dagger (sm_fused.cuh) against D^dagger-then-D (k_wilson twice), and the
    CG built on each, on ragged strips/chunks."""
    from oracle.port import Port, gaussian_fields
    for nx, nt, m0 in [(8, 8, 0.1), (37, 300, -0.02), (300, 37, 0.0), (64, 64, 0.0), (257, 509, 0.05)]:
        P = Port(nx, nt)
        U = P.hot_start(77)
        phi, _ = gaussian_fields(nx, nt, 78)
        os.environ["SM_DD_PATH"] = "twopass"
        two = sb.Lattice(nx, nt)
        os.environ["SM_DD_PATH"] = "onepass"
        one = sb.Lattice(nx, nt)
        os.environ.pop("SM_DD_PATH")
        a, b = one.D_D_dagger_phi(U, phi, m0), two.D_D_dagger_phi(U, phi, m0)
        assert relerr(a, b) <= 1e-14, (nx, nt)
        assert relerr(a, P.DDdag(U, phi, m0)) <= TOL_D
        xa, oka, ia = one.conjugate_gradient(U, phi, m0)
        xb, okb, ib = two.conjugate_gradient(U, phi, m0)
        assert oka == okb == 1 and abs(ia - ib) <= 1
        assert relerr(xa, xb) <= TOL_X
        for rows in (1, 3):      # extreme chunking: every row is a warm-up row of some block
            os.environ.update(SM_FUSED_ROWS=str(rows), SM_DD_PATH="onepass", SM_FUSED_BT="256" if rows == 3 else "128")
            tiny = sb.Lattice(nx, nt)
            for k in ("SM_FUSED_ROWS", "SM_DD_PATH", "SM_FUSED_BT"):
                os.environ.pop(k)
            assert relerr(tiny.D_D_dagger_phi(U, phi, m0), b) <= 1e-14
            tiny.close()
        one.close()
        two.close()


def test_cg_execution_strategies_agree(sb):
    """Cluster-resident CG (<= 4096 sites), CUDA-graph batches and plain launches run the same algorithm."""
    from oracle.port import Port, gaussian_fields
    for nx, nt, m0 in [(64, 64, 0.0), (16, 24, -0.05), (96, 100, 0.02), (5, 7, 0.3)]:   # cluster, cluster, grid, cluster
        P = Port(nx, nt)
        U = P.hot_start(5)
        phi, _ = gaussian_fields(nx, nt, 6)
        xo, oko, apps, _ = P.cg(U, phi, m0)
        got, full = {}, {}
        for name, env in [("cluster", {}), ("graphs", {"SM_CLUSTER_CG": "0", "SM_PDL": "0"}),
                          ("launches", {"SM_CLUSTER_CG": "0", "SM_GRAPHS": "0", "SM_PDL": "0"}),
                          # programmatic dependent launch of the two kernels of an iteration: the same arithmetic
                          ("graphs pdl", {"SM_CLUSTER_CG": "0", "SM_PDL": "1"}),
                          ("launches pdl", {"SM_CLUSTER_CG": "0", "SM_GRAPHS": "0", "SM_PDL": "1"}),
                          ("twopass", {"SM_CLUSTER_CG": "0", "SM_DD_PATH": "twopass"})]:
            os.environ.update(env)
            lat = sb.Lattice(nx, nt)
            for k in env:
                os.environ.pop(k)
            x, ok, its = lat.conjugate_gradient(U, phi, m0)
            x2, ok2, its2 = lat.conjugate_gradient(U, phi, m0)        # second solve replays the cached graph
            assert (ok, its) == (ok2, its2) and np.array_equal(x, x2), name
            assert ok == oko == 1 and abs(its + 2 - apps) <= 1, (name, its, apps)
            assert relerr(x, xo) <= TOL_X, name
            full[name] = x
            lat.set_cg(1e-10, 11)                                      # stop mid-way: same iterate everywhere
            xm, okm, itm = lat.conjugate_gradient(U, phi, m0)
            assert okm == 0 and itm == 11
            got[name] = xm
            lat.close()
        xref = P.cg(U, phi, m0, 1e-10, 11)[0]
        for name, xm in got.items():
            assert relerr(xm, xref) <= 1e-10, name
        assert np.array_equal(full["graphs pdl"], full["graphs"]) and np.array_equal(full["launches pdl"], full["graphs"])
        assert np.array_equal(got["graphs pdl"], got["graphs"]) and np.array_equal(got["launches pdl"], got["graphs"])


@pytest.mark.parametrize("nx,nt,m0", [(288, 288, 0.0), (300, 333, -0.03), (400, 401, 0.05), (512, 512, -0.1), (1100, 275, 0.0)])
def test_grid_resident_cg_several_sites_per_thread(sb, nx, nt, m0):
    """Lattices beyond one site per thread of a full cooperative grid (k_cg_cols: column segments of up to 8 rows per
    thread; ragged sizes leave empty slots; 512 x 512 is BASELINE configs[4]'s lattice): same iterate as the oracle and
    as the CUDA-graph path, also when stopped mid-way."""
    from oracle.port import Port, gaussian_fields
    P = Port(nx, nt)
    U = P.hot_start(8)
    phi, _ = gaussian_fields(nx, nt, 9)
    xo, oko, apps, _ = P.cg(U, phi, m0)
    xref = P.cg(U, phi, m0, 1e-10, 7)[0]
    for name, env in [("columns", {}), ("columns 4x512", {"SM_COLS": "4,512"}), ("graphs", {"SM_COLS": "0"})]:
        os.environ.update(env)
        lat = sb.Lattice(nx, nt)
        for k in env:
            os.environ.pop(k)
        x, ok, its = lat.conjugate_gradient(U, phi, m0)
        assert ok == oko == 1 and abs(its + 2 - apps) <= 1, (name, its, apps)
        assert relerr(x, xo) <= TOL_X, name
        res = np.linalg.norm(phi - lat.D_D_dagger_phi(U, x, m0)) / np.linalg.norm(phi)
        assert res <= 2e-10, (name, res)
        lat.set_cg(1e-10, 7)
        xm, okm, itm = lat.conjugate_gradient(U, phi, m0)
        assert okm == 0 and itm == 7 and relerr(xm, xref) <= 1e-10, name
        lat.close()


@pytest.mark.parametrize("nx,nt,m0", [(96, 100, 0.02), (70, 33, -0.05), (131, 32, 0.1), (37, 45, 0.0), (66, 64, 0.0),
                                      (256, 256, 0.0)])
def test_column_segment_cg_every_segment_length(sb, nx, nt, m0):
    """k_cg_cols forced with 1 ... 8 rows per thread (512- and 256-thread CTAs) on lattices the one-site kernels would
    take: rows that do not divide by the segment length, row wraps in the middle of a warp (width_t not a multiple of
    32), one-row groups."""
    from oracle.port import Port, gaussian_fields
    P = Port(nx, nt)
    U = P.hot_start(18)
    phi, _ = gaussian_fields(nx, nt, 19)
    xo, oko, apps, _ = P.cg(U, phi, m0)
    xref = P.cg(U, phi, m0, 1e-10, 9)[0]
    for s in ("1,512", "2,512", "3,512", "4,512", "2,256", "3,256", "5,256", "8,256"):
        os.environ["SM_COLS"] = s
        lat = sb.Lattice(nx, nt)
        os.environ.pop("SM_COLS")
        x, ok, its = lat.conjugate_gradient(U, phi, m0)
        assert ok == oko == 1 and abs(its + 2 - apps) <= 1, (s, its, apps)
        assert relerr(x, xo) <= TOL_X, s
        lat.set_cg(1e-10, 9)
        xm, okm, itm = lat.conjugate_gradient(U, phi, m0)
        assert okm == 0 and itm == 9 and relerr(xm, xref) <= 1e-10, s
        lat.close()


def test_config2_trajectory_256_vs_oracle(sb):
    """A full trajectory at BASELINE config 2's size (grid-resident CG path): dH against the oracle."""
    from oracle.port import Port, gaussian_fields
    n = 256
    P, lat = Port(n, n), sb.Lattice(n, n)
    U = P.hot_start(12345)
    chi, pi = gaussian_fields(n, n, 777)
    md, tau, beta, m0 = 4, 0.2, 2.0, 0.0
    t = P.trajectory(U, pi, chi, md, tau, beta, m0)
    lat.hmc_configure(beta, m0, md, tau)
    lat.hmc_set_gauge(U)
    lat.hmc_inject(pi, chi)
    r = lat.hmc_trajectory()
    assert r.cg_all_converged == 1 and t["cg_ok"] == 1
    assert abs(r.dd_applications - t["dd_apps"]) <= r.cg_solves
    assert abs(r.dH - t["dH"]) <= TOL_DH, (r.dH, t["dH"])
    assert np.abs(lat.hmc_get_gauge(True) - t["U"]).max() <= 1e-9
    lat.close()


def test_config5_near_critical_cg_vs_oracle(sb):
    """BASELINE config 5 parameters (m0 = -0.18, ill-conditioned, long CG) at 128x128: same convergence flag,
    iteration count within +-1, solution to 1e-9, true residual to the stated tolerance."""
    from oracle.port import Port, gaussian_fields
    n, m0 = 128, -0.18
    P, lat = Port(n, n), sb.Lattice(n, n)
    U = P.hot_start(2024)
    phi, _ = gaussian_fields(n, n, 31)
    xo, oko, apps, _ = P.cg(U, phi, m0)
    x, ok, its = lat.conjugate_gradient(U, phi, m0)
    assert ok == oko == 1
    assert abs(its + 2 - apps) <= 1, (its, apps)
    res = np.linalg.norm(phi - lat.D_D_dagger_phi(U, x, m0)) / np.linalg.norm(phi)
    assert res <= 2e-10
    assert relerr(x, xo) <= TOL_X
    lat.close()


# (lattice, beta, m0, MD, tau): BASELINE configs[4]'s parameters (near critical, MD = 20) on 64^2 and 128^2 and
# configs[2]'s (beta = 4, m0 = -0.05, MD = 10) on 256^2 -- the sizes at which H (~ 5 per site) still resolves 1e-8
# in double precision; CG::tol = 1e-10 (the reference's) and 1e-14 on both sides (hard part 1)
TRAJ_CASES = [(64, 2.0, -0.18, 20, 1.0), (128, 2.0, -0.18, 20, 1.0), (256, 4.0, -0.05, 10, 1.0)]


@pytest.mark.parametrize("tol", [1e-10, 1e-14], ids=["tol1e-10", "tol1e-14"])
@pytest.mark.parametrize("n,beta,m0,md,tau", TRAJ_CASES, ids=["cfg4_64", "cfg4_128", "cfg2_256"])
def test_trajectory_dH_at_baseline_parameters(sb, n, beta, m0, md, tau, tol):
    """One whole HMC_Update (src/hmc.cpp:151-181) at the parameters of BASELINE configs[2] and configs[4]:
    dH, both Hamiltonians, the proposal U' and pi', and the count of D D^dagger applications against the oracle."""
    from oracle.port import Port, gaussian_fields
    P, lat = Port(n, n), sb.Lattice(n, n)
    U = P.hot_start(12345)
    chi, pi = gaussian_fields(n, n, 777)
    t = P.trajectory(U, pi, chi, md, tau, beta, m0, tol=tol)
    lat.set_cg(tol, 10000)
    lat.hmc_configure(beta, m0, md, tau)
    lat.hmc_set_gauge(U)
    lat.hmc_inject(pi, chi)
    r = lat.hmc_trajectory()
    assert r.cg_all_converged == 1 and t["cg_ok"] == 1 and r.cg_solves == (md - 1) + 2
    assert abs(r.dd_applications - t["dd_apps"]) <= r.cg_solves, (r.dd_applications, t["dd_apps"])
    assert abs(r.H_old - t["H_old"]) <= TOL_DH and abs(r.H_new - t["H_new"]) <= TOL_DH, (r.H_old - t["H_old"], r.H_new - t["H_new"])
    assert abs(r.dH - t["dH"]) <= TOL_DH, (r.dH, t["dH"])
    assert np.abs(lat.hmc_get_gauge(True) - t["U"]).max() <= 1e-9
    assert np.abs(lat.hmc_get_momenta(True) - t["pi"]).max() <= 1e-8
    lat.close()


def test_mixed_precision_solver_opt_in(sb):
    """single-precision inner CG inside a double-precision defect correction.  Same stopping rule on the
    TRUE residual; the iterate differs from the reference's, so the bar is the residual and a loose solution match."""
    from oracle.port import gaussian_fields
    for n, m0 in [(384, 0.0), (512, -0.05)]:
        rng = np.random.default_rng(n)
        U = np.exp(2j * np.pi * rng.random((2, n * n)))
        phi, _ = gaussian_fields(n, n, 3)
        lat = sb.Lattice(n, n)
        x64, ok64, its64 = lat.conjugate_gradient(U, phi, m0)
        lat.set_solver(True)
        x32, ok32, its32 = lat.conjugate_gradient(U, phi, m0)
        assert ok64 == ok32 == 1
        res = np.linalg.norm(phi - lat.D_D_dagger_phi(U, x32, m0)) / np.linalg.norm(phi)
        assert res < 1e-10                                   # the reference's criterion, on the true residual
        assert relerr(x32, x64) <= 1e-7
        assert its32 <= 2 * its64 + 20
        # a trajectory with the opt-in solver: dH agrees with the reference solver far below Metropolis relevance
        chi, pi = gaussian_fields(n, n, 4)
        out = []
        for mixed in (False, True):
            lat.set_solver(mixed)
            lat.hmc_configure(2.0, m0, 4, 0.2)
            lat.hmc_set_gauge(U)
            lat.hmc_inject(pi, chi)
            out.append(lat.hmc_trajectory())
        assert out[1].cg_all_converged == 1
        assert abs(out[0].dH - out[1].dH) <= 1e-5 * max(1.0, abs(out[0].H_old) * 1e-3)
        lat.close()


@pytest.mark.parametrize("n", [64, 160, 512], ids=["cluster", "grid-resident", "columns"])
def test_chronological_start_vector_opt_in(sb, n):
    """SURVEY 8f.4, opt-in: inside a trajectory the force solves start from the previous solution (second solve) or
    from the extrapolation of the last two instead of from phi.  Same CG, same stopping rule: every solve converges,
    fewer D D^dagger applications (a modest saving: CG's count depends on the logarithm of the start residual), the proposal and dH equal the reference solver's far below Metropolis relevance
    (not to 1e-8: a different iterate), and a solve started from a vector of the caller's meets the stopping rule on
    the TRUE residual.  BASELINE configs[4] parameters (m0 = -0.18, MD = 20), on every resident CG kernel."""
    from oracle.port import Port, gaussian_fields
    beta, m0, md, tau = 2.0, -0.18, 20, 1.0
    U = Port(n, n).hot_start(12345)
    chi, pi = gaussian_fields(n, n, 777)
    lat = sb.Lattice(n, n)
    out = []
    for solver in ("reference", "chrono"):
        lat.set_solver(solver)
        lat.hmc_configure(beta, m0, md, tau)
        lat.hmc_set_gauge(U)
        lat.hmc_inject(pi, chi)
        r = lat.hmc_trajectory()
        out.append((r.dH, r.H_old, r.dd_applications, r.cg_all_converged, lat.hmc_get_gauge(True), lat.hmc_get_momenta(True)))
    lat.set_solver("reference")
    (dH0, H0, apps0, ok0, U0, p0), (dH1, H1, apps1, ok1, U1, p1) = out
    assert ok0 == 1 and ok1 == 1
    assert apps1 < 0.97 * apps0, (apps0, apps1)
    assert abs(H0 - H1) <= 1e-8                       # the old configuration's solve starts from phi in both
    assert abs(dH0 - dH1) <= 1e-5 * max(1.0, abs(H0) * 1e-3), (dH0, dH1)
    assert np.abs(U0 - U1).max() <= 1e-7 and np.abs(p0 - p1).max() <= 1e-6
    lat.close()


def _parity_masks(nx, nt):
    x, t = np.divmod(np.arange(nx * nt), nt)
    even = ((x + t) & 1) == 0
    return even, ~even


@pytest.mark.parametrize("nx,nt,m0", [(64, 48, 0.0), (96, 96, -0.18), (256, 256, -0.1)])
def test_evenodd_solver_opt_in(sb, nx, nt, m0):
    """SURVEY 8f.4, opt-in: CG on the Schur complement of D on the even sites.  The operator is rebuilt here from the
    plain stencil (D_phi / D_dagger_phi of the C ABI and parity masks), independently of the even-odd kernels:
    Dhat v = m v - (1/m) [D ([odd] D v)]_even.  The solution meets the reference's stopping rule on the TRUE residual of
    Dhat Dhat^dagger x = phi_e, vanishes on the odd sites, equals the even part of the full-lattice solve that the
    Schur complement stands for, and takes well under the iterations of the reference CG on D D^dagger."""
    from oracle.port import Port, gaussian_fields
    P, lat = Port(nx, nt), sb.Lattice(nx, nt)
    U = P.hot_start(31)
    phi, _ = gaussian_fields(nx, nt, 32)
    even, odd = _parity_masks(nx, nt)
    m = m0 + 2
    phi_e = phi * even

    def Dhat(v, dagger=False):
        op = lat.D_dagger_phi if dagger else lat.D_phi
        w = op(U, v, m0) * odd
        return (m * v - op(U, w, m0) / m) * even

    x, ok, its = lat.evenodd_solve(U, phi, m0)
    assert ok == 1 and np.abs(x[:, odd]).max() == 0.0
    res = phi_e - Dhat(Dhat(x, True))
    assert np.linalg.norm(res) <= 2e-10 * np.linalg.norm(phi_e)
    # Schur identity: with z = Dhat^dagger x the full-lattice solution of D y = (phi_e, 0) has y_even = z
    z = Dhat(x, True)
    y = z - (lat.D_phi(U, z, m0) * odd) / m               # y_odd = -(1/m) (D z)_odd
    assert np.linalg.norm(lat.D_phi(U, y, m0) - phi_e) <= 1e-9 * np.linalg.norm(phi_e)
    _, ok_ref, its_ref = lat.conjugate_gradient(U, phi_e, m0)
    assert ok_ref == 1 and its < 0.7 * its_ref, (its, its_ref)
    lat.close()


def test_evenodd_hmc_force_reversibility_and_chain(sb):
    """The opt-in even-odd HMC: pseudofermion on the even sites, S_pf = phi_e^dagger (Dhat Dhat^dagger)^-1 phi_e.
    Force = -dS/d(omega) by central differences through the Hamiltonian entry point, leapfrog reversibility, and a short
    chain on 32x32 (beta = 2, m0 = 0): the plaquette agrees with the reference-exact chain's, <exp(-dH)> ~ 1, and a
    near-critical trajectory spends less than half of the operator applications of the reference solver."""
    from oracle.port import Port, gaussian_fields
    g = load_golden(8, 8)
    lat = sb.Lattice(8, 8)
    lat.set_solver("evenodd")
    U, pi, phi = g["U"].copy(), g["pi"], g["phi"]
    m0, beta = float(g["m0"]), float(g["beta"])
    lat.set_cg(1e-13, 10000)
    lat.hmc_configure(beta, m0, 6, 0.5)
    lat.hmc_set_gauge(U)
    F, ok = lat.hmc_force(phi)
    assert ok == 1
    zero = np.zeros_like(pi)
    w = 1e-5
    for mu, n in [(0, 3), (1, 5), (0, 63), (1, 18)]:
        Up, Um = U.copy(), U.copy()
        Up[mu, n] *= np.exp(1j * w)
        Um[mu, n] *= np.exp(-1j * w)
        lat.hmc_set_gauge(Up)
        sp = lat.hmc_hamiltonian(zero, phi)
        lat.hmc_set_gauge(Um)
        sm = lat.hmc_hamiltonian(zero, phi)
        dS = (sp - sm) / (2 * w)
        assert abs(F[mu, n] + dS) < 2e-5 * max(1.0, abs(dS)), (mu, n, F[mu, n], dS)
    lat.hmc_set_gauge(U)
    U1, p1, _ = lat.hmc_leapfrog(pi, phi)
    lat.hmc_set_gauge(U1)
    U2, p2, _ = lat.hmc_leapfrog(-p1, phi)
    assert np.abs(U2 - U).max() < 1e-11 and np.abs(p2 + pi).max() < 1e-10
    lat.close()
    # chain statistics (64^2 gives <P> = 0.7187(11); 32^2 agrees within the tolerance used here)
    n = 32
    lat = sb.Lattice(n, n)
    lat.set_solver("evenodd")
    h = sb.HMC(lat, Port(n, n).hot_start(12345), 10, 1.0, 80, 40, 0, 2.0, 0.0, seed=7)
    Ep, dEp = h.HMC_algorithm()
    assert abs(Ep - 0.7187) < 0.02, Ep
    assert 0.3 < h.getacceptance_rate() <= 1.0
    dH = np.array([x[0] for x in h.history[80:]])
    assert abs(np.exp(-dH).mean() - 1.0) < 0.5
    assert all(ok for (_, _, _, ok, _) in h.history)
    lat.close()
    # near the critical mass (BASELINE configs[4] parameters) the even-odd trajectory needs far fewer applications
    n = 128
    lat = sb.Lattice(n, n)
    U = Port(n, n).hot_start(12345)
    chi, pi = gaussian_fields(n, n, 777)
    apps = {}
    for solver in ("reference", "evenodd"):
        lat.set_solver(solver)
        lat.hmc_configure(2.0, -0.18, 20, 1.0)
        lat.hmc_set_gauge(U)
        lat.hmc_inject(pi, chi)
        r = lat.hmc_trajectory()
        assert r.cg_all_converged == 1
        apps[solver] = r.dd_applications
    assert apps["evenodd"] < 0.5 * apps["reference"], apps
    lat.close()


@pytest.mark.parametrize("ghosts", ["t", "xt"])
def test_one_pass_with_ghost_columns_on_one_tile(sb, ghosts):
    """The one-pass D D^dagger of lattices split along t (2-deep ghost columns from packed strips, ghost rows widened by
    the corner entries, d's ghosts written by the CG pass) exercised on ONE GPU: SM_SELF_GHOSTS makes a single tile take
    its own opposite edges as ghosts through the same pack kernels, buffers and kernel paths a split lattice uses (only
    the NCCL send/recv is replaced by a device copy).  Against the oracle and against the plain single-tile path, on
    ragged shapes, incl. CG stopped mid-way (the ghost columns of d_k are rebuilt every iteration)."""
    from oracle.port import Port, gaussian_fields
    # (tiles with >= 3 strips take the overlapped form: edge strips and boundary bands behind the exchange on the comm stream,
    # interior strips on the compute stream, one reduction over the three launches)
    for nx, nt, m0 in [(64, 48, -0.05), (37, 300, 0.02), (300, 37, 0.0), (256, 256, 0.0), (130, 70, 0.1), (8, 8, 0.2),
                       (64, 1000, 0.0), (300, 600, -0.03)]:
        P = Port(nx, nt)
        U = P.hot_start(77)
        phi, _ = gaussian_fields(nx, nt, 78)
        os.environ.update(SM_SELF_GHOSTS=ghosts, SM_CLUSTER_CG="0", SM_DD_PATH="onepass")
        lat = sb.Lattice(nx, nt)
        for k in ("SM_SELF_GHOSTS", "SM_CLUSTER_CG", "SM_DD_PATH"):
            os.environ.pop(k)
        assert lat.one_pass_dd()
        want = P.DDdag(U, phi, m0)
        assert relerr(lat.D_D_dagger_phi(U, phi, m0), want) <= TOL_D, (nx, nt)
        xo, oko, apps, _ = P.cg(U, phi, m0)
        x, ok, its = lat.conjugate_gradient(U, phi, m0)
        assert ok == oko == 1 and abs(its + 2 - apps) <= 1, (nx, nt, its, apps)
        assert relerr(x, xo) <= TOL_X, (nx, nt)
        lat.set_cg(1e-10, 9)
        xm, okm, itm = lat.conjugate_gradient(U, phi, m0)
        assert okm == 0 and itm == 9 and relerr(xm, P.cg(U, phi, m0, 1e-10, 9)[0]) <= 1e-10, (nx, nt)
        lat.close()
