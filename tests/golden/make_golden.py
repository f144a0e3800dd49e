"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from /root/reference).

Run here (the container that has /root/reference):   python tests/golden/make_golden.py
The fixtures are what travels to the GPU box; the reference tree does not.

Inputs are deterministic: hot start = the reference's own GaugeConf::initialization after
srand(12345) (gauge_conf.cpp:23-36); sources/momenta = numpy default_rng Gaussians (seeded).
Every output array below was produced by the reference's own functions through
oracle/ref_build/ref_harness.cpp.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle.port import gaussian_fields  # noqa: E402  (input generator only)
from oracle.ref import Ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# (Nx, Nt, m0, beta, md, tau): tiny, non-square, and the 32x32 case of SURVEY section 4
CASES = [
    (8, 8, 0.1, 1.0, 4, 0.4),
    (16, 24, -0.05, 2.0, 6, 0.6),
    (32, 32, -0.05, 2.0, 10, 1.0),
]


def make(nx, nt, m0, beta, md, tau):
    R = Ref(nx, nt)
    U = R.hot_start(12345)
    chi, pi = gaussian_fields(nx, nt, 777)
    phi, _ = gaussian_fields(nx, nt, 778)
    g = dict(nx=nx, nt=nt, m0=m0, beta=beta, md=md, tau=tau, U=U, chi=chi, pi=pi, phi=phi)
    for name, tab in R.tables().items():
        g["tab_" + name] = tab
    g["tab22_rank3"] = np.concatenate([v.view(np.float64).ravel() if v.dtype == np.complex128 else v.astype(np.float64)
                                       for v in R.tables(2, 2, 3).values()])
    g["D"] = R.D(U, phi, m0)
    g["Ddag"] = R.D(U, phi, m0, True)
    g["DDdag"] = R.DDdag(U, phi, m0)
    z = R.dot(phi, chi)
    g["dot"] = np.array([z.real, z.imag])
    x, ok, apps, _ = R.cg(U, phi, m0)
    g["cg_x"], g["cg_ok"], g["cg_apps"] = x, ok, apps
    g["fforce"] = R.fermion_force(U, x, R.D(U, x, m0, True))
    g["staple"] = R.staple(U)
    P, sp, sg = R.plaquette(U, beta)
    g["plaq"], g["plaq_sums"] = P, np.array([sp, sg])
    F, fok = R.force(U, phi, beta, m0)
    g["force"] = F
    g["action"] = R.action(U, phi, beta, m0)
    g["hamiltonian"] = R.hamiltonian(U, pi, phi, beta, m0)
    Ul, pl = R.leapfrog(U, pi, phi, md, tau, beta, m0)
    g["lf_U"], g["lf_pi"] = Ul, pl
    tr = R.trajectory(U, pi, chi, md, tau, beta, m0)
    g["tr_phi"], g["tr_U"], g["tr_pi"] = tr["phi"], tr["U"], tr["pi"]
    g["tr_H"] = np.array([tr["H_old"], tr["H_new"]])
    g["tr_aux"] = np.array([tr["sp"], tr["sg"]])
    path = os.path.join(OUT, f"ref_{nx}x{nt}.npz")
    np.savez_compressed(path, **g)
    # the binary configuration file as the reference's SaveConf writes it (smallest case only)
    if nx * nt <= 64:
        R.save_conf(U, os.path.join(OUT, f"ref_{nx}x{nt}.ctxt"))
    print("wrote", path, os.path.getsize(path), "bytes; cg apps", apps)


def scalars_64():
    """Fingerprints on 64x64 (BASELINE config 1 size): CG iteration count and a trajectory's H values."""
    nx = nt = 64
    R = Ref(nx, nt)
    U = R.hot_start(12345)
    chi, pi = gaussian_fields(nx, nt, 777)
    phi, _ = gaussian_fields(nx, nt, 778)
    x, ok, apps, _ = R.cg(U, phi, 0.0)
    tr = R.trajectory(U, pi, chi, 10, 1.0, 2.0, 0.0)
    P, sp, sg = R.plaquette(tr["U"], 2.0)
    np.savez_compressed(os.path.join(OUT, "ref_64x64_scalars.npz"), cg_apps=apps, cg_ok=ok,
                        cg_x_sum=np.array([x.sum().real, x.sum().imag]), cg_x_norm=np.linalg.norm(x),
                        tr_H=np.array([tr["H_old"], tr["H_new"]]), tr_sp=sp, tr_sg=sg,
                        tr_U_sum=np.array([tr["U"].sum().real, tr["U"].sum().imag]),
                        tr_pi_norm=np.linalg.norm(tr["pi"]), strings=np.array([R.format(2.0), R.format(-0.18), R.format(0.0)]))
    print("64x64: cg apps", apps, "dH", tr["dH"])


def fingerprint_sites(n, count=256, seed=8192):
    """Sites sampled for the full-size fingerprint: seeded random ones plus the corners of the lattice, the
    antiperiodic seam columns and the wrap rows (where the one-pass kernel's strips and chunks begin and end)."""
    rng = np.random.default_rng(seed)
    xs = rng.integers(0, n, count - 16)
    ts = rng.integers(0, n, count - 16)
    ex = np.array([0, 0, n - 1, n - 1, 1, n - 2, 0, n - 1, 2, 3, n // 2, n // 2 - 1, 251, 252, 253, 4095])
    et = np.array([0, n - 1, 0, n - 1, 1, n - 2, n // 2, n // 2, n - 1, 0, 0, n - 1, 251, 252, 253, 4096])
    return np.concatenate([xs, ex]) * n + np.concatenate([ts, et])


def fingerprint_8192(ranks=8):
    """BASELINE configs[3] at its own size, produced by the UNMODIFIED reference over `ranks` forked ranks
    (ranks_x = ranks, ranks_t = 1; multi-rank == single-rank is pinned by tests/test_oracle_vs_reference.py):
    hot start srand(12345), Gaussian source (seed 777), m0 = 0:  D D^dagger phi and the CG solution at 256 sites,
    the CG's application count and norms.  ~10 minutes on 8 cores."""
    n = 8192
    R = Ref(n, n)
    U = R.hot_start(12345)
    phi, _ = gaussian_fields(n, n, 777)
    idx = fingerprint_sites(n)
    sec, _, dd = R.timed("dd", U, phi, 0.0, ranks, 1, reps=1, want_out=True)
    print(f"8192^2 D D^dagger by the reference on {ranks} ranks: {sec:.2f} s", flush=True)
    dd_s, dd_norm = dd[:, idx].copy(), float(np.linalg.norm(dd.ravel()))
    del dd
    sec, apps, x = R.timed("cg", U, phi, 0.0, ranks, 1, want_out=True)
    print(f"8192^2 CG by the reference: {sec:.1f} s, {apps} applications", flush=True)
    np.savez_compressed(os.path.join(OUT, "ref_8192x8192_fingerprint.npz"), sites=idx, m0=0.0, hot_start_seed=12345,
                        source_seed=777, ranks_x=ranks, U_s=U[:, idx], phi_s=phi[:, idx], dd_s=dd_s, dd_norm=dd_norm,
                        cg_apps=apps, cg_x_s=x[:, idx], cg_x_norm=float(np.linalg.norm(x.ravel())),
                        cg_x_sum=np.array([x.sum().real, x.sum().imag]), cg_seconds=sec)
    print("wrote ref_8192x8192_fingerprint.npz")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "8192":
        fingerprint_8192()
        sys.exit(0)
    for c in CASES:
        make(*c)
    scalars_64()
