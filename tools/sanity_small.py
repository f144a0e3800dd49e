"""Every kernel once on small, ragged lattices (for compute-sanitizer): checks against the oracle too."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import schwingermodel_b200 as sb  # noqa: E402
from oracle.port import Port, gaussian_fields  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


for nx, nt, env in [(6, 10, {}), (37, 70, {}), (37, 70, {"SM_CLUSTER_CG": "0"}), (37, 70, {"SM_CLUSTER_CG": "0", "SM_GRAPHS": "0"}),
                    (37, 70, {"SM_CLUSTER_CG": "0", "SM_DD_PATH": "twopass"}), (96, 130, {}),
                    (70, 300, {"SM_CLUSTER_CG": "0", "SM_FUSED_ROWS": "5", "SM_FUSED_BT": "256"})]:
    os.environ.update(env)
    lat = sb.Lattice(nx, nt)
    for k in env:
        os.environ.pop(k)
    P = Port(nx, nt)
    U = P.hot_start(3)
    chi, pi = gaussian_fields(nx, nt, 4)
    phi, _ = gaussian_fields(nx, nt, 5)
    m0, beta = 0.05, 2.0
    e = [rel(lat.D_phi(U, phi, m0), P.D(U, phi, m0)), rel(lat.D_dagger_phi(U, phi, m0), P.D(U, phi, m0, True)),
         rel(lat.D_D_dagger_phi(U, phi, m0), P.DDdag(U, phi, m0))]
    x, ok, its = lat.conjugate_gradient(U, phi, m0)
    xo = P.cg(U, phi, m0)[0]
    e.append(rel(x, xo))
    e.append(rel(lat.Compute_Staple(U), P.staple(U)))
    e.append(rel(lat.Compute_Plaquette01(U, beta)[0], P.plaquette(U, beta)[0]))
    e.append(rel(lat.phi_dag_partialD_phi(U, x, phi), P.fermion_force(U, x, phi)))
    lat.hmc_configure(beta, m0, 3, 0.3)
    lat.hmc_set_gauge(U)
    lat.hmc_inject(pi, chi)
    r = lat.hmc_trajectory()
    t = P.trajectory(U, pi, chi, 3, 0.3, beta, m0)
    lat.hmc_refresh(1, 2)
    lat.hmc_trajectory()
    lat.hmc_accept(True)
    print(nx, nt, env, "max rel err", max(e), "dH diff", abs(r.dH - t["dH"]), flush=True)
    assert max(e) < 1e-9 and abs(r.dH - t["dH"]) < 1e-8
    lat.close()
print("SANITY_OK")
