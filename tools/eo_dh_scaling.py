import sys, json, numpy as np
sys.path.insert(0, '/root/repo')
import schwingermodel_b200 as sb
from oracle.port import Port, gaussian_fields
n = 64
U = Port(n, n).hot_start(12345)
chi, pi = gaussian_fields(n, n, 777)
lat = sb.Lattice(n, n)
for solver in ("reference", "evenodd"):
    lat.set_solver(solver)
    out = []
    for md in (10, 20, 40, 80):
        lat.set_cg(1e-12, 10000)
        lat.hmc_configure(2.0, 0.0, md, 1.0)
        lat.hmc_set_gauge(U)
        lat.hmc_inject(pi, chi)
        r = lat.hmc_trajectory()
        out.append((md, r.dH, r.H_old))
    print(solver, json.dumps(out))
# thermalised start: run the reference chain 200 trajectories, then compare dH of both solvers from that configuration
lat.set_solver("reference")
h = sb.HMC(lat, U, 10, 1.0, 0, 0, 0, 2.0, 0.0, seed=3)
for _ in range(200):
    h.HMC_Update()
Ut = lat.hmc_get_gauge(False)
for solver in ("reference", "evenodd"):
    lat.set_solver(solver)
    dhs = []
    for s in range(6):
        chi, pi = gaussian_fields(n, n, 900 + s)
        lat.hmc_configure(2.0, 0.0, 10, 1.0)
        lat.hmc_set_gauge(Ut)
        lat.hmc_inject(pi, chi)
        dhs.append(round(lat.hmc_trajectory().dH, 4))
    print("thermalised", solver, dhs)
