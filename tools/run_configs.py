"""The five BASELINE.json configs on the GPU, with the reference's CPU code timed beside them where it
finishes in seconds (GPU box; writes one JSON document to stdout)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import schwingermodel_b200 as sb  # noqa: E402
from oracle import ref as refmod  # noqa: E402
from oracle.port import Port, gaussian_fields  # noqa: E402

out = {"host_cores": os.cpu_count()}

# configs[0]: 64x64, beta=2, m0=0, MD=10, tau=1, 100 trajectories
lat = sb.Lattice(64, 64)
U0 = Port(64, 64).hot_start(12345)
h = sb.HMC(lat, U0, 10, 1.0, 80, 20, 0, 2.0, 0.0, seed=3)
t0 = time.perf_counter()
h.HMC_algorithm()
dt = time.perf_counter() - t0
cfg0 = {"gpu_traj_per_s": 100 / dt, "seconds_100_traj": dt, "Ep": h.Ep, "dEp": h.dEp, "acceptance": h.getacceptance_rate(),
        "mean_dd_applications_per_traj": float(np.mean([x[2] for x in h.history])),
        "all_cg_converged": all(x[3] for x in h.history)}
if refmod.available(64, 64):
    R = refmod.Ref(64, 64)
    Utherm = lat.hmc_get_gauge(False)       # a thermalised field: what the reference spends its time on
    chi, pi = gaussian_fields(64, 64, 5)
    t1 = time.perf_counter()
    tr = R.trajectory(Utherm, pi, chi, 10, 1.0, 2.0, 0.0)
    cfg0["reference_seconds_per_traj_1core_thermalised"] = time.perf_counter() - t1
    cfg0["reference_traj_per_s_1core"] = 1.0 / cfg0["reference_seconds_per_traj_1core_thermalised"]
    lat.hmc_configure(2.0, 0.0, 10, 1.0)
    lat.hmc_set_gauge(Utherm)
    lat.hmc_inject(pi, chi)
    t1 = time.perf_counter()
    r = lat.hmc_trajectory()
    cfg0["gpu_seconds_same_traj"] = time.perf_counter() - t1
    cfg0["dH_gpu_minus_reference"] = r.dH - tr["dH"]
out["config0_64x64_hmc"] = cfg0
lat.close()

# configs[1]: 256x256, m0=0, one CG solve on a hot start
lat = sb.Lattice(256, 256)
P = Port(256, 256)
U = P.hot_start(12345)
phi, _ = gaussian_fields(256, 256, 777)
dU, dphi, dx = lat.new_field(True, U), lat.new_field(True, phi), lat.new_field(True)
for _ in range(3):
    lat.dev_cg(dU, dphi, dx, 0.0)
ts = []
for _ in range(20):
    ok, its = lat.dev_cg(dU, dphi, dx, 0.0)
    ts.append(lat.last_kernel_ms())
cfg1 = {"gpu_ms_per_solve": float(np.mean(ts)), "gpu_solves_per_s": 1e3 / float(np.mean(ts)), "iterations": its, "converged": ok}
lat.conjugate_gradient(U, phi, 0.0)      # first call allocates the staging fields
t1 = time.perf_counter()
x, ok2, its2 = lat.conjugate_gradient(U, phi, 0.0)
cfg1["gpu_ms_per_solve_host_buffers"] = (time.perf_counter() - t1) * 1e3
if refmod.available(256, 256):
    R = refmod.Ref(256, 256)
    xr, okr, apps, sec = R.cg(U, phi, 0.0)
    cfg1.update(reference_seconds_1core=sec, reference_dd_applications=apps,
                x_rel_err=float(np.abs(x - xr).max() / np.abs(xr).max()))
    cores = 1
    while cores * 2 <= min(os.cpu_count() or 1, 16):
        cores *= 2
    sec_n, apps_n, _ = R.timed("cg", U, phi, 0.0, cores, 1)
    cfg1.update(reference_seconds_all_cores=sec_n, reference_cores=cores)
out["config1_256x256_cg"] = cfg1
lat.close()

# configs[2]: 1024x1024, beta=4, m0=-0.05, trajectories (MD=10, tau=1) from a hot start
from bench import synthetic_links  # noqa: E402
lat = sb.Lattice(1024, 1024)
h = sb.HMC(lat, synthetic_links(1024 * 1024, 3), 10, 1.0, 0, 0, 0, 4.0, -0.05, seed=11)
h.HMC_Update()
t0 = time.perf_counter()
for _ in range(5):
    h.HMC_Update()
dt = time.perf_counter() - t0
out["config2_1024x1024_hmc"] = {"gpu_traj_per_s": 5 / dt, "mean_dd_applications_per_traj": float(np.mean([x[2] for x in h.history[1:]])),
                                "device_ms_per_traj": float(np.mean([x[4] for x in h.history[1:]])),
                                "all_cg_converged": all(x[3] for x in h.history)}
lat.close()

# configs[4]: 512x512, beta=2, m0=-0.18, MD=20 (near critical: long CG)
lat = sb.Lattice(512, 512)
Uh = synthetic_links(512 * 512, 4)
h = sb.HMC(lat, Uh, 20, 1.0, 0, 0, 0, 2.0, -0.18, seed=12)
h.HMC_Update()
t0 = time.perf_counter()
for _ in range(3):
    h.HMC_Update()
dt = time.perf_counter() - t0
c4 = {"gpu_traj_per_s": 3 / dt, "mean_dd_applications_per_traj": float(np.mean([x[2] for x in h.history[1:]])),
      "cg_solves_per_traj": 21, "all_cg_converged": all(x[3] for x in h.history), "dH": [x[0] for x in h.history]}
phi, _ = gaussian_fields(512, 512, 9)
t1 = time.perf_counter()
x, ok, its = lat.conjugate_gradient(Uh, phi, -0.18)
c4.update(gpu_cg_seconds_host_buffers=time.perf_counter() - t1, cg_iterations_hot_start=its)
t1 = time.perf_counter()
xo, oko, apps, _ = Port(512, 512).cg(Uh, phi, -0.18)
c4.update(port_cg_seconds_1core=time.perf_counter() - t1, port_dd_applications=apps,
          x_rel_err=float(np.abs(x - xo).max() / np.abs(xo).max()))
out["config4_512x512_near_critical"] = c4
lat.close()
print(json.dumps(out, indent=1))
