"""Trajectories per second with the reference-exact solver and the opt-in even-odd HMC (GPU box).
usage: eo_traj.py N beta m0 MD [ntraj]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import schwingermodel_b200 as sb  # noqa: E402
from bench import synthetic_links  # noqa: E402

n, beta, m0, md = int(sys.argv[1]), float(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
ntr = int(sys.argv[5]) if len(sys.argv) > 5 else 3
for solver in ("reference", "evenodd"):
    lat = sb.Lattice(n, n)
    lat.set_solver(solver)
    h = sb.HMC(lat, synthetic_links(n * n, 4), md, 1.0, 0, 0, 0, beta, m0, seed=12)
    h.HMC_Update()
    t0 = time.perf_counter()
    for _ in range(ntr):
        h.HMC_Update()
    dt = time.perf_counter() - t0
    print(json.dumps({"n": n, "beta": beta, "m0": m0, "md": md, "solver": solver, "traj_per_s": round(ntr / dt, 3),
                      "applications_per_traj": int(np.mean([x[2] for x in h.history[1:]])),
                      "kernel_ms_per_traj": round(float(np.mean([x[4] for x in h.history[1:]])), 2),
                      "dH": [round(x[0], 3) for x in h.history], "all_converged": all(x[3] for x in h.history)}), flush=True)
    lat.close()
