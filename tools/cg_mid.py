"""CG on mid-size lattices (the working set just exceeds L2): L2-persistence window and row-chunk variants (GPU box).
usage: cg_mid.py N "env1;env2;..." with env = comma list of KEY=VALUE"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import schwingermodel_b200 as sb  # noqa: E402
from bench import synthetic_tile  # noqa: E402

n = int(sys.argv[1])
variants = sys.argv[2].split(";")
V = n * n
U, phi = synthetic_tile("links", 3, n, n), synthetic_tile("spinor", 2, n, n)
m0 = float(sys.argv[3]) if len(sys.argv) > 3 else -0.05
ref = None
for v in variants:
    env = dict(kv.split("=") for kv in v.split(",") if kv)
    os.environ.update(env)
    lat = sb.Lattice(n, n)
    for k in env:
        os.environ.pop(k)
    dU, dphi, dx = lat.new_field(True, U), lat.new_field(True, phi), lat.new_field()
    for _ in range(3):
        ok, its = lat.dev_cg(dU, dphi, dx, m0)
    t = []
    for _ in range(10):
        t0 = time.perf_counter()
        ok, its = lat.dev_cg(dU, dphi, dx, m0)
        t.append(time.perf_counter() - t0)
    x = dx.download()
    if ref is None:
        ref = x
    import numpy as np
    dt = min(t)
    print(json.dumps({"n": n, "env": env, "its": its, "ok": ok, "ms": round(dt * 1e3, 3), "us_per_it": round(dt / (its + 1) * 1e6, 2),
                      "GBs_320": round(320 * V * (its + 1) / dt / 1e9), "relerr_vs_first": float(np.abs(x - ref).max() / np.abs(ref).max())}),
          flush=True)
    lat.close()
