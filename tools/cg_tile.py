"""One-pass CG on a non-square single-GPU lattice with the tile shape of a split 8192^2 (e.g. 1024 x 8192 = the 8 x 1 tile):
geometry variants (strips, rows per chunk) of the one-pass kernel.  usage: cg_tile.py NX NT "env1;env2;..." """
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import schwingermodel_b200 as sb  # noqa: E402
from bench import synthetic_tile  # noqa: E402

nx, nt = int(sys.argv[1]), int(sys.argv[2])
V = nx * nt
U, phi = synthetic_tile("links", 3, nx, nt), synthetic_tile("spinor", 2, nx, nt)
for v in sys.argv[3].split(";"):
    env = dict(kv.split("=") for kv in v.split(",") if kv)
    os.environ.update(env)
    lat = sb.Lattice(nx, nt)
    for k in env:
        os.environ.pop(k)
    dU, dphi, dx = lat.new_field(True, U), lat.new_field(True, phi), lat.new_field()
    lat.dev_DDdag_loop(dU, dphi, dx, 0.0, 20)
    dd = min(lat.dev_DDdag_loop(dU, dphi, dx, 0.0, 50) for _ in range(3)) / 50
    for _ in range(2):
        ok, its = lat.dev_cg(dU, dphi, dx, 0.0)
    t = []
    for _ in range(5):
        t0 = time.perf_counter()
        ok, its = lat.dev_cg(dU, dphi, dx, 0.0)
        t.append(time.perf_counter() - t0)
    dt = min(t)
    print(json.dumps({"nx": nx, "nt": nt, "env": env, "dd_us": round(dd * 1e3, 1), "dd_GBs_96": round(96 * V / dd / 1e6),
                      "cg_its": its, "cg_us_per_it": round(dt / (its + 1) * 1e6, 1), "cg_GBs_320": round(320 * V * (its + 1) / dt / 1e9)}),
          flush=True)
    lat.close()
