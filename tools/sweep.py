"""Size sweep of the D D^dagger paths and the CG built on them (GPU box): prints one JSON line per case."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import schwingermodel_b200 as sb  # noqa: E402
from bench import synthetic_links, synthetic_spinor  # noqa: E402

sizes = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [64, 256, 512, 1024, 2048, 4096, 8192]
paths = sys.argv[2].split(",") if len(sys.argv) > 2 else ["twopass", "onepass"]
for n in sizes:
    V = n * n
    U, phi = synthetic_links(V, 1), synthetic_spinor(V, 2)
    for path in paths:
        env = dict(p.split("=") for p in path.split(":")[1:])
        os.environ["SM_DD_PATH"] = path.split(":")[0]
        os.environ.update(env)
        mixed = env.pop("MIXED", None)
        lat = sb.Lattice(n, n)
        for k in ["SM_DD_PATH", *env]:
            os.environ.pop(k)
        if mixed:
            lat.set_solver(True)
        dU, dphi, dout, dx = lat.new_field(True, U), lat.new_field(True, phi), lat.new_field(), lat.new_field()
        reps = max(3, min(200, int(2e9 / V / 20)))
        lat.dev_DDdag_loop(dU, dphi, dout, 0.0, 3)
        ms = min(lat.dev_DDdag_loop(dU, dphi, dout, 0.0, reps) for _ in range(3)) / reps
        lat.dev_cg(dU, dphi, dx, 0.0)
        t = []
        for _ in range(3):
            t0 = time.perf_counter()
            ok, its = lat.dev_cg(dU, dphi, dx, 0.0)
            t.append(time.perf_counter() - t0)
        cg = min(t)
        print(json.dumps({"n": n, "path": path, "dd_us": ms * 1e3, "dd_su_per_s": V / ms * 1e3,
                          "dd_GBs_192": 192 * V / ms / 1e6, "cg_ms": cg * 1e3, "cg_its": its, "cg_ok": ok,
                          "cg_us_per_it": cg * 1e6 / (its + 1), "cg_GBs_512": 512 * V * (its + 1) / cg / 1e9}), flush=True)
        lat.close()
