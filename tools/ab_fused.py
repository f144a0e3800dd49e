"""A/B of the one-pass D D^dagger variants on the GPU box: per-thread cp.async staging (k_dd_fused) against TMA bulk
staging (k_dd_tma, 3 or 4 stages), sustained under load (the power cap lowers the SM clock after ~0.3 s), plus the CG
built on each.  usage: ab_fused.py [N] [variants: comma list of tma,stages[,rows]]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import schwingermodel_b200 as sb  # noqa: E402
from bench import synthetic_tile  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
variants = sys.argv[2].split(";") if len(sys.argv) > 2 else ["0,3", "1,3", "1,4"]
do_cg = len(sys.argv) > 3 and sys.argv[3] == "cg"
V = n * n
U, phi = synthetic_tile("links", 1000, n, n), synthetic_tile("spinor", 2000, n, n)
ref = None
for v in variants:
    parts = v.split(",")
    env = {"SM_FUSED_TMA": parts[0], "SM_FUSED_STAGES": parts[1]}
    if len(parts) > 2:
        env["SM_FUSED_ROWS"] = parts[2]
    if len(parts) > 3:
        env["SM_FUSED_BT"] = parts[3]
    os.environ.update(env)
    lat = sb.Lattice(n, n)
    for k in env:
        os.environ.pop(k)
    dU, dphi, dout = lat.new_field(True, U), lat.new_field(True, phi), lat.new_field()
    lat.dev_DDdag_loop(dU, dphi, dout, 0.0, 3)
    got = dout.download()
    if ref is None:
        ref = got
    err = float(np.abs(got - ref).max() / np.abs(ref).max())
    reps = max(5, int(0.5e3 / (V / 6.1e10 * 1e3)))          # ~0.5 s per timed loop
    t0 = time.time()
    burst = lat.dev_DDdag_loop(dU, dphi, dout, 0.0, 20) / 20
    while time.time() - t0 < 1.0:
        lat.dev_DDdag_loop(dU, dphi, dout, 0.0, reps)
    ms = min(lat.dev_DDdag_loop(dU, dphi, dout, 0.0, reps) for _ in range(3)) / reps
    out = {"n": n, "variant": env, "burst_ms": round(burst, 4), "sustained_ms": round(ms, 4),
           "GBs_96_sustained": round(96 * V / ms / 1e6), "relerr_vs_first": err}
    if do_cg:
        dx = lat.new_field()
        ok, its = lat.dev_cg(dU, dphi, dx, 0.0)
        t0 = time.perf_counter()
        ok, its = lat.dev_cg(dU, dphi, dx, 0.0)
        dt = time.perf_counter() - t0
        out.update({"cg_its": its, "cg_ok": ok, "cg_s": round(dt, 4), "cg_us_per_it": round(dt / (its + 1) * 1e6, 1),
                    "cg_GBs_320": round(320 * V * (its + 1) / dt / 1e9)})
    print(json.dumps(out), flush=True)
    lat.close()
