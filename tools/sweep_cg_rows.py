"""CG time per iteration vs rows-per-chunk / block width of the one-pass kernel on mid-size lattices (GPU box)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import schwingermodel_b200 as sb  # noqa: E402
from bench import synthetic_links, synthetic_spinor  # noqa: E402

sizes = [int(v) for v in sys.argv[1].split(",")]
rows_list = [int(v) for v in sys.argv[2].split(",")]
for n in sizes:
    V = n * n
    U, phi = synthetic_links(V, 1), synthetic_spinor(V, 2)
    for bt in ("128", "256"):
        for rows in rows_list:
            os.environ.update(SM_FUSED_ROWS=str(rows), SM_FUSED_BT=bt, SM_CLUSTER_CG="0")
            lat = sb.Lattice(n, n)
            dU, dphi, dx = lat.new_field(True, U), lat.new_field(True, phi), lat.new_field()
            lat.dev_cg(dU, dphi, dx, 0.0)
            t = []
            for _ in range(3):
                t0 = time.perf_counter()
                ok, its = lat.dev_cg(dU, dphi, dx, 0.0)
                t.append(time.perf_counter() - t0)
            print(json.dumps({"n": n, "bt": bt, "rows": rows, "cg_us_per_it": round(min(t) * 1e6 / (its + 1), 2), "its": its}), flush=True)
            lat.close()
