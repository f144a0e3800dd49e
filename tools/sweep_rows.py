"""Row-chunk sweep of the one-pass D D^dagger on tile shapes (GPU box)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import schwingermodel_b200 as sb  # noqa: E402
from bench import synthetic_links, synthetic_spinor  # noqa: E402

shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1].split(",")]
rows_list = [int(r) for r in sys.argv[2].split(",")]
bts = sys.argv[3].split(",") if len(sys.argv) > 3 else ["256"]
for nx, nt in shapes:
    V = nx * nt
    U, phi = synthetic_links(V, 1), synthetic_spinor(V, 2)
    for bt in bts:
        for rows in rows_list:
            os.environ.update(SM_DD_PATH="onepass", SM_FUSED_ROWS=str(rows), SM_FUSED_BT=bt)
            lat = sb.Lattice(nx, nt)
            dU, dphi, dout = lat.new_field(True, U), lat.new_field(True, phi), lat.new_field()
            reps = max(5, min(200, int(4e9 / V / 20)))
            lat.dev_DDdag_loop(dU, dphi, dout, 0.0, 3)
            ms = min(lat.dev_DDdag_loop(dU, dphi, dout, 0.0, reps) for _ in range(3)) / reps
            print(json.dumps({"nx": nx, "nt": nt, "bt": bt, "rows": rows, "dd_us": round(ms * 1e3, 2),
                              "su_per_s": round(V / ms * 1e3 / 1e9, 2), "GBs_96": round(96 * V / ms / 1e6)}), flush=True)
            lat.close()
