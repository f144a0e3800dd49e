"""Single Wilson stencil application (k_wilson PLAIN) vs resident blocks per SM (GPU box)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import schwingermodel_b200 as sb  # noqa: E402
from bench import synthetic_links, synthetic_spinor  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
V = n * n
U, phi = synthetic_links(V, 1), synthetic_spinor(V, 2)
for occ in sys.argv[2].split(","):
    os.environ["SM_DD_PATH"] = "twopass"
    if occ != "auto":
        os.environ["SM_WILSON_BLOCKS_PER_SM"] = occ
    lat = sb.Lattice(n, n)
    os.environ.pop("SM_DD_PATH")
    os.environ.pop("SM_WILSON_BLOCKS_PER_SM", None)
    dU, dphi, dout = lat.new_field(True, U), lat.new_field(True, phi), lat.new_field()
    lat.dev_DDdag_loop(dU, dphi, dout, 0.0, 3)
    ms = min(lat.dev_DDdag_loop(dU, dphi, dout, 0.0, 20) for _ in range(3)) / 40      # two stencils per D D^dagger
    print(json.dumps({"n": n, "blocks_per_sm": occ, "stencil_us": round(ms * 1e3, 1), "GBs_96": round(96 * V / ms / 1e6)}), flush=True)
    lat.close()
