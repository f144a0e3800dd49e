"""One short CG solve on 1024^2 (configs[2]: beta=4 parameters, m0=-0.05) with plain launches (SM_GRAPHS=0), the
target of the ncu capture of the CG pass at the size where it runs at 0.80 of its 320-B roofline (GPU box):
    python tools/ncu_cg_1024.py && ncu --set full --clock-control none --import-source on \
        -k regex:"k_dd_tma|k_cg_resid" -s 12 -c 4 -o gpurun_out/r02_cg_1024 python tools/ncu_cg_1024.py"""
import os
import sys

os.environ["SM_GRAPHS"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import schwingermodel_b200 as sb  # noqa: E402
from bench import synthetic_tile  # noqa: E402

n = 1024
lat = sb.Lattice(n, n)
U, phi = synthetic_tile("links", 3, n, n), synthetic_tile("spinor", 2, n, n)
dU, dphi, dx = lat.new_field(True, U), lat.new_field(True, phi), lat.new_field()
lat.set_cg(1e-10, 24)                    # 24 iterations: enough launches to skip the cold ones
ok, its = lat.dev_cg(dU, dphi, dx, -0.05)
print("cg 1024^2:", ok, its, "launches", lat.launch_count())
lat.close()
