"""A few CG solves on one mid-size lattice (several-sites-per-thread resident kernels): profiling target.
usage: cg512.py [n=512] [m0=0.0]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import schwingermodel_b200 as sb  # noqa: E402
from bench import synthetic_links, synthetic_spinor  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
m0 = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
lat = sb.Lattice(n, n)
dU, dphi, dx = lat.new_field(True, synthetic_links(n * n, 1)), lat.new_field(True, synthetic_spinor(n * n, 2)), lat.new_field()
for _ in range(3):
    ok, its = lat.dev_cg(dU, dphi, dx, m0)
print(n, ok, its, lat.last_kernel_ms(), "ms")
lat.close()
