"""A few CG solves on 64x64 (cluster-resident kernel) and 256x256 (grid-resident kernel): profiling target."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import schwingermodel_b200 as sb  # noqa: E402
from bench import synthetic_links, synthetic_spinor  # noqa: E402

for n in (64, 256):
    lat = sb.Lattice(n, n)
    dU, dphi, dx = lat.new_field(True, synthetic_links(n * n, 1)), lat.new_field(True, synthetic_spinor(n * n, 2)), lat.new_field()
    for _ in range(4):
        ok, its = lat.dev_cg(dU, dphi, dx, 0.0)
    print(n, ok, its, lat.last_kernel_ms(), "ms")
    lat.close()
