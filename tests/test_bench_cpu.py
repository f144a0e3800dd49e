"""bench.py on the CPU box: the reference arm runs here (it times the reference's CPU code), the B200 arm
must refuse to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-lattice", "256")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "DD^dagger site-updates/s" and d["unit"] == "site-updates/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "0", "--lattice", "64", "--skip-extra")
    assert r.returncode != 0          # no CPU fallback: it must not print a number
    assert not any(l.startswith("{") for l in r.stdout.splitlines())


def test_bench_tiles_come_from_one_global_field_and_seam_check_detects_errors():
    """bench.py's inputs are tiles of one global lattice at every GPU count, and its seam-band oracle check accepts the
    oracle's own D D^dagger on every decomposition (rows and, for splits along t, columns across the antiperiodic
    seam) while flagging a single wrong site next to a seam."""
    import numpy as np

    sys.path.insert(0, ROOT)
    import bench
    from oracle.port import Port
    from schwingermodel_b200.tiles import tile_of

    nx, nt, m0 = 192, 160, -0.03
    U = bench.synthetic_tile("links", 1000, nx, nt)
    phi = bench.synthetic_tile("spinor", 2000, nx, nt)
    assert abs(np.abs(U) - 1).max() < 1e-14 and abs(phi.real.var() - 0.5) < 0.02
    full = Port(nx, nt).DDdag(U, phi, m0)
    for rx, rt in [(1, 1), (3, 1), (2, 2), (1, 4), (4, 2)]:
        for rank in range(rx * rt):
            assert np.array_equal(bench.synthetic_tile("links", 1000, nx, nt, rx, rt, rank), tile_of(U, nx, nt, rx, rt, rank))
            assert np.array_equal(bench.synthetic_tile("spinor", 2000, nx, nt, rx, rt, rank), tile_of(phi, nx, nt, rx, rt, rank))
            got = tile_of(full, nx, nt, rx, rt, rank)
            assert bench.seam_band_check(got, nx, nt, rx, rt, rank, m0, 1000, 2000) <= 1e-14, (rx, rt, rank)
        bad = tile_of(full, nx, nt, rx, rt, 0).copy()
        bad[1, 3] += 1e-9                                   # row 0, column 3 of rank 0's tile
        assert bench.seam_band_check(bad, nx, nt, rx, rt, 0, m0, 1000, 2000) > 1e-11
    # rows wrap around the lattice
    a = bench._rows("links", 7, nt, -2, 3, nx)
    assert np.array_equal(a[:, :2], bench._rows("links", 7, nt, nx - 2, nx, nx)) and np.array_equal(a[:, 2:], bench._rows("links", 7, nt, 0, 3, nx))
