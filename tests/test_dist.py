"""N > 1: (a) CPU, gloo, world size 2 -- the host-side plumbing of a split lattice (tile placement,
NCCL-id style broadcast, max-over-ranks timing) without a GPU; (b) GPU -- tests/dist/dist_check.py
under torchrun on every GPU of the box (skipped with fewer than 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from schwingermodel_b200.tiles import assemble, tile_of
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nx, nt = 8, 12
        rng = np.random.default_rng(7)               # same global field on every rank
        g = rng.normal(size=(2, nx * nt)) + 1j * rng.normal(size=(2, nx * nt))
        for rx, rt in [(2, 1), (1, 2)]:
            mine = tile_of(g, nx, nt, rx, rt, rank)
            # broadcast of an opaque 128-byte id from rank 0, as bench.py does with the NCCL id
            idt = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                idt.copy_(torch.arange(128, dtype=torch.uint8))
            dist.broadcast(idt, 0)
            assert idt.tolist() == list(range(128))
            parts = [torch.zeros(mine.shape, dtype=torch.complex128) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(mine))
            back = assemble([p.numpy() for p in parts], nx, nt, rx, rt)
            assert np.array_equal(back, g)
            # a local-site index maps to the global site the reference's tables imply
            wx, wt = nx // rx, nt // rt
            cx, ct = divmod(rank, rt)
            n = 3 * wt + 1
            assert mine[0, n] == g[0, (cx * wx + 3) * nt + ct * wt + 1]
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)     # max over ranks, the bench's timing rule
        assert t.item() == float(world)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _halo_worker(rank, world, port, q):
    """One-pass D D^dagger on a lattice split over two ranks, the exchange carried by gloo: every rank sends the two boundary
    rows (split along x) or columns (split along t) of U and psi in the library's ghost layouts, pads its tile with what it
    receives, puts the antiperiodic sign on the seam links as the kernels do, and the D D^dagger of the padded tile (an
    exact sub-problem for the tile's own sites: the operator reaches two sites) must equal the tile of the global result."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle.port import Port, gaussian_fields
    from schwingermodel_b200.tiles import boundary_cols2, boundary_rows2, neighbours, seam_signs, tile_of
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nx, nt, m0 = 12, 16, -0.07
        G = Port(nx, nt)
        U = G.hot_start(31)
        psi, _ = gaussian_fields(nx, nt, 32)
        want_global = G.DDdag(U, psi, m0)

        def swap(lo, hi, to_lo, to_hi):
            """send `lo` to rank to_lo and `hi` to rank to_hi; receive the neighbours' pieces: (from to_lo, from to_hi)."""
            mine = torch.from_numpy(np.stack([lo, hi]))
            parts = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            # the -side neighbour's "hi" piece is my "lo" ghost, the +side neighbour's "lo" piece my "hi" ghost
            return parts[to_lo][1].numpy(), parts[to_hi][0].numpy()

        for rx, rt in [(2, 1), (1, 2)]:
            wx, wt = nx // rx, nt // rt
            nb = neighbours(rx, rt, rank)
            sR, sL = seam_signs(rt, rank)
            tU, tp = tile_of(U, nx, nt, rx, rt, rank), tile_of(psi, nx, nt, rx, rt, rank)
            padded = []
            for tile in (tU, tp):
                f = tile.reshape(2, wx, wt)
                if rt == 1:
                    lo, hi = boundary_rows2(tile, wx, wt)
                    g_lo, g_hi = swap(lo, hi, nb["xm"], nb["xp"])
                    padded.append(np.concatenate([g_lo, f, g_hi], axis=1))          # rows -2 .. wx+1
                else:
                    lo, hi = boundary_cols2(tile, wx, wt)
                    g_lo, g_hi = swap(lo, hi, nb["tm"], nb["tp"])
                    padded.append(np.concatenate([g_lo, f, g_hi], axis=2))          # columns -2 .. wt+1
            pU, pp = padded
            if rt == 1:
                got = Port(wx + 4, nt).DDdag(pU.reshape(2, -1), pp.reshape(2, -1), m0).reshape(2, wx + 4, nt)[:, 2:wx + 2, :]
            else:
                # the time link of column -1 carries the sign of the -t hop into column 0, that of column wt-1 the sign of
                # the +t hop out of it (csrc/sm_fused_tma.cuh: h0); the padded problem's own seam only reaches its outer
                # two columns on each side, which are dropped
                pU = pU.copy()
                pU[0, :, 1] *= sL
                pU[0, :, wt + 1] *= sR
                got = Port(nx, wt + 4).DDdag(pU.reshape(2, -1), pp.reshape(2, -1), m0).reshape(2, nx, wt + 4)[:, :, 2:wt + 2]
            want = tile_of(want_global, nx, nt, rx, rt, rank).reshape(2, wx, wt)
            err = float(np.abs(got - want).max() / np.abs(want).max())
            assert err <= 1e-14, (rx, rt, rank, err)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_one_pass_halo_conventions_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_halo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_neighbours_and_seam_signs_match_the_reference_tables():
    """neighbours() / seam_signs() against the oracle's per-rank tables (the reference's periodic_boundary, bit-exact to
    the reference build): a rank's SignR is -1 on its last column exactly when seam_signs says so, SignL on its first."""
    from oracle.port import Port
    from schwingermodel_b200.tiles import neighbours, seam_signs
    nx, nt = 8, 12
    P = Port(nx, nt)
    for rx, rt in [(1, 1), (2, 1), (1, 2), (2, 2), (4, 3), (2, 6)]:
        wx, wt = nx // rx, nt // rt
        for rank in range(rx * rt):
            T = P.tables(rx, rt, rank)
            sR, sL = seam_signs(rt, rank)
            n_last, n_first = 0 * wt + wt - 1, 0
            assert T["SignR"][2 * n_last + 0].real == sR and T["SignL"][2 * n_first + 0].real == sL
            assert T["SignR"][2 * n_first + 0].real == (sR if wt == 1 else 1.0)
            nb = neighbours(rx, rt, rank)
            cx, ct = divmod(rank, rt)
            assert nb["xp"] == ((cx + 1) % rx) * rt + ct and nb["tm"] == cx * rt + (ct - 1) % rt
            assert neighbours(rx, rt, nb["xp"])["xm"] == rank and neighbours(rx, rt, nb["tp"])["tm"] == rank


def test_split_lattice_plumbing_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_tiles_reject_uneven_split():
    from schwingermodel_b200.tiles import tile_shape
    with pytest.raises(ValueError):
        tile_shape(8, 12, 3, 1)


@pytest.mark.gpu
def test_split_lattice_parity_on_all_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        os.path.join(ROOT, "tests", "dist", "dist_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
