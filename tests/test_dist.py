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
