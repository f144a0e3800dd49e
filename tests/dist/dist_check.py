"""Split-lattice parity, one process per GPU:  torchrun --nproc-per-node N tests/dist/dist_check.py
Every rank holds a tile; results are compared with the single-rank oracle on the global lattice
(the reference's multi-rank code equals its single-rank code, tests/test_oracle_vs_reference.py)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import schwingermodel_b200 as sb  # noqa: E402
from oracle.port import Port, gaussian_fields  # noqa: E402
from schwingermodel_b200.tiles import tile_of  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def check_lattice(nx, nt, decomps, rank, world, local):
    """Every kernel path of a split lattice against the single-rank oracle on the global nx x nt lattice."""
    m0, beta, md, tau = -0.05, 2.0, 5, 0.5
    P = Port(nx, nt)
    U = P.hot_start(12345)
    chi, pi = gaussian_fields(nx, nt, 777)
    phi, _ = gaussian_fields(nx, nt, 778)
    want = dict(D=P.D(U, phi, m0), Ddag=P.D(U, phi, m0, True), DD=P.DDdag(U, phi, m0), dot=P.dot(phi, chi),
                staple=P.staple(U))
    xo, oko, apps, _ = P.cg(U, phi, m0)
    want["ff"] = P.fermion_force(U, xo, P.D(U, xo, m0, True))
    Pl, sp, sg = P.plaquette(U, beta)
    F, _ = P.force(U, phi, beta, m0)
    tr = P.trajectory(U, pi, chi, md, tau, beta, m0)
    report = []
    # every decomposition with the default path choice, and the x-only splits again with the one-pass
    # D D^dagger forced (2-row ghosts), which small lattices would not pick on their own
    # x-only splits connect peer-memory windows by default (halo rows and CG sums by the kernels themselves, CG batches
    # as CUDA graphs); "nccl" runs them again with SM_P2P=0 (ncclSend/Recv halos, ncclAllReduce sums, plain launches),
    # "p2p-explicit" through the sm_p2p_handle / sm_p2p_connect entry points
    cases = [(rx, rt, None) for rx, rt in decomps] + [(rx, rt, "nccl") for rx, rt in decomps if rt == 1 and rx > 1]
    cases += [(rx, rt, "p2p-explicit") for rx, rt in decomps if rt == 1 and rx > 1 and nx // rx >= 4]
    for rx, rt, path in cases:
        os.environ["SM_P2P"] = "1" if path is None else "0"
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(sb.Lattice.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        lat = sb.Lattice(nx, nt, device=local, ranks_x=rx, ranks_t=rt, rank=rank, nccl_id=idt.cpu().numpy().tobytes())
        os.environ.pop("SM_P2P", None)
        one_pass = lat.one_pass_dd()
        if path == "p2p-explicit":
            lat.p2p_connect_all(dist)     # halo rows by peer-memory stores instead of NCCL send/recv
        peer_mode = lat.peer_mode()
        if path is None and rt == 1 and rx > 1 and nx // rx >= 4:
            assert peer_mode == 2, peer_mode          # the default on x-only splits: everything over peer memory
        if path == "nccl":
            assert peer_mode == 0
        T = lambda f: tile_of(f, nx, nt, rx, rt, rank)   # noqa: E731
        tabs, otabs = lat.periodic_boundary(rx, rt, rank), P.tables(rx, rt, rank)
        assert all(np.array_equal(tabs[k], otabs[k]) for k in tabs)
        e = {}
        e["D"] = rel(lat.D_phi(T(U), T(phi), m0), T(want["D"]))
        e["Ddag"] = rel(lat.D_dagger_phi(T(U), T(phi), m0), T(want["Ddag"]))
        e["DD"] = rel(lat.D_D_dagger_phi(T(U), T(phi), m0), T(want["DD"]))
        z = lat.dot(T(phi), T(chi))
        e["dot"] = abs(z - want["dot"]) / abs(want["dot"])
        x, ok, its = lat.conjugate_gradient(T(U), T(phi), m0)
        assert ok == 1 and abs(its + 2 - apps) <= 1, (ok, its, apps)
        e["cg"] = rel(x, T(xo))
        e["ff"] = rel(lat.phi_dag_partialD_phi(T(U), T(xo), T(P.D(U, xo, m0, True))), T(want["ff"]))
        e["staple"] = rel(lat.Compute_Staple(T(U)), T(want["staple"]))
        Pg, spg, sgg = lat.Compute_Plaquette01(T(U), beta)
        e["plaq"] = rel(Pg[None, :], T(Pl[None, :]))
        e["sp"] = abs(spg - sp) / abs(sp)
        e["sg"] = abs(sgg - sg) / abs(sg)
        lat.hmc_configure(beta, m0, md, tau)
        lat.hmc_set_gauge(T(U))
        Fg, fok = lat.hmc_force(T(phi))
        e["force"] = rel(Fg, T(F))
        lat.hmc_inject(T(pi), T(chi))
        r = lat.hmc_trajectory()
        e["dH"] = abs(r.dH - tr["dH"])
        e["U'"] = float(np.abs(lat.hmc_get_gauge(True) - T(tr["U"])).max())
        e["pi'"] = float(np.abs(lat.hmc_get_momenta(True) - T(tr["pi"])).max())
        # device RNG is indexed by the global site: tiles of the same global field on every decomposition
        lat.hmc_refresh(99, 3)
        mine = lat.hmc_get_momenta(False)
        ref = None
        if rx * rt > 1:
            one = sb.Lattice(nx, nt, device=local)
            one.hmc_configure(beta, m0, md, tau)
            one.hmc_refresh(99, 3)
            ref = T(one.hmc_get_momenta(False))
            one.close()
            assert np.array_equal(mine, ref)
        lat.close()
        tol = dict(D=1e-13, Ddag=1e-13, DD=1e-13, dot=1e-12, cg=1e-9, ff=1e-12, staple=1e-14, plaq=1e-14, sp=1e-12,
                   sg=1e-12, force=1e-8, dH=1e-8)
        tol["U'"] = 1e-9
        tol["pi'"] = 1e-8
        bad = {k: v for k, v in e.items() if not v <= tol[k]}
        report.append({"lattice": [nx, nt], "ranks_x": rx, "ranks_t": rt, "path": path or "default",
                       "one_pass": one_pass, "peer_mode": peer_mode, "errors": e, "bad": bad})
        assert not bad, (rx, rt, bad)
        dist.barrier()
    return report


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    all_decomps = [(rx, world // rx) for rx in range(1, world + 1) if world % rx == 0]
    report = []
    # 64 x 48: every decomposition.  512 x 48 split along x only: tiles of >= 64 rows at 8 ranks, so the one-pass pass is
    # the split launch (interior chunks on the compute stream, boundary bands behind the exchange on the comm stream)
    # at every GPU count.  48 x 512 split along t only: tiles >= 64 columns wide, the overlapped k_wilson_boundary path.
    # 256 x 256: the 2-D decompositions at a size where every tile still has an interior.
    lattices = [(64, 48, all_decomps), (512, 48, [(world, 1)]), (48, 512, [(1, world)])]
    if world >= 4:
        lattices.append((256, 256, [d for d in all_decomps if d[0] > 1 and d[1] > 1]))
    for nx, nt, decomps in lattices:
        report += check_lattice(nx, nt, decomps, rank, world, local)
    if rank == 0:
        print(json.dumps({"world": world, "checked": report}))
        print("DIST_CHECK_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
