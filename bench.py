#!/usr/bin/env python
"""bench.py -- the headline measurement of BASELINE.json: DD^dagger site-updates/s on the
8192x8192 lattice (beta=2, m0=0; configs[3]), with CG solves/s (configs[1], 256^2) and HMC
trajectories/s (configs[2], 1024^2) reported beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W]             the B200 path (libschwinger_b200.so)
    python bench.py --impl reference [...]                           the reference's CPU code on host cores

One "step" is one D D^dagger application over the whole lattice (one launch of the one-pass kernel per GPU; on a split
lattice an interior and a boundary launch that run concurrently).
`value` is site-updates/s with every field resident in HBM; `e2e` is the same unit measured
through the reference-facing conjugate_gradient() call of the C ABI with pinned HOST buffers
(U and phi copied in, x copied out inside the timed region): DD^dagger applications the solve
performed x sites / wall time.  N > 1 splits the same 8192^2 lattice over ranks_x = N GPUs
(strong scaling, the decomposition configs[3] names); launch with torchrun as the driver does.  Every N works on
tiles of the same global synthetic lattice, and the line's `parity` block checks this run against the CPU oracle (seam
bands of D D^dagger) and against the single-GPU checksums in tests/golden/bench_expect.json.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")     # keep stdout to the one JSON line

METRIC = "DD^dagger site-updates/s"
UNIT = "site-updates/s"
BYTES_PER_STENCIL_SITE = 96      # read psi 32 + read U 32 + write 32  
BYTES_PER_DD_SITE = 192


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin=None, t_end=None, t_load=None):
        """Median SM clock and throttle reasons over the samples taken inside [t_begin, t_end] (the timed
        region); if the region was shorter than the sampling period, over [t_load, t_end] (GPU under the
        same load since the warm-up) -- `window` says which."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def pick(lo, hi):
            sm, mx, reasons = [], [], set()
            for ts, r in self.rows:
                if lo is not None and not (lo - 0.03 <= ts <= hi + 0.03):
                    continue
                try:
                    sm.append(float(r[1]))
                    mx.append(float(r[2]))
                except (ValueError, IndexError):
                    continue
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons

        window = "timed region"
        sm, mx, reasons = pick(t_begin, t_end)
        if not sm and t_load is not None:
            window = "warm-up + timed region (timed region shorter than the sampling period)"
            sm, mx, reasons = pick(t_load, t_end)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


ROW_BLOCK = 64     # rows per generator block


def _rows(kind, seed, Nt, x0, x1, Nx):
    """Rows [x0, x1) (taken modulo Nx) of ONE global synthetic field, whole in t: shape (2, x1-x0, Nt).
    The field is a pure function of (kind, seed, Nx, Nt): block b = rows [64 b, 64 b + 64) comes from
    default_rng([seed, b]), so every rank of every decomposition builds tiles of the same global lattice.
      kind "links":  hot-start gauge field U = exp(i theta), theta ~ U[0, 2 pi)   (gauge_conf.cpp:23-36)
      kind "spinor": Gaussian pseudofermion source, re, im ~ N(0, 1/sqrt 2)       (hmc.cpp:19-28)"""
    out = np.empty((2, x1 - x0, Nt), np.complex128)
    x = x0
    while x < x1:
        b = (x % Nx) // ROW_BLOCK
        lo = b * ROW_BLOCK
        take = min(x1 - x, lo + ROW_BLOCK - (x % Nx), Nx - (x % Nx))
        rng = np.random.default_rng([seed, b])
        nb = min(ROW_BLOCK, Nx - lo)
        if kind == "links":
            th = rng.random((2, nb, Nt)) * (2.0 * np.pi)
            blk = np.cos(th) + 1j * np.sin(th)
        else:
            g = rng.standard_normal((2, nb, Nt, 2)) * np.sqrt(0.5)
            blk = g[..., 0] + 1j * g[..., 1]
        o = (x % Nx) - lo
        out[:, x - x0:x - x0 + take, :] = blk[:, o:o + take, :]
        x += take
    return out


def synthetic_tile(kind, seed, Nx, Nt, rx=1, rt=1, rank=0):
    """(2, wx*wt) tile of `rank` of the global synthetic field (rank = cx*ranks_t + ct, include/mpi_setup.h:39-47)."""
    wx, wt = Nx // rx, Nt // rt
    cx, ct = divmod(rank, rt)
    rows = _rows(kind, seed, Nt, cx * wx, (cx + 1) * wx, Nx)
    return np.ascontiguousarray(rows[:, :, ct * wt:(ct + 1) * wt]).reshape(2, wx * wt)


def synthetic_links(V, seed, Nx=None, Nt=None):
    n = int(round(np.sqrt(V))) if Nx is None else Nx
    return synthetic_tile("links", seed, n, V // n if Nt is None else Nt)


def synthetic_spinor(V, seed, Nx=None, Nt=None):
    n = int(round(np.sqrt(V))) if Nx is None else Nx
    return synthetic_tile("spinor", seed, n, V // n if Nt is None else Nt)


def seam_band_check(got_tile, Nx, Nt, rx, rt, rank, m0, seed_U, seed_phi, R=12):
    """D D^dagger of this rank's tile against the CPU oracle on the bands of rows (and, when the lattice is split along
    t, columns) next to the tile's edges -- the sites whose stencil reaches into the neighbour's tile.  A band with two
    extra rows on each side is an exact sub-problem for its inner rows (D D^dagger reaches two sites); the inputs are
    regenerated from the global field's seeds, so no rank needs another rank's data.  -> max relative error."""
    from oracle.port import Port
    wx, wt = Nx // rx, Nt // rt
    cx, ct = divmod(rank, rt)
    got = got_tile.reshape(2, wx, wt)
    worst, scale = 0.0, float(np.abs(got[:, :R]).max())
    R = min(R, wx)
    for lo in (cx * wx, (cx + 1) * wx - R):                       # first and last R rows of the tile
        U = _rows("links", seed_U, Nt, lo - 2, lo + R + 2, Nx).reshape(2, -1)
        phi = _rows("spinor", seed_phi, Nt, lo - 2, lo + R + 2, Nx).reshape(2, -1)
        want = Port(R + 4, Nt).DDdag(U, phi, m0).reshape(2, R + 4, Nt)[:, 2:R + 2, ct * wt:(ct + 1) * wt]
        x = lo - cx * wx
        worst = max(worst, float(np.abs(got[:, x:x + R] - want).max()))
    if rt > 1:
        C = min(R, wt)
        Ug = _rows("links", seed_U, Nt, cx * wx - 2, (cx + 1) * wx + 2, Nx)       # the tile's rows + 2 either side
        pg = _rows("spinor", seed_phi, Nt, cx * wx - 2, (cx + 1) * wx + 2, Nx)
        Ug[0, :, Nt - 1] *= -1.0          # antiperiodic in t == periodic with the time links of the last column negated
        for lo in (ct * wt, (ct + 1) * wt - C):
            cols = np.arange(lo - 2, lo + C + 2) % Nt
            U = np.ascontiguousarray(Ug[:, :, cols]).reshape(2, -1)
            phi = np.ascontiguousarray(pg[:, :, cols]).reshape(2, -1)
            want = Port(wx + 4, C + 4).DDdag(U, phi, m0).reshape(2, wx + 4, C + 4)[:, 2:wx + 2, 2:C + 2]
            t = lo - ct * wt
            worst = max(worst, float(np.abs(got[:, :, t:t + C] - want).max()))
    return worst / scale


def pinned_like(a):
    """numpy view of pinned host memory holding a copy of `a` (torch is plumbing only)."""
    import torch
    t = torch.empty(a.shape, dtype=torch.complex128 if a.dtype == np.complex128 else torch.float64, pin_memory=True)
    n = t.numpy()
    n[...] = a
    return t, n


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code (oracle/_ref) on the box's host cores
# ---------------------------------------------------------------------------------------------------
def reference_lattice(args):
    """The lattice the reference arm runs: the workload's own 8192^2 when its build (oracle/_ref/libref_8192x8192.so)
    is here and the host has the memory for it (same config as the B200 arm), else the 2048^2 sample."""
    from oracle import ref as refmod
    if args.ref_lattice:
        return args.ref_lattice
    try:
        import psutil
        roomy = psutil.virtual_memory().available >= 40 * 2 ** 30
    except Exception:  # noqa: BLE001
        roomy = False
    if roomy and refmod.available(args.lattice, args.lattice, build=False):
        return args.lattice
    return 2048


def reference_ranks(nx, n_threads=None):
    cores = os.cpu_count() or 1
    if n_threads is None:
        n_threads = cores
    rx = 1
    while rx * 2 <= min(n_threads, 64) and nx % (rx * 2) == 0 and nx // (rx * 2) >= 2:
        rx *= 2
    return rx


def run_reference(args):
    """DD^dagger site-updates/s of the UNMODIFIED reference (oracle/_ref), forked over the host cores through the
    mini-MPI shim (ranks_x = cores, ranks_t = 1), on the B200 arm's own workload when it fits the host."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ref as refmod
    n = reference_lattice(args)
    U, phi = synthetic_tile("links", 1000, n, n), synthetic_tile("spinor", 2000, n, n)
    times = None
    if refmod.available(n, n):
        kind = "reference"
        for attempt in (n, 2048):          # the bench lattice itself; if this host cannot run it, the 2048^2 sample
            try:
                if attempt != n:
                    n = attempt
                    U, phi = synthetic_tile("links", 1000, n, n), synthetic_tile("spinor", 2000, n, n)
                    if not refmod.available(n, n):
                        break
                R = refmod.Ref(n, n)
                rx = cores = reference_ranks(n)
                reps = 1 if n >= 4096 else 2
                times = []
                for i in range(args.warmup + args.steps):
                    sec, count, _ = R.timed("dd", U, phi, 0.0, rx, 1, reps=reps)
                    times.append(sec / reps)
                break
            except (RuntimeError, OSError, MemoryError) as e:
                print(f"reference arm: {attempt}x{attempt} failed ({e!r})", file=sys.stderr)
                times = None
                if attempt == 2048:
                    break
    if times is None:
        # the reference did not compile (or run) here: time the C port (1 core)
        from oracle.port import Port
        n = min(n, 2048)
        U, phi = synthetic_tile("links", 1000, n, n), synthetic_tile("spinor", 2000, n, n)
        P = Port(n, n)
        times = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            P.DDdag(U, phi, 0.0)
            times.append(time.perf_counter() - t0)
        rx, kind, cores = 1, "port", 1
    per = float(np.mean(times[args.warmup:]))
    V = n * n
    val = V / per
    frac = "" if n == args.lattice else f" (1/{(args.lattice // n) ** 2} of the {args.lattice}^2 workload)"
    sample = (f"{n}x{n} lattice{frac}, D_D_dagger_phi of the reference over {cores} forked ranks "
              f"(ranks_x={rx}, ranks_t=1), {args.steps} timed steps of one application each")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per * 1e3 * (args.lattice / n) ** 2,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"DD^dagger on {args.lattice}x{args.lattice}, beta=2, m0=0, hot-start links, Gaussian "
                               f"source (BASELINE configs[3])", "sample": sample, "same_lattice_as_b200_arm": n == args.lattice},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cores": os.cpu_count(),
    }
    print(json.dumps(line))
    return 0


def cpu_baseline(args, U=None, phi=None):
    """rank 0, N=1 only: bounded samples of the same workloads on the host cores, by the unmodified reference
    (oracle/_ref; the C port of oracle/ if the reference did not build): (i) D_D_dagger_phi on the bench lattice,
    (ii) one conjugate_gradient on 256^2 (configs[1]), (iii) HMC_Update on 64^2 (configs[0])  -- BASELINE.md 3."""
    from oracle import ref as refmod
    t_start = time.time()
    n = reference_lattice(args)
    if U is None or U.shape[1] != n * n:
        U, phi = synthetic_tile("links", 1000, n, n), synthetic_tile("spinor", 2000, n, n)
    if not refmod.available(n, n):
        from oracle.port import Port
        n = 1024
        P = Port(n, n)
        U, phi = synthetic_tile("links", 1, n, n), synthetic_tile("spinor", 2, n, n)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            P.DDdag(U, phi, 0.0)
        per = (time.perf_counter() - t0) / reps
        return {"value": n * n / per, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"{reps} x DD^dagger on a {n}x{n} lattice with the C port, 1 core"}
    R = refmod.Ref(n, n)
    rx = reference_ranks(n)
    best = None
    reps, rounds = (1 if n >= 4096 else 2), 0
    while time.time() - t_start < 15 and rounds < 6:
        sec, count, _ = R.timed("dd", U, phi, 0.0, rx, 1, reps=reps)
        per = sec / reps
        best = per if best is None else min(best, per)
        rounds += 1
    frac = "the bench lattice itself" if n == args.lattice else f"1/{(args.lattice // n) ** 2} of the workload"
    out = {"value": n * n / best, "unit": UNIT, "cores": rx, "kind": "reference",
           "sample": f"best of {rounds} x {reps} D_D_dagger_phi on a {n}x{n} lattice ({frac}) by the unmodified reference "
                     f"over {rx} forked ranks (mini-MPI shim), {os.cpu_count()} host cores present"}
    # (ii) configs[1]: one conjugate_gradient on the 256^2 hot start (src/conjugate_gradient.cpp:4), all host cores
    try:
        if refmod.available(256, 256):
            R2 = refmod.Ref(256, 256)
            U2, p2 = synthetic_tile("links", 1, 256, 256), synthetic_tile("spinor", 2, 256, 256)
            r2 = reference_ranks(256, min(os.cpu_count() or 1, 16))
            best2, apps = None, 0
            for _ in range(3):
                sec, apps, _ = R2.timed("cg", U2, p2, 0.0, r2, 1)
                best2 = sec if best2 is None else min(best2, sec)
            sec1, apps1, _ = R2.timed("cg", U2, p2, 0.0, 1, 1)
            out["cg_256"] = {"solves_per_s": 1.0 / best2, "cores": r2, "dd_applications": apps, "seconds": best2,
                             "solves_per_s_1core": 1.0 / sec1, "kind": "reference",
                             "sample": "conjugate_gradient on 256x256, hot start, m0=0, tol 1e-10 (configs[1]); best of 3"}
        # (iii) configs[0]: HMC_Update on 64^2, beta=2, m0=0, MD=10, tau=1 (src/hmc.cpp:151), one core
        if refmod.available(64, 64):
            from oracle.port import gaussian_fields
            R3 = refmod.Ref(64, 64)
            U3 = synthetic_tile("links", 5, 64, 64)
            secs = []
            for i in range(3):
                chi, pi = gaussian_fields(64, 64, 100 + i)
                tr = R3.trajectory(U3, pi, chi, 10, 1.0, 2.0, 0.0)
                secs.append(tr["seconds"])
                U3 = tr["U"]
            out["hmc_64"] = {"traj_per_s": 1.0 / float(np.mean(secs)), "cores": 1, "kind": "reference",
                             "sample": "3 x HMC_Update (injected pi, chi; always accepted) on 64x64, beta=2, m0=0, "
                                       "MD=10, tau=1 from a hot start (configs[0])"}
    except Exception as e:  # noqa: BLE001
        out["extra_error"] = repr(e)
    return out


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    if os.environ.get("SM_BENCH_TRACE"):     # debugging aid: dump every rank's Python stack after that many seconds
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["SM_BENCH_TRACE"]), exit=True)

    import schwingermodel_b200 as sb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    N = args.gpus
    torch.cuda.set_device(local_rank)
    nccl_id = None
    if N > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(sb.Lattice.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if N == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    L = args.lattice
    weak = args.scaling == "weak"
    rt = args.ranks_t                  # default 1: split along x only (contiguous halo rows, one-pass D D^dagger)
    if N % rt:
        raise SystemExit("--ranks-t must divide the number of GPUs")
    rx = N // rt
    Lx = L * rx if weak else L         # weak: every GPU keeps an L x L tile
    Lt = L * rt if weak else L
    sites = Lx * Lt
    m0, beta = 0.0, 2.0
    lat = sb.Lattice(Lx, Lt, device=local_rank, ranks_x=rx, ranks_t=rt, rank=rank, nccl_id=nccl_id)
    halo = "none"
    if N > 1:
        # x-only splits map every rank's window into every other rank at creation (CUDA IPC over NVLink): halo rows
        # are stored straight into the neighbour's HBM and the CG sums are gathered by the kernels; SM_P2P=0 or a
        # split along t keeps ncclSend/Recv + ncclAllReduce
        halo = {0: "nccl send/recv, sums by ncclAllReduce", 1: "peer-memory stores (CUDA IPC) + flags, sums by ncclAllReduce",
                2: "peer-memory stores (CUDA IPC) + flags; CG sums gathered by the kernels over peer memory, "
                   "CG batches as CUDA graphs"}[lat.peer_mode()]
    V = lat.V
    SEED_U, SEED_PHI = 1000, 2000
    U_h = synthetic_tile("links", SEED_U, Lx, Lt, rx, rt, rank)        # tiles of ONE global field: the same lattice at every N
    phi_h = synthetic_tile("spinor", SEED_PHI, Lx, Lt, rx, rt, rank)
    dU, dphi, dout = lat.new_field(True, U_h), lat.new_field(True, phi_h), lat.new_field(True)

    # ---- device-resident DD^dagger: W warm-up steps, then exactly K timed steps --------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_load = time.time()
    lat.dev_DDdag_loop(dU, dphi, dout, m0, max(args.warmup, 3))
    # keep the GPU under this load long enough for clocks to settle; every rank must make the same number of calls
    # (each one exchanges halos with its neighbours), so the ranks agree on the elapsed time
    while max_over_ranks(time.time() - t_load) < 0.4:
        lat.dev_DDdag_loop(dU, dphi, dout, m0, max(args.warmup, 3))
    l0 = lat.launch_count()
    barrier()
    t_begin = time.time()
    ms = lat.dev_DDdag_loop(dU, dphi, dout, m0, args.steps)     # CUDA events on the launching stream
    barrier()
    t_end = time.time()
    launches = lat.launch_count() - l0
    clocks = sampler.stop(t_begin, t_end, t_load + 0.2) if rank == 0 else None
    ms = max_over_ranks(ms)
    ms_per_step = ms / args.steps
    value = sites / (ms_per_step * 1e-3)

    peak, peak_src = measured_peaks()
    # dominant kernel: one k_dd_tma pass per step (on a split lattice the pass is an interior launch plus a
    # boundary-band launch that run concurrently) or two k_wilson launches (lattice split along t).
    # Algorithmic bytes per launch = 96 B x sites of the tile for BOTH kernels: a Wilson-stencil pass reads psi and U
    # and writes out once (96 B per stencil site), and the one-pass D D^dagger kernel likewise reads psi and
    # U once and writes out once per D D^dagger site-update -- the intermediate D^dagger psi never reaches HBM.
    one_pass = lat.one_pass_dd()
    # rows staged by TMA bulk copies (k_dd_tma, csrc/sm_fused_tma.cuh) unless SM_FUSED_TMA=0 selects the cp.async kernel
    try:
        tma = int(os.environ.get("SM_FUSED_TMA", "1")) != 0
    except ValueError:
        tma = False                      # the library reads the variable with atoi()
    one_pass_kernel = "k_dd_tma" if tma else "k_dd_fused"
    passes = args.steps if one_pass else 2 * args.steps
    avg_launch_ms = ms / passes
    alg_bytes = BYTES_PER_STENCIL_SITE * V
    achieved = alg_bytes / (avg_launch_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and N == 1:
        with open(tp) as f:
            traffic = json.load(f).get("k_dd_fused_bytes_per_launch" if one_pass else "k_wilson_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": (one_pass_kernel + " (one-pass D D^dagger)") if one_pass else "k_wilson (Wilson stencil D / D^dagger)",
                "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_site": BYTES_PER_STENCIL_SITE,
                "avg_launch_ms": avg_launch_ms}
    if one_pass:
        # SURVEY 8(d) quotes 192 B per D D^dagger site-update for the two-pass form and asks that a temporally blocked
        # kernel be flagged against it without changing that denominator: kept here under an explicit name; it is an
        # equivalent rate, not memory traffic, and may exceed 1
        roofline["two_pass_equivalent_192B_survey"] = {"achieved": 2 * achieved, "frac_192B_survey": 2 * achieved / peak,
                                                       "note": "site-updates/s x 192 B (two-pass accounting of SURVEY "
                                                               "8d); not a physical bandwidth"}

    # ---- parity, visible in this line: the seam bands of D D^dagger against the CPU oracle, and checksums that must
    #      agree between GPU counts (every N works on tiles of the same global lattice) ----------------------------
    parity = None
    if not args.skip_parity:
        got = dout.download()
        band = seam_band_check(got, Lx, Lt, rx, rt, rank, m0, SEED_U, SEED_PHI)
        del got
        chk = lat.dev_dot(dout, dphi)                     # <DD^dagger phi, phi>, all-reduced over the ranks
        nrm = lat.dev_dot(dout, dout).real
        parity = {"dd_seam_band_max_rel_err_vs_oracle": max_over_ranks(band), "dd_tolerance": 1e-13,
                  "dd_checked": "first and last 12 rows%s of every rank's tile against oracle/liboracle.so (CPU) on bands "
                                "regenerated from the global field's seeds" % (" and columns" if rt > 1 else ""),
                  "dd_dot_phi": [chk.real, chk.imag], "dd_norm2": nrm}
    # ---- e2e: the reference-facing conjugate_gradient() with pinned host buffers --------------------
    keepU, U_p = pinned_like(U_h)
    keepP, phi_p = pinned_like(phi_h)
    keepX, x_p = pinned_like(phi_h)
    lat.set_cg(1e-10, 10000)
    e2e_steps = max(0, min(args.e2e_steps, args.steps))
    if args.warmup > 0 and e2e_steps > 0:
        _cg_into(lat, U_p, phi_p, x_p, m0)      # one untimed solve
    barrier()
    t0 = time.perf_counter()
    apps, ok = 0, -1
    for _ in range(e2e_steps):
        x, ok, its = _cg_into(lat, U_p, phi_p, x_p, m0)
        apps += its + 2
    barrier()
    e2e_s = max(max_over_ranks(time.perf_counter() - t0), 1e-9)
    # for transparency: ONE D_D_dagger_phi call through the host-buffer ABI (what the reference only does inside CG)
    barrier()
    t1 = time.perf_counter()
    p_ = lambda r_: r_.ctypes.data_as(sb._abi.dp)   # noqa: E731
    sb._abi.check(lat.lib.sm_D_D_dagger_phi(lat.ctx, p_(U_p[0]), p_(U_p[1]), p_(phi_p[0]), p_(phi_p[1]), p_(x_p[0]),
                                            p_(x_p[1]), float(m0)))
    barrier()
    dd_call_s = max_over_ranks(time.perf_counter() - t1)
    e2e = {"value": apps * sites / e2e_s, "unit": UNIT,
           "single_dd_call": {"value": sites / dd_call_s, "unit": UNIT, "seconds": dd_call_s,
                              "note": "one sm_D_D_dagger_phi with host buffers: 96 B/site over PCIe for 192 B/site of "
                                      "algorithmic work, i.e. bound by the host link, not by the GPU"},
           "h2d_bytes_per_step": int(N * 2 * U_p.nbytes), "d2h_bytes_per_step": int(N * x_p.nbytes),
           "call": "sm_conjugate_gradient (host buffers)", "solves": e2e_steps, "dd_applications": apps,
           "cg_converged": int(ok), "seconds": e2e_s, "solves_per_s": e2e_steps / e2e_s}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"DD^dagger on {Lx}x{Lt}, beta=2, m0=0, hot-start links, Gaussian source "
                               f"(BASELINE configs[3]{' tile per GPU' if weak else ''}); ranks_x={rx}, ranks_t={rt}",
                   "l2": "inputs larger than L2 (each field %.0f MiB per GPU)" % (V * 32 / 2 ** 20),
                   "step": "one D D^dagger application over the whole lattice", "halo_exchange": halo},
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "hbm_gbs_per_gpu": BYTES_PER_STENCIL_SITE * (1 if one_pass else 2) * value / 1e9 / N,
    }

    # ---- the other two parts of BASELINE.json's metric: CG solves/s and HMC trajectories/s -----------
    if not args.skip_extra:
        # on the bench lattice itself, at this GPU count (device-resident, max over ranks)
        big = {}
        dx = lat.new_field(True)
        lat.set_cg(1e-10, 10000)
        barrier()
        t0 = time.perf_counter()
        ok, its = lat.dev_cg(dU, dphi, dx, m0)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        x_norm2 = lat.dev_dot(dx, dx).real
        if parity is not None:
            parity.update({"cg_iterations": its, "cg_converged": ok, "cg_x_norm2": x_norm2})
        big["cg"] = {"solves_per_s": 1.0 / dt, "iterations": its, "converged": ok, "seconds": dt,
                     "GBs_per_gpu_320B": 320.0 * V * (its + 1) / dt / 1e9,
                     "config": f"one (D D^dagger)^-1 solve on {Lx}x{Lt}, hot start, m0=0, tol 1e-10, device-resident"}
        if N == 1:
            # opt-in solver upgrade: same stopping criterion on the true residual, different iterate
            lat.set_solver(True)
            lat.dev_cg(dU, dphi, dx, m0)
            t0 = time.perf_counter()
            okm, itm = lat.dev_cg(dU, dphi, dx, m0)
            dtm = time.perf_counter() - t0
            lat.set_solver(False)
            big["cg_mixed_precision_opt_in"] = {"solves_per_s": 1.0 / dtm, "iterations": itm, "converged": okm, "seconds": dtm,
                                                "note": "single-precision inner CG inside a double-precision defect "
                                                        "correction; not used by any other number of this line"}
        for f in (dx, dout):
            f.free()
        h = sb.HMC(lat, U_h, 10, 1.0, 0, 0, 0, beta, m0, seed=11)
        barrier()
        t0 = time.perf_counter()
        r, acc = h.HMC_Update()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        big["hmc"] = {"traj_per_s": 1.0 / dt, "seconds": dt, "dd_applications": int(r.dd_applications),
                      "cg_solves": int(r.cg_solves), "all_cg_converged": bool(r.cg_all_converged), "dH": r.dH,
                      "config": f"one HMC trajectory on {Lx}x{Lt}, beta=2, m0=0, MD=10, tau=1, hot start, device-resident"}
        if parity is not None:
            parity.update({"hmc_H_old": r.H_old, "hmc_H_new": r.H_new, "hmc_dH": r.dH,
                           "hmc_dd_applications": int(r.dd_applications)})
        line["extra"] = {f"lattice_{L}": big}
        if N == 1:
            try:
                line["extra"].update(extra_metrics(sb, args))
            except Exception as e:  # noqa: BLE001  -- the smaller configs must not cost the headline line
                line["extra"]["error"] = repr(e)
            try:
                line["cpu_baseline"] = cpu_baseline(args, U_h, phi_h)
            except Exception as e:  # noqa: BLE001  -- e.g. the host cannot fork the 8192^2 reference: take the 2048^2 sample
                args.ref_lattice = 2048
                try:
                    line["cpu_baseline"] = cpu_baseline(args)
                    line["cpu_baseline"]["note"] = f"bench lattice failed on this host ({e!r}); 2048^2 sample instead"
                except Exception as e2:  # noqa: BLE001
                    line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                            "sample": "none", "error": repr(e2)}
    if parity is not None:
        parity["vs_single_gpu"] = compare_with_expected(parity, Lx, Lt, args)
        parity["ok"] = bool(parity["dd_seam_band_max_rel_err_vs_oracle"] <= parity["dd_tolerance"] and
                            (parity["vs_single_gpu"] is None or parity["vs_single_gpu"]["ok"]))
        line["parity"] = parity
    if rank == 0:
        print(json.dumps(line))
    lat.close()
    if N > 1:
        dist.destroy_process_group()
    return 0


EXPECT = os.path.join(ROOT, "tests", "golden", "bench_expect.json")


def compare_with_expected(parity, Lx, Lt, args):
    """Checksums of this run against the single-GPU run of the same global lattice (tests/golden/bench_expect.json,
    written by `bench.py --gpus 1 --write-expect`): the D D^dagger checksums to 1e-13, CG iteration count equal,
    |x|^2 to 1e-10, H to 1e-10 relative and dH to 1e-8 absolute x (H / 2e4)."""
    key = f"{Lx}x{Lt}"
    if args.write_expect:
        d = {}
        if os.path.exists(EXPECT):
            with open(EXPECT) as f:
                d = json.load(f)
        d[key] = {k: v for k, v in parity.items() if k.startswith(("dd_dot", "dd_norm", "cg_", "hmc_"))}
        if int(os.environ.get("RANK", "0")) == 0:
            with open(EXPECT, "w") as f:
                json.dump(d, f, indent=1, sort_keys=True)
        return None
    if not os.path.exists(EXPECT):
        return None
    with open(EXPECT) as f:
        e = json.load(f).get(key)
    if e is None:
        return None
    out = {"expected_from": "tests/golden/bench_expect.json (1 GPU)"}

    def rel(a, b):
        return abs(a - b) / max(abs(b), 1e-300)

    out["dd_dot_phi_rel"] = rel(complex(*parity["dd_dot_phi"]), complex(*e["dd_dot_phi"]))
    out["dd_norm2_rel"] = rel(parity["dd_norm2"], e["dd_norm2"])
    ok = out["dd_dot_phi_rel"] <= 1e-12 and out["dd_norm2_rel"] <= 1e-12
    if "cg_iterations" in parity and "cg_iterations" in e:
        out["cg_iterations"] = [parity["cg_iterations"], e["cg_iterations"]]
        out["cg_x_norm2_rel"] = rel(parity["cg_x_norm2"], e["cg_x_norm2"])
        ok = ok and parity["cg_iterations"] == e["cg_iterations"] and out["cg_x_norm2_rel"] <= 1e-10
    if "hmc_H_old" in parity and "hmc_H_old" in e:
        out["hmc_H_old_rel"] = rel(parity["hmc_H_old"], e["hmc_H_old"])
        out["hmc_dH_abs"] = abs(parity["hmc_dH"] - e["hmc_dH"])
        out["hmc_dH_tolerance"] = 1e-8 * max(1.0, abs(e["hmc_H_old"]) / 2e4)
        out["hmc_dd_applications"] = [parity["hmc_dd_applications"], e["hmc_dd_applications"]]
        ok = ok and out["hmc_H_old_rel"] <= 1e-10 and out["hmc_dH_abs"] <= out["hmc_dH_tolerance"]
    out["ok"] = bool(ok)
    return out


def _cg_into(lat, U, phi, x, m0):
    """sm_conjugate_gradient writing into a caller-owned (pinned) x."""
    import ctypes as C

    from schwingermodel_b200._abi import check, dp
    ok, its = C.c_int(0), C.c_int(0)
    p = lambda r: r.ctypes.data_as(dp)   # noqa: E731
    check(lat.lib.sm_conjugate_gradient(lat.ctx, p(U[0]), p(U[1]), p(phi[0]), p(phi[1]), p(x[0]), p(x[1]), float(m0),
                                        C.byref(ok), C.byref(its)))
    return x, ok.value, its.value


def extra_metrics(sb, args):
    out = {}
    # configs[1]: 256x256, beta=2, m0=0, one CG solve of (DD^dagger)^-1 on a hot-start field
    lat = sb.Lattice(256, 256)
    U, phi = synthetic_links(256 * 256, 1), synthetic_spinor(256 * 256, 2)
    dU, dphi, dx = lat.new_field(True, U), lat.new_field(True, phi), lat.new_field(True)
    for _ in range(3):
        ok, its = lat.dev_cg(dU, dphi, dx, 0.0)
    t = []
    for _ in range(10):
        ok, its = lat.dev_cg(dU, dphi, dx, 0.0)
        t.append(lat.last_kernel_ms())
    out["cg_256"] = {"solves_per_s": 1e3 / float(np.mean(t)), "iterations": its, "converged": ok,
                     "ms_per_solve": float(np.mean(t)), "config": "256x256 hot start, m0=0, tol 1e-10 (configs[1])"}
    lat.close()
    # configs[0]: 64x64, beta=2, m0=0, MD=10, tau=1
    lat = sb.Lattice(64, 64)
    h = sb.HMC(lat, synthetic_tile("links", 5, 64, 64), 10, 1.0, 0, 0, 0, 2.0, 0.0, seed=11)
    for _ in range(3):
        h.HMC_Update()
    t0 = time.perf_counter()
    ntr = 20
    for _ in range(ntr):
        h.HMC_Update()
    dt = time.perf_counter() - t0
    out["hmc_64"] = {"traj_per_s": ntr / dt, "dd_applications_per_traj": int(np.mean([x[2] for x in h.history[3:]])),
                     "all_cg_converged": all(x[3] for x in h.history),
                     "config": "64x64, beta=2, m0=0, MD=10, tau=1, from a hot start (configs[0])"}
    lat.close()
    # configs[2]: 1024x1024, beta=4, m0=-0.05, full HMC trajectories device-resident (MD=10, tau=1)
    n = 1024
    lat = sb.Lattice(n, n)
    h = sb.HMC(lat, synthetic_links(n * n, 3), 10, 1.0, 0, 0, 0, 4.0, -0.05, seed=11)
    h.HMC_Update()
    t0 = time.perf_counter()
    ntr = 3
    for _ in range(ntr):
        h.HMC_Update()
    dt = time.perf_counter() - t0
    out["hmc_1024"] = {"traj_per_s": ntr / dt, "dd_applications_per_traj": int(np.mean([x[2] for x in h.history[1:]])),
                       "kernel_ms_per_traj": float(np.mean([x[4] for x in h.history[1:]])),
                       "dH": [x[0] for x in h.history], "all_cg_converged": all(x[3] for x in h.history),
                       "config": "1024x1024, beta=4, m0=-0.05, MD=10, tau=1, from a hot start (configs[2])"}
    lat.close()
    # configs[4]: 512x512, beta=2, m0=-0.18 (near critical: ~850 CG iterations per solve), MD=20
    n = 512
    lat = sb.Lattice(n, n)
    h = sb.HMC(lat, synthetic_links(n * n, 4), 20, 1.0, 0, 0, 0, 2.0, -0.18, seed=12)
    h.HMC_Update()
    t0 = time.perf_counter()
    ntr = 2
    for _ in range(ntr):
        h.HMC_Update()
    dt = time.perf_counter() - t0
    # the same with the opt-in chronological start vectors (a different iterate, so not used by any other number)
    hc = sb.HMC(lat, synthetic_links(n * n, 4), 20, 1.0, 0, 0, 0, 2.0, -0.18, seed=12)
    lat.set_solver("chrono")
    hc.HMC_Update()
    t0c = time.perf_counter()
    for _ in range(ntr):
        hc.HMC_Update()
    dtc = time.perf_counter() - t0c
    lat.set_solver("reference")
    out["hmc_512_near_critical_chrono_opt_in"] = {"traj_per_s": ntr / dtc,
                                                  "dd_applications_per_traj": int(np.mean([x[2] for x in hc.history[1:]])),
                                                  "all_cg_converged": all(x[3] for x in hc.history),
                                                  "note": "sm_set_solver(SM_SOLVER_CHRONO): force solves start from the "
                                                          "extrapolated previous solutions"}
    # ... and the opt-in even-odd HMC (pseudofermion on the even sites, CG on the Schur complement: a different Markov chain
    # with the same stationary distribution; not used by any other number of this line)
    lat.set_solver("evenodd")
    he = sb.HMC(lat, synthetic_links(n * n, 4), 20, 1.0, 0, 0, 0, 2.0, -0.18, seed=12)
    he.HMC_Update()
    t0e = time.perf_counter()
    for _ in range(ntr):
        he.HMC_Update()
    dte = time.perf_counter() - t0e
    lat.set_solver("reference")
    out["hmc_512_near_critical_evenodd_opt_in"] = {"traj_per_s": ntr / dte,
                                                   "schur_applications_per_traj": int(np.mean([x[2] for x in he.history[1:]])),
                                                   "all_cg_converged": all(x[3] for x in he.history),
                                                   "note": "sm_set_solver(SM_SOLVER_EVENODD)"}
    out["hmc_512_near_critical"] = {"traj_per_s": ntr / dt,
                                    "dd_applications_per_traj": int(np.mean([x[2] for x in h.history[1:]])),
                                    "kernel_ms_per_traj": float(np.mean([x[4] for x in h.history[1:]])),
                                    "all_cg_converged": all(x[3] for x in h.history),
                                    "config": "512x512, beta=2, m0=-0.18, MD=20, tau=1, from a hot start (configs[4])"}
    lat.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lattice", type=int, default=8192)
    ap.add_argument("--ref-lattice", type=int, default=0,
                    help="lattice of the reference arm (default: the bench lattice if its reference build and the host "
                         "memory allow, else 2048)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--skip-extra", action="store_true")
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--write-expect", action="store_true",
                    help="(1 GPU) record this run's checksums in tests/golden/bench_expect.json for the N > 1 parity block")
    ap.add_argument("--ranks-t", type=int, default=1,
                    help="GPUs along t (default 1: all GPUs along x); > 1 exercises the strided-halo two-pass path")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): the 8192^2 lattice of configs[3] split over N GPUs; weak: an 8192^2 tile per GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
