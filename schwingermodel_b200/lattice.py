"""Host-side mirror of the reference's operator interface, on top of the C ABI.

Names and argument meaning follow the reference headers (include/dirac_operator.h:71-93,
include/conjugate_gradient.h:16, include/gauge_conf.h:79-114, include/hmc.h:8-66):

    lat = Lattice(Nx, Nt)
    Dphi = lat.D_phi(U, phi, m0)               # D_phi(U, phi, Dphi, m0)
    x, ok, its = lat.conjugate_gradient(U, phi, m0)

Fields are numpy arrays of shape (2, V): complex128 for spinors / gauge links (row mu0, row mu1 =
the reference's spinor.mu0 / spinor.mu1), float64 for momenta and forces (re_field); site
n = x*Nt + t.  All arithmetic happens in libschwinger_b200.so on the GPU; nothing here computes.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _abi
from ._abi import HmcParams, TrajResult, check, dp


def _c2(a, V):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    if a.shape != (2, V):
        raise ValueError(f"expected a complex field of shape (2, {V}), got {a.shape}")
    return a


def _r2(a, V):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.shape != (2, V):
        raise ValueError(f"expected a real field of shape (2, {V}), got {a.shape}")
    return a


def _p(row):
    return row.ctypes.data_as(dp)


class DeviceField:
    """A field resident in HBM (sm_field_alloc)."""

    def __init__(self, lat: "Lattice", complex_field: bool = True):
        self.lat, self.complex = lat, bool(complex_field)
        ptr = dp()
        check(lat.lib.sm_field_alloc(lat.ctx, int(self.complex), C.byref(ptr)))
        self.ptr = ptr

    def upload(self, a):
        V = self.lat.V
        a = _c2(a, V) if self.complex else _r2(a, V)
        check(self.lat.lib.sm_field_upload(self.lat.ctx, self.ptr, _p(a[0]), _p(a[1]), int(self.complex)))
        return self

    def download(self):
        V = self.lat.V
        out = np.empty((2, V), np.complex128 if self.complex else np.float64)
        check(self.lat.lib.sm_field_download(self.lat.ctx, self.ptr, _p(out[0]), _p(out[1]), int(self.complex)))
        return out

    def free(self):
        if self.ptr and getattr(self.lat, "ctx", None):     # sm_destroy already released every field of a closed lattice
            check(self.lat.lib.sm_field_free(self.lat.ctx, self.ptr))
        self.ptr = None


class Lattice:
    """One lattice (or one rank's tile of it) bound to one GPU."""

    def __init__(self, Nx: int, Nt: int, device: int = 0, ranks_x: int = 1, ranks_t: int = 1, rank: int = 0,
                 nccl_id: bytes | None = None):
        self.lib = _abi.load()
        self.Nx, self.Nt = int(Nx), int(Nt)
        self.ranks_x, self.ranks_t, self.rank = int(ranks_x), int(ranks_t), int(rank)
        ctx = _abi.ctx_p()
        if ranks_x * ranks_t == 1:
            check(self.lib.sm_create(self.Nx, self.Nt, int(device), C.byref(ctx)))
        else:
            if nccl_id is None or len(nccl_id) != _abi.SM_NCCL_ID_BYTES:
                raise ValueError("a split lattice needs the 128-byte NCCL id from nccl_unique_id()")
            buf = C.create_string_buffer(nccl_id, _abi.SM_NCCL_ID_BYTES)
            check(self.lib.sm_create_dist(self.Nx, self.Nt, self.ranks_x, self.ranks_t, self.rank, int(device), buf,
                                          C.byref(ctx)))
        self.ctx = ctx
        dims = (C.c_int * 4)()
        check(self.lib.sm_local_dims(self.ctx, dims))
        self.width_x, self.width_t = dims[0], dims[1]
        self.V = self.width_x * self.width_t      # mpi::maxSize
        self.tol, self.max_iter = 1e-10, 10000    # CG::tol, CG::max_iter (src/main.cpp:26-27)

    @staticmethod
    def nccl_unique_id() -> bytes:
        lib = _abi.load()
        buf = C.create_string_buffer(_abi.SM_NCCL_ID_BYTES)
        check(lib.sm_nccl_unique_id(buf))
        return buf.raw

    # -- peer-memory halos (x-only splits): export, all-gather by the caller, connect ------------------
    def p2p_handle(self) -> bytes:
        buf = C.create_string_buffer(_abi.SM_P2P_HANDLE_BYTES)
        check(self.lib.sm_p2p_handle(self.ctx, buf))
        return buf.raw

    def p2p_connect(self, handles_in_rank_order):
        blob = b"".join(handles_in_rank_order)
        if len(blob) != _abi.SM_P2P_HANDLE_BYTES * self.ranks_x * self.ranks_t:
            raise ValueError("need one 64-byte handle per rank")
        buf = C.create_string_buffer(blob, len(blob))
        check(self.lib.sm_p2p_connect(self.ctx, buf))

    def p2p_connect_all(self, dist_module, device="cuda"):
        """Gather every rank's window handle with torch.distributed (plumbing only) and connect."""
        import torch
        mine = torch.frombuffer(bytearray(self.p2p_handle()), dtype=torch.uint8).to(device)
        parts = [torch.zeros_like(mine) for _ in range(self.ranks_x * self.ranks_t)]
        dist_module.all_gather(parts, mine)
        self.p2p_connect([bytes(p.cpu().numpy().tobytes()) for p in parts])

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.sm_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- bookkeeping ---------------------------------------------------------------------------
    def set_cg(self, tol: float = 1e-10, max_iter: int = 10000):
        check(self.lib.sm_set_cg(self.ctx, float(tol), int(max_iter)))
        self.tol, self.max_iter = float(tol), int(max_iter)

    def set_solver(self, mixed_precision=False):
        """False / "reference": the reference's double-precision CG (default).  True / "mixed": opt-in mixed-precision
        defect correction.  "chrono": opt-in chronological start vectors inside a trajectory."""
        code = {False: _abi.SM_SOLVER_REFERENCE, True: _abi.SM_SOLVER_MIXED, "reference": _abi.SM_SOLVER_REFERENCE,
                "mixed": _abi.SM_SOLVER_MIXED, "chrono": _abi.SM_SOLVER_CHRONO,
                "evenodd": _abi.SM_SOLVER_EVENODD}[mixed_precision]
        check(self.lib.sm_set_solver(self.ctx, code))

    def last_kernel_ms(self) -> float:
        ms = C.c_double()
        check(self.lib.sm_last_kernel_ms(self.ctx, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        n = C.c_longlong()
        check(self.lib.sm_launch_count(self.ctx, C.byref(n)))
        return n.value

    def one_pass_dd(self) -> bool:
        v = C.c_int()
        check(self.lib.sm_one_pass_dd(self.ctx, C.byref(v)))
        return bool(v.value)

    def peer_mode(self) -> int:
        """0: single tile / NCCL; 1: halo rows by peer-memory stores; 2: halos and CG sums by the kernels over peer memory."""
        v = C.c_int()
        check(self.lib.sm_peer_mode(self.ctx, C.byref(v)))
        return v.value

    def new_field(self, complex_field=True, init=None) -> DeviceField:
        f = DeviceField(self, complex_field)
        if init is not None:
            f.upload(init)
        return f

    # -- geometry (include/dirac_operator.h:35-62) -----------------------------------------------
    def periodic_boundary(self, ranks_x=1, ranks_t=1, rank=0):
        m = (self.Nx // ranks_x) * (self.Nt // ranks_t)
        rpb, lpb = np.zeros(2 * m, np.int32), np.zeros(2 * m, np.int32)
        sr, sl = np.zeros(2 * m, np.complex128), np.zeros(2 * m, np.complex128)
        a, b = np.zeros(m, np.int32), np.zeros(m, np.int32)
        ip = _abi.ip
        check(self.lib.sm_tables(self.ctx, ranks_x, ranks_t, rank, rpb.ctypes.data_as(ip), lpb.ctypes.data_as(ip),
                                 sr.ctypes.data_as(dp), sl.ctypes.data_as(dp), a.ctypes.data_as(ip),
                                 b.ctypes.data_as(ip)))
        return dict(RightPB=rpb, LeftPB=lpb, SignR=sr, SignL=sl, x_1_t1=a, x1_t_1=b)

    # -- operators on host arrays (drop-in calls: copies inside) -----------------------------------
    def _stencil(self, fn, U, phi, m0):
        U, phi = _c2(U, self.V), _c2(phi, self.V)
        out = np.empty_like(phi)
        check(fn(self.ctx, _p(U[0]), _p(U[1]), _p(phi[0]), _p(phi[1]), _p(out[0]), _p(out[1]), float(m0)))
        return out

    def D_phi(self, U, phi, m0):
        return self._stencil(self.lib.sm_D_phi, U, phi, m0)

    def D_dagger_phi(self, U, phi, m0):
        return self._stencil(self.lib.sm_D_dagger_phi, U, phi, m0)

    def D_D_dagger_phi(self, U, phi, m0):
        return self._stencil(self.lib.sm_D_D_dagger_phi, U, phi, m0)

    def dot(self, x, y) -> complex:
        x, y = _c2(x, self.V), _c2(y, self.V)
        o = np.zeros(2)
        check(self.lib.sm_dot(self.ctx, _p(x[0]), _p(x[1]), _p(y[0]), _p(y[1]), _p(o)))
        return complex(o[0], o[1])

    def conjugate_gradient(self, U, phi, m0):
        """-> (x, converged, iterations); converged is the reference's return value (1/0)."""
        U, phi = _c2(U, self.V), _c2(phi, self.V)
        x = np.empty_like(phi)
        ok, its = C.c_int(0), C.c_int(0)
        check(self.lib.sm_conjugate_gradient(self.ctx, _p(U[0]), _p(U[1]), _p(phi[0]), _p(phi[1]), _p(x[0]), _p(x[1]),
                                             float(m0), C.byref(ok), C.byref(its)))
        return x, ok.value, its.value

    def evenodd_solve(self, U, phi, m0):
        """x_e = (Dhat Dhat^dagger)^-1 phi_e on the even sites (opt-in even-odd solver) -> (x, converged, iterations)"""
        U, phi = _c2(U, self.V), _c2(phi, self.V)
        x = np.empty_like(phi)
        ok, its = C.c_int(0), C.c_int(0)
        check(self.lib.sm_evenodd_solve(self.ctx, _p(U[0]), _p(U[1]), _p(phi[0]), _p(phi[1]), _p(x[0]), _p(x[1]), float(m0),
                                        C.byref(ok), C.byref(its)))
        return x, ok.value, its.value

    def phi_dag_partialD_phi(self, U, left, right):
        U, left, right = _c2(U, self.V), _c2(left, self.V), _c2(right, self.V)
        F = np.empty((2, self.V))
        check(self.lib.sm_phi_dag_partialD_phi(self.ctx, _p(U[0]), _p(U[1]), _p(left[0]), _p(left[1]), _p(right[0]),
                                               _p(right[1]), _p(F[0]), _p(F[1])))
        return F

    def Compute_Staple(self, U):
        U = _c2(U, self.V)
        K = np.empty_like(U)
        check(self.lib.sm_compute_staple(self.ctx, _p(U[0]), _p(U[1]), _p(K[0]), _p(K[1])))
        return K

    def Compute_Plaquette01(self, U, beta=1.0, want_field=True):
        """-> (Plaquette01 or None, MeasureSp_HMC, Compute_gaugeAction(beta))"""
        U = _c2(U, self.V)
        P = np.empty(self.V, np.complex128) if want_field else None
        s = np.zeros(2)
        check(self.lib.sm_compute_plaquette(self.ctx, _p(U[0]), _p(U[1]), float(beta),
                                            _p(P) if want_field else None, _p(s)))
        return P, float(s[0]), float(s[1])

    # -- device-resident operators -----------------------------------------------------------------
    def dev_D(self, U: DeviceField, src: DeviceField, dst: DeviceField, m0, dagger=False):
        check(self.lib.sm_dev_D(self.ctx, U.ptr, src.ptr, dst.ptr, float(m0), int(dagger)))

    def dev_DDdag(self, U: DeviceField, src: DeviceField, dst: DeviceField, m0):
        check(self.lib.sm_dev_DDdag(self.ctx, U.ptr, src.ptr, dst.ptr, float(m0)))

    def dev_DDdag_loop(self, U, src, dst, m0, reps) -> float:
        ms = C.c_double()
        check(self.lib.sm_dev_DDdag_loop(self.ctx, U.ptr, src.ptr, dst.ptr, float(m0), int(reps), C.byref(ms)))
        return ms.value

    def dev_dot(self, x: DeviceField, y: DeviceField) -> complex:
        o = np.zeros(2)
        check(self.lib.sm_dev_dot(self.ctx, x.ptr, y.ptr, _p(o)))
        return complex(o[0], o[1])

    def dev_cg(self, U: DeviceField, phi: DeviceField, x: DeviceField, m0):
        ok, its = C.c_int(0), C.c_int(0)
        check(self.lib.sm_dev_cg(self.ctx, U.ptr, phi.ptr, x.ptr, float(m0), C.byref(ok), C.byref(its)))
        return ok.value, its.value

    # -- HMC on device-resident state (src/hmc.cpp) ------------------------------------------------
    def hmc_configure(self, beta, m0, md_steps, trajectory_length):
        p = HmcParams(float(beta), float(m0), float(trajectory_length), int(md_steps))
        check(self.lib.sm_hmc_configure(self.ctx, C.byref(p)))

    def hmc_set_gauge(self, U):
        U = _c2(U, self.V)
        check(self.lib.sm_hmc_set_gauge(self.ctx, _p(U[0]), _p(U[1])))

    def hmc_get_gauge(self, proposal=False):
        U = np.empty((2, self.V), np.complex128)
        check(self.lib.sm_hmc_get_gauge(self.ctx, _p(U[0]), _p(U[1]), int(proposal)))
        return U

    def hmc_get_momenta(self, proposal=False):
        p = np.empty((2, self.V))
        check(self.lib.sm_hmc_get_momenta(self.ctx, _p(p[0]), _p(p[1]), int(proposal)))
        return p

    def hmc_get_phi(self):
        p = np.empty((2, self.V), np.complex128)
        check(self.lib.sm_hmc_get_phi(self.ctx, _p(p[0]), _p(p[1])))
        return p

    def hmc_get_chi(self):
        p = np.empty((2, self.V), np.complex128)
        check(self.lib.sm_hmc_get_chi(self.ctx, _p(p[0]), _p(p[1])))
        return p

    def hmc_refresh(self, seed, trajectory_index):
        check(self.lib.sm_hmc_refresh(self.ctx, int(seed), int(trajectory_index)))

    def hmc_inject(self, pi, chi):
        pi, chi = _r2(pi, self.V), _c2(chi, self.V)
        check(self.lib.sm_hmc_inject(self.ctx, _p(pi[0]), _p(pi[1]), _p(chi[0]), _p(chi[1])))

    def hmc_trajectory(self) -> TrajResult:
        r = TrajResult()
        check(self.lib.sm_hmc_trajectory(self.ctx, C.byref(r)))
        return r

    def hmc_accept(self, accept: bool):
        check(self.lib.sm_hmc_accept(self.ctx, int(bool(accept))))

    def hmc_force(self, phi):
        phi = _c2(phi, self.V)
        F = np.empty((2, self.V))
        ok = C.c_int(0)
        check(self.lib.sm_hmc_force(self.ctx, _p(phi[0]), _p(phi[1]), _p(F[0]), _p(F[1]), C.byref(ok)))
        return F, ok.value

    def hmc_hamiltonian(self, pi, phi) -> float:
        pi, phi = _r2(pi, self.V), _c2(phi, self.V)
        H = C.c_double()
        check(self.lib.sm_hmc_hamiltonian(self.ctx, _p(pi[0]), _p(pi[1]), _p(phi[0]), _p(phi[1]), C.byref(H)))
        return H.value

    def hmc_leapfrog(self, pi, phi):
        """-> (U', pi', all CG converged)"""
        pi, phi = _r2(pi, self.V), _c2(phi, self.V)
        ok = C.c_int(0)
        check(self.lib.sm_hmc_leapfrog(self.ctx, _p(pi[0]), _p(pi[1]), _p(phi[0]), _p(phi[1]), C.byref(ok)))
        return self.hmc_get_gauge(True), self.hmc_get_momenta(True), ok.value


# -- configuration files (src/gauge_conf.cpp:378-423, :495-546) -------------------------------------
def SaveConf(U, Nx, Nt, name: str):
    lib = _abi.load()
    U = _c2(U, Nx * Nt)
    check(lib.sm_save_conf(int(Nx), int(Nt), _p(U[0]), _p(U[1]), name.encode()))


def readBinary(Nx, Nt, name: str):
    lib = _abi.load()
    U = np.empty((2, Nx * Nt), np.complex128)
    check(lib.sm_read_conf(int(Nx), int(Nt), name.encode(), _p(U[0]), _p(U[1])))
    return U


def format_tag(x: float) -> str:
    """`format()` of include/variables.h:197-203: fixed, 4 decimals, decimal point removed."""
    return f"{x:.4f}".replace(".", "", 1)


def hot_start_angles_to_links(theta):
    """U = exp(i theta) for (2, V) angles -- convenience for synthetic inputs (host-side numpy)."""
    theta = np.asarray(theta, dtype=np.float64)
    return np.cos(theta) + 1j * np.sin(theta)


class HMC:
    """HMC driver with the reference's control flow (src/hmc.cpp:151-215): the trajectory runs on
    the GPU, the Metropolis test and the measurement bookkeeping stay on the host."""

    def __init__(self, lat: Lattice, U, MD_steps, trajectory_length, Ntherm, Nmeas, Nsteps, beta, m0, saveconf=0,
                 seed=12345, rng=None):
        self.lat = lat
        self.MD_steps, self.trajectory_length = int(MD_steps), float(trajectory_length)
        self.Ntherm, self.Nmeas, self.Nsteps = int(Ntherm), int(Nmeas), int(Nsteps)
        self.beta, self.m0, self.saveconf = float(beta), float(m0), int(saveconf)
        self.seed = int(seed)
        self.rng = rng or np.random.default_rng(seed)      # Metropolis uniforms (the reference uses rand())
        self.acceptance = 0.0
        self.therm = False
        self.traj_index = 0
        self.Ep = self.dEp = self.gS = self.dgS = 0.0
        self.sum_re_plaq = float("nan")
        self.gauge_action = float("nan")
        self.history = []
        lat.hmc_configure(beta, m0, MD_steps, trajectory_length)
        lat.hmc_set_gauge(U)

    def HMC_Update(self, pi=None, chi=None):
        lat = self.lat
        if pi is None:
            lat.hmc_refresh(self.seed, self.traj_index)
        else:
            lat.hmc_inject(pi, chi)
        self.traj_index += 1
        r = lat.hmc_trajectory()
        u = self.rng.random()
        accept = u <= math.exp(-r.dH) if r.dH > -700 else True
        lat.hmc_accept(accept)
        if accept:
            self.sum_re_plaq, self.gauge_action = r.sum_re_plaq_new, r.gauge_action_new
            if self.therm:
                self.acceptance += 1.0
        else:
            self.sum_re_plaq, self.gauge_action = r.sum_re_plaq_old, r.gauge_action_old
        self.history.append((r.dH, bool(accept), r.dd_applications, r.cg_all_converged, r.kernel_ms))
        return r, accept

    def HMC_algorithm(self, on_conf=None):
        Ntot = self.lat.Nx * self.lat.Nt
        for _ in range(self.Ntherm):
            self.HMC_Update()
        self.therm = True
        sp, ga = [], []
        n_after_therm = 0
        for i in range(self.Nmeas):
            self.HMC_Update()
            n_after_therm += 1
            sp.append(self.sum_re_plaq)
            ga.append(self.gauge_action)
            if on_conf is not None:
                on_conf(i, self)
            if i != self.Nmeas - 1:
                for _ in range(self.Nsteps):
                    self.HMC_Update()
                    n_after_therm += 1
        self.n_after_therm = n_after_therm
        self.Ep = float(np.mean(sp)) / Ntot
        self.gS = float(np.mean(ga)) / Ntot
        self.dEp = jackknife_error(sp, 20) / Ntot if len(sp) >= 20 else float("nan")
        self.dgS = jackknife_error(ga, 20) / Ntot if len(ga) >= 20 else float("nan")
        return self.Ep, self.dEp

    def getacceptance_rate(self, conf_number=None):
        n = conf_number if conf_number else max(1, getattr(self, "n_after_therm", 1))
        return self.acceptance / n


def jackknife_error(dat, bins: int) -> float:
    """Jackknife_error of src/statistics.cpp:6-34 (leave-one-bin-out), host bookkeeping."""
    dat = np.asarray(dat, dtype=np.float64)
    n = len(dat)
    per = n // bins
    mean = dat.sum() / n
    tot = dat[: bins * per].reshape(bins, per).sum(axis=1)
    all_sum = tot.sum()
    means = (all_sum - tot) / (n - per)
    return float(np.sqrt(((means - mean) ** 2).sum() * (bins - 1) / bins))
