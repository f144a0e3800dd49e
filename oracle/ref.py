"""ctypes loader for oracle/_ref/libref_<NS>x<NT>.so — the UNMODIFIED reference, compiled in place.

TEST INFRASTRUCTURE ONLY.  Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product (schwingermodel_b200) never touches it.

Field convention: complex fields are numpy complex128 arrays of shape (2, V) (mu, site) with
site n = x*Nt + t; real fields are float64 arrays of shape (2, V).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
_BUILD = os.path.join(_HERE, "ref_build", "build_ref.sh")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def lib_path(nx: int, nt: int) -> str:
    return os.path.join(_REF_DIR, f"libref_{nx}x{nt}.so")


def available(nx: int, nt: int, build: bool = True) -> bool:
    """True if the reference library for this lattice exists (building it if the reference tree is here)."""
    if os.path.exists(lib_path(nx, nt)):
        return True
    if not build or not os.path.isdir(os.environ.get("SM_REFERENCE", "/root/reference")):
        return False
    r = subprocess.run([_BUILD, str(nx), str(nt)], capture_output=True, text=True)
    return r.returncode == 0 and os.path.exists(lib_path(nx, nt))


class Ref:
    """The reference's own functions for one compile-time lattice size."""

    def __init__(self, nx: int, nt: int):
        if not available(nx, nt):
            raise FileNotFoundError(f"no reference build for {nx}x{nt} (and /root/reference absent)")
        self.lib = C.CDLL(lib_path(nx, nt))
        self.nx, self.nt, self.V = nx, nt, nx * nt
        L = self.lib
        assert L.ref_nx() == nx and L.ref_nt() == nt
        L.ref_action.restype = C.c_double
        L.ref_hamiltonian.restype = C.c_double
        L.ref_jackknife.restype = C.c_double

    # -- helpers
    def _c(self):
        return np.zeros((2, self.V), dtype=np.complex128)

    def _r(self):
        return np.zeros((2, self.V), dtype=np.float64)

    @staticmethod
    def _cc(a):
        a = np.ascontiguousarray(a, dtype=np.complex128)
        return a

    # -- geometry
    def tables(self, ranks_x=1, ranks_t=1, rank=0):
        m = self.V // (ranks_x * ranks_t)
        rpb = np.zeros(2 * m, np.int32)
        lpb = np.zeros(2 * m, np.int32)
        sr = np.zeros(2 * m, np.complex128)
        sl = np.zeros(2 * m, np.complex128)
        a = np.zeros(m, np.int32)
        b = np.zeros(m, np.int32)
        self.lib.ref_tables(ranks_x, ranks_t, rank, _i(rpb), _i(lpb), _d(sr), _d(sl), _i(a), _i(b))
        return dict(RightPB=rpb, LeftPB=lpb, SignR=sr, SignL=sl, x_1_t1=a, x1_t_1=b)

    def hot_start(self, seed: int):
        U = self._c()
        self.lib.ref_hot_start(C.c_uint(seed), _d(U))
        return U

    # -- operators
    def D(self, U, phi, m0, dagger=False):
        U, phi, out = self._cc(U), self._cc(phi), self._c()
        self.lib.ref_D(_d(U), _d(phi), _d(out), C.c_double(m0), int(dagger))
        return out

    def DDdag(self, U, phi, m0):
        U, phi, out = self._cc(U), self._cc(phi), self._c()
        self.lib.ref_DDdag(_d(U), _d(phi), _d(out), C.c_double(m0))
        return out

    def dot(self, x, y):
        x, y = self._cc(x), self._cc(y)
        o = np.zeros(2)
        self.lib.ref_dot(_d(x), _d(y), _d(o))
        return complex(o[0], o[1])

    def cg(self, U, phi, m0, tol=1e-10, max_iter=10000):
        """-> (x, converged, DD^dagger applications, seconds)"""
        U, phi, x = self._cc(U), self._cc(phi), self._c()
        apps = C.c_int(0)
        sec = C.c_double(0)
        ok = self.lib.ref_cg(_d(U), _d(phi), _d(x), C.c_double(m0), C.c_double(tol), int(max_iter),
                             C.byref(apps), C.byref(sec))
        return x, int(ok), apps.value, sec.value

    def fermion_force(self, U, left, right):
        U, left, right, F = self._cc(U), self._cc(left), self._cc(right), self._r()
        self.lib.ref_fermion_force(_d(U), _d(left), _d(right), _d(F))
        return F

    def staple(self, U):
        U, K = self._cc(U), self._c()
        self.lib.ref_staple(_d(U), _d(K))
        return K

    def plaquette(self, U, beta=1.0):
        """-> (P[V] complex, sum Re P, beta*sum Re(1-P))"""
        U = self._cc(U)
        P = np.zeros(self.V, np.complex128)
        s = np.zeros(2)
        self.lib.ref_plaquette(_d(U), C.c_double(beta), _d(P), _d(s))
        return P, float(s[0]), float(s[1])

    # -- HMC internals
    def force(self, U, phi, beta, m0):
        U, phi, F = self._cc(U), self._cc(phi), self._r()
        ok = self.lib.ref_force(_d(U), _d(phi), C.c_double(beta), C.c_double(m0), _d(F))
        return F, int(ok)

    def action(self, U, phi, beta, m0):
        U, phi = self._cc(U), self._cc(phi)
        return float(self.lib.ref_action(_d(U), _d(phi), C.c_double(beta), C.c_double(m0)))

    def hamiltonian(self, U, pi, phi, beta, m0):
        U, phi = self._cc(U), self._cc(phi)
        pi = np.ascontiguousarray(pi, np.float64)
        return float(self.lib.ref_hamiltonian(_d(U), _d(pi), _d(phi), C.c_double(beta), C.c_double(m0)))

    def leapfrog(self, U, pi, phi, md, tau, beta, m0):
        U, phi = self._cc(U), self._cc(phi)
        pi = np.ascontiguousarray(pi, np.float64)
        Uo, po = self._c(), self._r()
        self.lib.ref_leapfrog(_d(U), _d(pi), _d(phi), int(md), C.c_double(tau), C.c_double(beta), C.c_double(m0),
                              _d(Uo), _d(po))
        return Uo, po

    def trajectory(self, U, pi, chi, md, tau, beta, m0, tol=1e-10):
        """One HMC_Update with injected pi, chi.  -> dict(phi, U, pi, H_old, H_new, dH, sp, sg, cg_ok, seconds)"""
        U, chi = self._cc(U), self._cc(chi)
        pi = np.ascontiguousarray(pi, np.float64)
        phi, Uo, po = self._c(), self._c(), self._r()
        H = np.zeros(2)
        aux = np.zeros(2)
        sec = C.c_double(0)
        ok = self.lib.ref_trajectory(_d(U), _d(pi), _d(chi), int(md), C.c_double(tau), C.c_double(beta),
                                     C.c_double(m0), C.c_double(tol), _d(phi), _d(Uo), _d(po), _d(H), _d(aux),
                                     C.byref(sec))
        return dict(phi=phi, U=Uo, pi=po, H_old=float(H[0]), H_new=float(H[1]), dH=float(H[1] - H[0]),
                    sp=float(aux[0]), sg=float(aux[1]), cg_ok=int(ok), seconds=sec.value)

    # -- files
    def save_conf(self, U, name):
        U = self._cc(U)
        self.lib.ref_save_conf(_d(U), name.encode())

    def read_binary(self, name):
        U = self._c()
        self.lib.ref_read_binary(name.encode(), _d(U))
        return U

    def format(self, v):
        buf = C.create_string_buffer(64)
        self.lib.ref_format(C.c_double(v), buf, 64)
        return buf.value.decode()

    def jackknife(self, dat, bins):
        dat = np.ascontiguousarray(dat, np.float64)
        return float(self.lib.ref_jackknife(_d(dat), len(dat), int(bins)))

    # -- CPU baseline
    def timed(self, op, U, phi, m0, ranks_x=1, ranks_t=1, reps=1, tol=1e-10, max_iter=10000, want_out=False):
        """op: 'dd' (reps x DD^dagger), 'cg' (one solve), 'cgiter' (reps CG-iteration bodies).
        -> (seconds, count, out|None); count = DD^dagger applications performed."""
        code = {"dd": 0, "cg": 1, "cgiter": 2}[op]
        U, phi = self._cc(U), self._cc(phi)
        out = self._c() if want_out else None
        sec = C.c_double(0)
        its = C.c_int(0)
        rc = self.lib.ref_timed(int(ranks_x), int(ranks_t), code, _d(U), _d(phi), C.c_double(m0), int(reps),
                                C.c_double(tol), int(max_iter), _d(out) if want_out else None, C.byref(sec),
                                C.byref(its))
        if rc != 0:
            raise RuntimeError("reference multi-rank run failed")
        return sec.value, its.value, out
