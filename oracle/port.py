"""ctypes loader for oracle/liboracle.so — the C restatement of the reference's hot path.

TEST INFRASTRUCTURE ONLY (see the header of schwinger_oracle.c).  Same method names and array
conventions as oracle.ref.Ref, but the lattice size is a constructor argument, not a build.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "schwinger_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(src) > os.path.getmtime(_LIB):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return _LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        for f in ("so_action", "so_hamiltonian", "so_jackknife"):
            getattr(_lib, f).restype = C.c_double
    return _lib


class Port:
    def __init__(self, nx: int, nt: int):
        self.lib = _load()
        self.nx, self.nt, self.V = nx, nt, nx * nt

    def _c(self):
        return np.zeros((2, self.V), dtype=np.complex128)

    def _r(self):
        return np.zeros((2, self.V), dtype=np.float64)

    @staticmethod
    def _cc(a):
        return np.ascontiguousarray(a, dtype=np.complex128)

    def tables(self, ranks_x=1, ranks_t=1, rank=0):
        m = self.V // (ranks_x * ranks_t)
        rpb = np.zeros(2 * m, np.int32)
        lpb = np.zeros(2 * m, np.int32)
        sr = np.zeros(2 * m, np.complex128)
        sl = np.zeros(2 * m, np.complex128)
        a = np.zeros(m, np.int32)
        b = np.zeros(m, np.int32)
        self.lib.so_tables(self.nx, self.nt, ranks_x, ranks_t, rank, _i(rpb), _i(lpb), _d(sr), _d(sl), _i(a), _i(b))
        return dict(RightPB=rpb, LeftPB=lpb, SignR=sr, SignL=sl, x_1_t1=a, x1_t_1=b)

    def hot_start(self, seed: int):
        U = self._c()
        self.lib.so_hot_start(C.c_uint(seed), self.nx, self.nt, _d(U))
        return U

    def D(self, U, phi, m0, dagger=False):
        U, phi, out = self._cc(U), self._cc(phi), self._c()
        self.lib.so_D(self.nx, self.nt, _d(U), _d(phi), _d(out), C.c_double(m0), int(dagger))
        return out

    def DDdag(self, U, phi, m0):
        U, phi, out = self._cc(U), self._cc(phi), self._c()
        self.lib.so_DDdag(self.nx, self.nt, _d(U), _d(phi), _d(out), C.c_double(m0))
        return out

    def dot(self, x, y):
        x, y = self._cc(x), self._cc(y)
        o = np.zeros(2)
        self.lib.so_dot(self.nx, self.nt, _d(x), _d(y), _d(o))
        return complex(o[0], o[1])

    def cg(self, U, phi, m0, tol=1e-10, max_iter=10000):
        """-> (x, converged, DD^dagger applications, seconds=None)"""
        U, phi, x = self._cc(U), self._cc(phi), self._c()
        apps = C.c_int(0)
        ok = self.lib.so_cg(self.nx, self.nt, _d(U), _d(phi), _d(x), C.c_double(m0), C.c_double(tol), int(max_iter),
                            C.byref(apps))
        return x, int(ok), apps.value, None

    def fermion_force(self, U, left, right):
        U, left, right, F = self._cc(U), self._cc(left), self._cc(right), self._r()
        self.lib.so_fermion_force(self.nx, self.nt, _d(U), _d(left), _d(right), _d(F))
        return F

    def staple(self, U):
        U, K = self._cc(U), self._c()
        self.lib.so_staple(self.nx, self.nt, _d(U), _d(K))
        return K

    def plaquette(self, U, beta=1.0):
        U = self._cc(U)
        P = np.zeros(self.V, np.complex128)
        s = np.zeros(2)
        self.lib.so_plaquette(self.nx, self.nt, _d(U), C.c_double(beta), _d(P), _d(s))
        return P, float(s[0]), float(s[1])

    def force(self, U, phi, beta, m0, tol=1e-10, max_iter=10000):
        U, phi, F = self._cc(U), self._cc(phi), self._r()
        ok = self.lib.so_force(self.nx, self.nt, _d(U), _d(phi), C.c_double(beta), C.c_double(m0), C.c_double(tol),
                               int(max_iter), _d(F))
        return F, int(ok)

    def action(self, U, phi, beta, m0, tol=1e-10, max_iter=10000):
        U, phi = self._cc(U), self._cc(phi)
        return float(self.lib.so_action(self.nx, self.nt, _d(U), _d(phi), C.c_double(beta), C.c_double(m0),
                                        C.c_double(tol), int(max_iter)))

    def hamiltonian(self, U, pi, phi, beta, m0, tol=1e-10, max_iter=10000):
        U, phi = self._cc(U), self._cc(phi)
        pi = np.ascontiguousarray(pi, np.float64)
        return float(self.lib.so_hamiltonian(self.nx, self.nt, _d(U), _d(pi), _d(phi), C.c_double(beta),
                                             C.c_double(m0), C.c_double(tol), int(max_iter)))

    def leapfrog(self, U, pi, phi, md, tau, beta, m0, tol=1e-10, max_iter=10000):
        U, phi = self._cc(U), self._cc(phi)
        pi = np.ascontiguousarray(pi, np.float64)
        Uo, po = self._c(), self._r()
        self.lib.so_leapfrog(self.nx, self.nt, _d(U), _d(pi), _d(phi), int(md), C.c_double(tau), C.c_double(beta),
                             C.c_double(m0), C.c_double(tol), int(max_iter), _d(Uo), _d(po))
        return Uo, po

    def trajectory(self, U, pi, chi, md, tau, beta, m0, tol=1e-10, max_iter=10000):
        U, chi = self._cc(U), self._cc(chi)
        pi = np.ascontiguousarray(pi, np.float64)
        phi, Uo, po = self._c(), self._c(), self._r()
        H = np.zeros(2)
        aux = np.zeros(3)
        ok = self.lib.so_trajectory(self.nx, self.nt, _d(U), _d(pi), _d(chi), int(md), C.c_double(tau),
                                    C.c_double(beta), C.c_double(m0), C.c_double(tol), int(max_iter), _d(phi), _d(Uo),
                                    _d(po), _d(H), _d(aux))
        return dict(phi=phi, U=Uo, pi=po, H_old=float(H[0]), H_new=float(H[1]), dH=float(H[1] - H[0]),
                    sp=float(aux[0]), sg=float(aux[1]), dd_apps=int(aux[2]), cg_ok=int(ok))

    def save_conf(self, U, name):
        U = self._cc(U)
        rc = self.lib.so_save_conf(self.nx, self.nt, _d(U), name.encode())
        if rc:
            raise OSError(f"cannot write {name}")

    def read_binary(self, name):
        U = self._c()
        rc = self.lib.so_read_binary(self.nx, self.nt, name.encode(), _d(U))
        if rc:
            raise OSError(f"cannot read {name} (rc={rc})")
        return U

    def jackknife(self, dat, bins):
        dat = np.ascontiguousarray(dat, np.float64)
        return float(self.lib.so_jackknife(_d(dat), len(dat), int(bins)))


def gaussian_fields(nx, nt, seed):
    """Deterministic sources shared by tests and benches: chi (re,im ~ N(0,1/sqrt2)), pi ~ N(0,1)."""
    V = nx * nt
    rng = np.random.default_rng(seed)
    chi = (rng.normal(size=(2, V)) + 1j * rng.normal(size=(2, V))) / np.sqrt(2.0)
    pi = rng.normal(size=(2, V))
    return chi, pi
