"""ctypes binding of libschwinger_b200.so (include/schwinger_b200.h).

The library is the product; this file only declares its symbols.  There is no fallback: if the
shared library is missing or a symbol is absent, import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libschwinger_b200.so")
HEADER_PATH = os.path.join(ROOT, "include", "schwinger_b200.h")

SM_OK, SM_ERR_ARG, SM_ERR_CUDA, SM_ERR_NCCL, SM_ERR_IO, SM_ERR_STATE = range(6)
SM_NCCL_ID_BYTES = 128
SM_P2P_HANDLE_BYTES = 64
SM_SOLVER_REFERENCE, SM_SOLVER_MIXED, SM_SOLVER_CHRONO, SM_SOLVER_EVENODD = 0, 1, 2, 3

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
ctx_p = C.c_void_p


class HmcParams(C.Structure):
    _fields_ = [("beta", C.c_double), ("m0", C.c_double), ("trajectory_length", C.c_double), ("md_steps", C.c_int)]


class TrajResult(C.Structure):
    _fields_ = [
        ("dH", C.c_double), ("H_old", C.c_double), ("H_new", C.c_double),
        ("sum_re_plaq_new", C.c_double), ("gauge_action_new", C.c_double),
        ("sum_re_plaq_old", C.c_double), ("gauge_action_old", C.c_double),
        ("dd_applications", C.c_longlong), ("cg_solves", C.c_int), ("cg_all_converged", C.c_int),
        ("kernel_ms", C.c_double), ("cg_force_failures", C.c_int), ("reserved_", C.c_int),
    ]


class SchwingerError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"libschwinger_b200 error {code}: {text}")
        self.code = code


def declared_symbols(header: str = HEADER_PATH):
    """Every entry point include/schwinger_b200.h declares (used by the CPU export test)."""
    with open(header) as f:
        return re.findall(r"^SM_API\s+(?:const\s+char\*|int)\s+(sm_\w+)\s*\(", f.read(), flags=re.M)


_SIGS = {
    "sm_create": [C.c_int, C.c_int, C.c_int, C.POINTER(ctx_p)],
    "sm_create_dist": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(ctx_p)],
    "sm_nccl_unique_id": [C.c_void_p],
    "sm_p2p_handle": [ctx_p, C.c_void_p],
    "sm_p2p_connect": [ctx_p, C.c_void_p],
    "sm_destroy": [ctx_p],
    "sm_local_dims": [ctx_p, ip],
    "sm_set_cg": [ctx_p, C.c_double, C.c_int],
    "sm_set_solver": [ctx_p, C.c_int],
    "sm_last_kernel_ms": [ctx_p, dp],
    "sm_launch_count": [ctx_p, C.POINTER(C.c_longlong)],
    "sm_one_pass_dd": [ctx_p, ip],
    "sm_peer_mode": [ctx_p, ip],
    "sm_device_count": [ip],
    "sm_host_register": [C.c_int],
    "sm_host_forget": [C.c_void_p],
    "sm_tables": [ctx_p, C.c_int, C.c_int, C.c_int, ip, ip, dp, dp, ip, ip],
    "sm_D_phi": [ctx_p, dp, dp, dp, dp, dp, dp, C.c_double],
    "sm_D_dagger_phi": [ctx_p, dp, dp, dp, dp, dp, dp, C.c_double],
    "sm_D_D_dagger_phi": [ctx_p, dp, dp, dp, dp, dp, dp, C.c_double],
    "sm_dot": [ctx_p, dp, dp, dp, dp, dp],
    "sm_conjugate_gradient": [ctx_p, dp, dp, dp, dp, dp, dp, C.c_double, ip, ip],
    "sm_evenodd_solve": [ctx_p, dp, dp, dp, dp, dp, dp, C.c_double, ip, ip],
    "sm_phi_dag_partialD_phi": [ctx_p, dp, dp, dp, dp, dp, dp, dp, dp],
    "sm_compute_staple": [ctx_p, dp, dp, dp, dp],
    "sm_compute_plaquette": [ctx_p, dp, dp, C.c_double, dp, dp],
    "sm_field_alloc": [ctx_p, C.c_int, C.POINTER(dp)],
    "sm_field_free": [ctx_p, dp],
    "sm_field_upload": [ctx_p, dp, dp, dp, C.c_int],
    "sm_field_download": [ctx_p, dp, dp, dp, C.c_int],
    "sm_dev_D": [ctx_p, dp, dp, dp, C.c_double, C.c_int],
    "sm_dev_DDdag": [ctx_p, dp, dp, dp, C.c_double],
    "sm_dev_dot": [ctx_p, dp, dp, dp],
    "sm_dev_cg": [ctx_p, dp, dp, dp, C.c_double, ip, ip],
    "sm_dev_DDdag_loop": [ctx_p, dp, dp, dp, C.c_double, C.c_int, dp],
    "sm_hmc_configure": [ctx_p, C.POINTER(HmcParams)],
    "sm_hmc_set_gauge": [ctx_p, dp, dp],
    "sm_hmc_get_gauge": [ctx_p, dp, dp, C.c_int],
    "sm_hmc_get_momenta": [ctx_p, dp, dp, C.c_int],
    "sm_hmc_get_phi": [ctx_p, dp, dp],
    "sm_hmc_get_chi": [ctx_p, dp, dp],
    "sm_hmc_refresh": [ctx_p, C.c_uint64, C.c_uint64],
    "sm_hmc_inject": [ctx_p, dp, dp, dp, dp],
    "sm_hmc_trajectory": [ctx_p, C.POINTER(TrajResult)],
    "sm_hmc_accept": [ctx_p, C.c_int],
    "sm_hmc_force": [ctx_p, dp, dp, dp, dp, ip],
    "sm_hmc_hamiltonian": [ctx_p, dp, dp, dp, dp, dp],
    "sm_hmc_leapfrog": [ctx_p, dp, dp, dp, dp, ip],
    "sm_save_conf": [C.c_int, C.c_int, dp, dp, C.c_char_p],
    "sm_read_conf": [C.c_int, C.c_int, C.c_char_p, dp, dp],
}

_lib = None


def load():
    """Load the shared library and type every symbol.  Raises if the CUDA extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with schwingermodel_b200/csrc/build.sh "
            "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    lib.sm_last_error.restype = C.c_char_p
    lib.sm_last_error.argtypes = []
    for name, args in _SIGS.items():
        fn = getattr(lib, name)   # AttributeError if the library lacks a declared symbol
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int):
    if rc != SM_OK:
        raise SchwingerError(rc, load().sm_last_error().decode(errors="replace"))
