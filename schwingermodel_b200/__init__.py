"""schwingermodel_b200 -- B200 (sm_100a) implementation of the Schwinger-model HMC fermion hot path.

The product is schwingermodel_b200/libschwinger_b200.so (hand-written CUDA behind the C ABI of
include/schwinger_b200.h).  This package is its Python host mirror: the reference's operator
interface (D_phi, D_dagger_phi, D_D_dagger_phi, conjugate_gradient, phi_dag_partialD_phi, the
GaugeConf observables and the HMC driver) with numpy arrays at the boundary.  No CPU fallback.
"""
from ._abi import LIB_PATH, SchwingerError, declared_symbols, load  # noqa: F401
from .lattice import HMC, DeviceField, Lattice, SaveConf, format_tag, jackknife_error, readBinary  # noqa: F401
