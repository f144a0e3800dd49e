"""Live check of the C restatement against oracle/_ref (the unmodified reference compiled from
/root/reference).  Skipped where neither the prebuilt library nor the reference tree exists."""
import numpy as np
import pytest

from oracle import ref as refmod
from oracle.port import Port, gaussian_fields

SIZES = [(8, 8), (16, 24)]


@pytest.mark.parametrize("nx,nt", SIZES)
def test_port_equals_reference_on_fresh_seeds(nx, nt):
    if not refmod.available(nx, nt):
        pytest.skip("no reference build")
    R, P = refmod.Ref(nx, nt), Port(nx, nt)
    U = R.hot_start(99)
    assert np.array_equal(U, P.hot_start(99))
    chi, pi = gaussian_fields(nx, nt, 5)
    phi, _ = gaussian_fields(nx, nt, 6)
    m0, beta = -0.1, 3.0
    assert np.array_equal(R.DDdag(U, phi, m0), P.DDdag(U, phi, m0))
    xa, oka, na, _ = R.cg(U, phi, m0)
    xb, okb, nb, _ = P.cg(U, phi, m0)
    assert (oka, na) == (okb, nb) and np.array_equal(xa, xb)
    ta = R.trajectory(U, pi, chi, 5, 0.5, beta, m0)
    tb = P.trajectory(U, pi, chi, 5, 0.5, beta, m0)
    assert ta["dH"] == tb["dH"] and np.array_equal(ta["U"], tb["U"]) and np.array_equal(ta["pi"], tb["pi"])


def test_reference_multirank_equals_single_rank():
    """The reference's own halo branches (forked ranks over minimpi) agree with its serial branch,
    so the single-rank oracle is valid for every decomposition."""
    nx, nt = 16, 24
    if not refmod.available(nx, nt):
        pytest.skip("no reference build")
    R = refmod.Ref(nx, nt)
    U = R.hot_start(12345)
    phi, _ = gaussian_fields(nx, nt, 778)
    one = R.DDdag(U, phi, -0.05)
    x1, _, n1, _ = R.cg(U, phi, -0.05)
    for rx, rt in [(2, 2), (4, 2), (1, 2), (2, 1)]:
        _, _, out = R.timed("dd", U, phi, -0.05, rx, rt, reps=1, want_out=True)
        assert np.array_equal(out, one), (rx, rt)
        _, n, x = R.timed("cg", U, phi, -0.05, rx, rt, want_out=True)
        assert n == n1 and np.abs(x - x1).max() < 1e-13 * np.abs(x1).max()
