"""CPU model of the data routing of k_cg_cols (schwingermodel_b200/csrc/sm_column_cg.cuh).

The kernel gives every thread a column segment (S rows at one t) and gets a site's four neighbours from three places:
the thread's own slots, a warp shuffle, or -- for the segment's two end rows, lanes 0 / 31 and the row wrap -- L2 buffers
that only the matching neighbours publish into, fetched by "loader" lanes and handed over through a 32-entry line per
warp.  This file restates that routing thread by thread in numpy (same predicates, same index arithmetic, unpublished
buffer entries are NaN) and checks one stencil application against the oracle on many shapes the GPU tests cannot all
afford: every segment length, widths that wrap in the middle of a warp, row counts that leave a short last segment,
thread counts that end in a partly empty warp.  It tests the DESIGN of the routing; the kernel itself is tested on the
GPU (tests/test_gpu_parity.py::test_column_segment_cg_every_segment_length)."""
import numpy as np
import pytest

from oracle.port import Port, gaussian_fields


def _hop(dag):
    s = 1.0 if dag else -1.0
    return dict(
        s=s,
        from_tp=lambda p0, p1: p0 + s * p1,
        from_xp=lambda p0, p1: p0 - 1j * s * p1,
        from_tm=lambda p0, p1: p0 - s * p1,
        from_xm=lambda p0, p1: p0 + 1j * s * p1,
        add_tp=lambda v: (v, s * v),
        add_xp=lambda v: (v, 1j * s * v),
        add_tm=lambda v: (v, -s * v),
        add_xm=lambda v: (v, -1j * s * v),
    )


def cols_apply_model(wx, wt, S, U, p, m0, dag):
    """One stencil application routed as k_cg_cols routes it.  U, p: (2, V) complex, site n = x wt + t."""
    V = wx * wt
    H = _hop(dag)
    groups = (wx + S - 1) // S
    nthreads = groups * wt
    nthreads_padded = (nthreads + 31) // 32 * 32          # dead lanes of the last warp take part in ballots / shuffles
    mass = m0 + 2.0

    # per-thread geometry (lines "const int groups ... m_xp" of the kernel)
    th = []
    for c in range(nthreads_padded):
        live = c < nthreads
        g = c // wt if live else 0
        t = c - g * wt if live else 0
        nslots = min(S, wx - g * S) if live else 0
        nb = g * S * wt + t
        lane = c & 31
        th.append(dict(live=live, g=g, t=t, nslots=nslots, nb=nb, lane=lane,
                       tp_l2=live and (lane == 31 or t == wt - 1), tm_l2=live and (lane == 0 or t == 0),
                       m_tp0=(nb - (wt - 1)) if t == wt - 1 else nb + 1,
                       m_tm0=(nb + (wt - 1)) if t == 0 else nb - 1,
                       m_xm=(nb - wt) % V, m_xp=(nb + nslots * wt) % V))

    def link(c, mu, j):          # -1/2 (sign) U, zero in an empty slot
        T = th[c]
        if j >= T["nslots"]:
            return 0.0
        n = T["nb"] + j * wt
        sg = -0.5 * (-1.0 if (mu == 0 and T["t"] == wt - 1) else 1.0)
        return sg * U[mu, n]

    def val(c, j):               # the thread's register copy of p, zero in an empty slot
        T = th[c]
        if j >= T["nslots"]:
            return 0.0, 0.0
        n = T["nb"] + j * wt
        return p[0, n], p[1, n]

    # publish: what reaches the L2 buffers (everything else stays NaN and must never be read)
    hop = np.full((4, V), np.nan + 1j * np.nan)
    for c in range(nthreads):
        T = th[c]
        ns, nb = T["nslots"], T["nb"]
        hop[2, nb] = H["from_xp"](*val(c, 0))
        hop[3, nb + (ns - 1) * wt] = np.conj(link(c, 1, ns - 1)) * H["from_xm"](*val(c, ns - 1))
        for j in range(ns):
            if T["tm_l2"]:
                hop[0, nb + j * wt] = H["from_tp"](*val(c, j))
            if T["tp_l2"]:
                hop[1, nb + j * wt] = np.conj(link(c, 0, j)) * H["from_tm"](*val(c, j))

    out = np.full((2, V), np.nan + 1j * np.nan)
    for w0 in range(0, nthreads_padded, 32):
        warp = th[w0:w0 + 32]
        mask_tp = [T["lane"] for T in warp if T["tp_l2"]]
        mask_tm = [T["lane"] for T in warp if T["tm_l2"]]
        assert len(mask_tp) <= 2 and len(mask_tm) <= 2, "more edge readers than ports (needs width_t >= 32)"
        # loader lanes fill the staging line: lane q S + j loads slot j of port q
        stage = np.full(32, np.nan + 1j * np.nan)
        for lane in range(32):
            q, j = divmod(lane, S)
            if q >= 4:
                continue
            ports = mask_tp if q < 2 else mask_tm
            if (q & 1) >= len(ports):
                continue
            src = warp[ports[q & 1]]
            if j < src["nslots"]:
                site = (src["m_tp0"] if q < 2 else src["m_tm0"]) + j * wt
                stage[lane] = hop[0 if q < 2 else 1, site]
        for T in warp:
            if not T["live"]:
                continue
            c = w0 + T["lane"]
            my_tp = S * sum(1 for l in mask_tp if l < T["lane"])
            my_tm = S * (2 + sum(1 for l in mask_tm if l < T["lane"]))
            in_xm, in_xp = hop[3, T["m_xm"]], hop[2, T["m_xp"]]
            for j in range(T["nslots"]):
                q0, q1 = val(c, j)
                u0, u1 = link(c, 0, j), link(c, 1, j)
                a0, a1 = mass * q0, mass * q1
                # +t: the next lane's value by shuffle, or what a loader lane fetched
                if T["tp_l2"]:
                    h_tp = stage[my_tp + j]
                else:
                    assert T["lane"] < 31 and th[c + 1]["g"] == T["g"] and th[c + 1]["t"] == T["t"] + 1
                    h_tp = H["from_tp"](*val(c + 1, j))
                if T["tm_l2"]:
                    h_tm = stage[my_tm + j]
                else:
                    assert T["lane"] > 0 and th[c - 1]["g"] == T["g"] and th[c - 1]["t"] == T["t"] - 1
                    h_tm = np.conj(link(c - 1, 0, j)) * H["from_tm"](*val(c - 1, j))
                # +x / -x: the thread's own neighbouring slot, or the row beyond the segment through L2
                h_xp = in_xp if j == T["nslots"] - 1 else H["from_xp"](*val(c, j + 1))
                h_xm = in_xm if j == 0 else np.conj(link(c, 1, j - 1)) * H["from_xm"](*val(c, j - 1))
                for add, v in (("add_tp", u0 * h_tp), ("add_xp", u1 * h_xp), ("add_tm", h_tm), ("add_xm", h_xm)):
                    d0, d1 = H[add](v)
                    a0, a1 = a0 + d0, a1 + d1
                n = T["nb"] + j * wt
                out[0, n], out[1, n] = a0, a1
    return out


SHAPES = [(wx, wt) for wt in (32, 33, 40, 63, 64, 65, 96) for wx in (2, 3, 5, 8, 9, 16, 17)]


@pytest.mark.parametrize("S", [1, 2, 3, 4, 5, 6, 7, 8])
def test_column_routing_reproduces_the_stencil(S):
    for wx, wt in SHAPES:
        P = Port(wx, wt)
        U = P.hot_start(100 + wx + wt)
        p, _ = gaussian_fields(wx, wt, 7 + S)
        for dag in (False, True):
            got = cols_apply_model(wx, wt, S, U, p, -0.07, dag)
            want = P.D(U, p, -0.07, dag)
            assert not np.isnan(got).any(), (wx, wt, S, dag)
            err = np.abs(got - want).max() / np.abs(want).max()
            assert err <= 1e-13, (wx, wt, S, dag, err)


def test_column_routing_l2_share():
    """512 x 512 with 8 rows per thread: the share of half-spinors that still goes through L2 (the reason the kernel is
    not L2-bound): 2 row-end halves per thread and stencil + the t halves of 1/16 of the lanes, against 4 per site."""
    wx = wt = 512
    S = 8
    groups = (wx + S - 1) // S
    threads = groups * wt
    row_end = 2 * threads
    t_edges = 2 * (threads // 32 + groups) * S        # lanes 0 / 31 of every warp (+ the row wrap, here on warp edges)
    assert (row_end + t_edges) / (4 * wx * wt) < 0.11
