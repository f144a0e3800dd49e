"""Tiles of a split lattice: which sites of the global (2, Nx*Nt) field a rank owns.

Same placement as the reference: rank r = (cx, ct) = (r // ranks_t, r % ranks_t) owns rows
[cx*width_x, (cx+1)*width_x) x columns [ct*width_t, (ct+1)*width_t), local site
n = x_local*width_t + t_local (include/mpi_setup.h:39-47, src/gauge_conf.cpp:383-387)."""
from __future__ import annotations

import numpy as np


def tile_shape(Nx, Nt, ranks_x, ranks_t):
    if Nx % ranks_x or Nt % ranks_t:
        raise ValueError("Nx (Nt) is not exactly divisible by rank_x (rank_t)")
    return Nx // ranks_x, Nt // ranks_t


def tile_of(field, Nx, Nt, ranks_x, ranks_t, rank):
    """(C, Nx*Nt) global field -> (C, width_x*width_t) tile of `rank` (a copy)."""
    wx, wt = tile_shape(Nx, Nt, ranks_x, ranks_t)
    cx, ct = divmod(rank, ranks_t)
    f = np.asarray(field).reshape(field.shape[0], Nx, Nt)
    return np.ascontiguousarray(f[:, cx * wx:(cx + 1) * wx, ct * wt:(ct + 1) * wt]).reshape(field.shape[0], wx * wt)


def assemble(tiles, Nx, Nt, ranks_x, ranks_t):
    """list of per-rank tiles (rank order) -> global field."""
    wx, wt = tile_shape(Nx, Nt, ranks_x, ranks_t)
    C = tiles[0].shape[0]
    out = np.empty((C, Nx, Nt), dtype=tiles[0].dtype)
    for r, t in enumerate(tiles):
        cx, ct = divmod(r, ranks_t)
        out[:, cx * wx:(cx + 1) * wx, ct * wt:(ct + 1) * wt] = np.asarray(t).reshape(C, wx, wt)
    return out.reshape(C, Nx * Nt)


# ---- what travels between tiles (the conventions of csrc/sm_ops.cuh / sm_dist.cuh, stated once in Python so that the
#      CPU tests can carry the exchange over gloo and check it against the single-rank oracle) -----------------------------
def neighbours(ranks_x, ranks_t, rank):
    """Periodic Cartesian neighbours of `rank` (include/mpi_setup.h:39-71): x-1 "top", x+1 "bot", t-1 "left", t+1 "right"."""
    cx, ct = divmod(rank, ranks_t)

    def rk(a, b):
        return (a % ranks_x) * ranks_t + b % ranks_t
    return {"xm": rk(cx - 1, ct), "xp": rk(cx + 1, ct), "tm": rk(cx, ct - 1), "tp": rk(cx, ct + 1)}


def seam_signs(ranks_t, rank):
    """(sR_edge, sL_edge): the antiperiodic sign of the fermions in t multiplies the hop across the GLOBAL seam only, i.e. the
    +t hop out of the last column of tiles with ct = ranks_t-1 and the -t hop into the first column of tiles with ct = 0
    (include/dirac_operator.h:53-58)."""
    ct = rank % ranks_t
    return (-1.0 if ct == ranks_t - 1 else 1.0), (-1.0 if ct == 0 else 1.0)


def boundary_rows2(tile, wx, wt):
    """The two boundary rows of each side of a (C, wx*wt) tile, layout [component][2 rows][wt] (contiguous in HBM, sent as they
    are): (rows 0, 1 -> the -x neighbour's "hi" ghost, rows wx-2, wx-1 -> the +x neighbour's "lo" ghost)."""
    f = np.asarray(tile).reshape(tile.shape[0], wx, wt)
    return np.ascontiguousarray(f[:, :2, :]), np.ascontiguousarray(f[:, wx - 2:, :])


def boundary_cols2(tile, wx, wt):
    """The two boundary columns of each side, layout [component][wx rows][2] (strided in HBM: packed by k_pack_cols2 / stored by
    k_push_cols): (columns 0, 1 -> the -t neighbour's "hi" ghost, columns wt-2, wt-1 -> the +t neighbour's "lo" ghost)."""
    f = np.asarray(tile).reshape(tile.shape[0], wx, wt)
    return np.ascontiguousarray(f[:, :, :2]), np.ascontiguousarray(f[:, :, wt - 2:])
