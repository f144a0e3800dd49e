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
