"""Host-side symbolic phase of the table-driven (higher-order) assembly path.

For ``(flux_degree, pressure_degree) != (1, 0)`` the matrix is described by *contribution lists*:
every stored entry has at most two sources ``(cell, coefficient, scale)`` with value
``coefficient * (R h of the cell | 1)``, the right-hand side rows have short source lists
(``f h w_r`` per cell, ``+-p_bc`` at boundary vertices).  The device kernels
(``generic.cuh``) gather them in a fixed order -- no atomics, deterministic -- which is the
"cell-to-nnz map" of the north_star in gather form.  Numbering: SURVEY Appendix C (flux dofs of an
edge = N+1 vertex dofs then the interior dofs cell by cell; continuous pressure = mesh vertices
then interior dofs; DG0 pressure = cells).
"""

from __future__ import annotations

import dataclasses

import numpy as np

from . import elements

RH_FLAG = 1 << 30  # source value is coefficient * R*h of the cell
VERTEX_FLAG = 1 << 30  # rhs source is coefficient * p_bc(vertex)


@dataclasses.dataclass
class GenericSystem:
    n_dofs: int
    n_flux: int
    n_pressure: int
    rowptr: np.ndarray
    colidx: np.ndarray
    src_id: np.ndarray  # [nnz, 2] int32: cell | RH_FLAG, or -1
    src_coef: np.ndarray  # [nnz, 2] float64
    bsrc_ptr: np.ndarray
    bsrc_id: np.ndarray  # cell, or vertex | VERTEX_FLAG
    bsrc_coef: np.ndarray
    block_sizes: list
    cell_flux_dofs: np.ndarray
    cell_pressure_dofs: np.ndarray
    flux_per_edge: int


def build_generic_system(nm, flux_degree: int, pressure_degree: int) -> GenericSystem:
    fd, pd = int(flux_degree), int(pressure_degree)
    M, B, w, t0, t1 = elements.tables(fd, pd)
    edges = nm.graph_edges
    E, N = edges.shape[0], nm.cells_per_edge
    nc = E * N
    n_nodes = nm._n_nodes
    nv = n_nodes + (N - 1) * E
    per_edge = fd * N + 1
    fb = nm.edge_slot.astype(np.int64) * per_edge
    nq = E * per_edge
    n_p = nc if pd == 0 else nv + (pd - 1) * nc
    loff = nq + n_p
    lm = nm.node_multiplier_index.astype(np.int64)
    n = loff + nm.bifurcation_values.size
    e = np.repeat(np.arange(E), N)
    j = np.tile(np.arange(N), E)
    cell = np.arange(nc)
    qd = [fb[e] + j, fb[e] + j + 1] + [fb[e] + (N + 1) + j * (fd - 1) + i for i in range(fd - 1)]
    qd = np.stack(qd, axis=1)
    if pd == 0:
        pdofs = (nq + cell)[:, None]
    else:
        cells_v = nm._cells()
        pdofs = np.stack([nq + cells_v[:, 0], nq + cells_v[:, 1]]
                         + [nq + nv + cell * (pd - 1) + i for i in range(pd - 1)], axis=1)
    rows, cols, ids, coefs = [], [], [], []

    def add(r, c, cells_, coef, rh):
        rows.append(np.asarray(r)); cols.append(np.asarray(c))
        ids.append(np.asarray(cells_) | (RH_FLAG if rh else 0))
        coefs.append(np.broadcast_to(np.float64(coef), np.asarray(r).shape))

    for a in range(fd + 1):
        for b_ in range(fd + 1):
            add(qd[:, a], qd[:, b_], cell, M[a, b_], True)  # assembly.py:253
    for r in range(pd + 1):
        for a in range(fd + 1):
            add(pdofs[:, r], qd[:, a], cell, B[r, a], False)  # assembly.py:254
            add(qd[:, a], pdofs[:, r], cell, -B[r, a], False)  # assembly.py:255
    u, v = edges[:, 0], edges[:, 1]
    ein, eout = np.flatnonzero(lm[v] >= 0), np.flatnonzero(lm[u] >= 0)
    last, first = ein * N + N - 1, eout * N
    lin, lout = loff + lm[v[ein]], loff + lm[u[eout]]
    for a in range(fd + 1):  # assembly.py:271-277, whole cell rows (explicit zeros)
        add(lin, qd[last, a], last, t1[a], False)
        add(qd[last, a], lin, last, t1[a], False)
        add(lout, qd[first, a], first, -t0[a], False)
        add(qd[first, a], lout, first, -t0[a], False)
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    ids, coefs = np.concatenate(ids), np.concatenate(coefs)
    order = np.lexsort((ids & (RH_FLAG - 1), cols, rows))
    rows, cols, ids, coefs = rows[order], cols[order], ids[order], coefs[order]
    newent = np.ones(rows.size, dtype=bool)
    newent[1:] = (rows[1:] != rows[:-1]) | (cols[1:] != cols[:-1])
    ent = np.cumsum(newent) - 1
    nnz = int(ent[-1]) + 1
    pos = np.arange(rows.size) - np.flatnonzero(newent)[ent]
    if pos.max() > 1:
        raise AssertionError("an entry with more than two cell contributions")
    src_id = np.full((nnz, 2), -1, dtype=np.int32)
    src_coef = np.zeros((nnz, 2))
    src_id[ent, pos] = ids
    src_coef[ent, pos] = coefs
    colidx = cols[newent].astype(np.int32)
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(rows[newent], minlength=n))]).astype(np.int32)
    # right-hand side sources
    brow = [pdofs[:, r] for r in range(pd + 1)]
    bid = [cell for _ in range(pd + 1)]
    bco = [np.full(nc, w[r]) for r in range(pd + 1)]
    outlet, inlet = np.flatnonzero(lm[v] < 0), np.flatnonzero(lm[u] < 0)
    brow += [fb[outlet] + N, fb[inlet]]
    bid += [v[outlet] | VERTEX_FLAG, u[inlet] | VERTEX_FLAG]
    bco += [np.ones(outlet.size), -np.ones(inlet.size)]
    brow, bid, bco = np.concatenate(brow), np.concatenate(bid), np.concatenate(bco)
    order = np.lexsort((bid, brow))
    brow, bid, bco = brow[order], bid[order], bco[order]
    bptr = np.concatenate([[0], np.cumsum(np.bincount(brow, minlength=n))]).astype(np.int32)
    counts = nm._color_count
    return GenericSystem(
        n, nq, n_p, rowptr, colidx, src_id, src_coef, bptr, bid.astype(np.int32), bco,
        [int(c) * per_edge for c in counts] + [n_p, int(nm.bifurcation_values.size)], qd, pdofs, per_edge,
    )
