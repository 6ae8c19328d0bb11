"""Host-side tables of the exact network condensation for general polynomial degrees
(flux P_fd, pressure DG0 or continuous P_pd; assembly.py:121-146) -- the analysis phase of the direct
solve the reference obtains from MUMPS (solver.py:58-65).

Everything that lives on ONE graph edge is *local* to it: its ``fd N + 1`` flux dofs, the pressure
dofs inside it (cells, interior vertices, cell-interior dofs) and the pressure dof of a boundary node
at its end.  What couples edges sits at the bifurcations: the multiplier ``lam_b`` and, for a
continuous pressure, the nodal pressure ``P_b``.  Per edge ``e = (u, v)`` with nodal unknowns
``z_e = (P_u, lam_u, P_v, lam_v)``::

    K_e y + C_e z_e = r_loc          (rows of the local unknowns)
    sum_e D_e y_e   = r_z            (rows of the nodal unknowns; no nodal-nodal entries, assembly.py:284-287)

``K_e`` is a banded saddle matrix when the local unknowns are ordered by their position along the edge.
The device (``condense.cuh``) factorises it per edge (banded LU, partial pivoting), forms
``Y_e = K_e^{-1} C_e`` and the 4 x 4 Schur contribution ``S_e = -D_e Y_e``; the bifurcation system with
2 x 2 blocks has the network's own topology and is eliminated leaf -> root without fill on a tree
(the same schedule as the P1/DG0 path).  This module only lists *which* entries ``K_e, C_e, D_e`` have,
per edge type ``(u is a bifurcation) + 2 (v is a bifurcation)``; the values are ``coef`` or
``coef * (R h)_cell``, taken from the same reference-element tables as the assembly (elements.py).
"""

from __future__ import annotations

import dataclasses

import numpy as np

from . import elements

# kinds of local unknowns -> global dof index on the device
K_FLUX, K_PCELL, K_PVERT, K_PU, K_PV = 0, 1, 2, 3, 4
# nodal slots of an edge
Z_PU, Z_LU, Z_PV, Z_LV = 0, 1, 2, 3


@dataclasses.dataclass
class EdgeTypeTables:
    n: int
    loc_kind: np.ndarray
    loc_off: np.ndarray
    k_row: np.ndarray
    k_col: np.ndarray
    k_cell: np.ndarray  # cell of the edge whose R*h scales the entry, or -1 (constant)
    k_coef: np.ndarray
    c_row: np.ndarray
    c_slot: np.ndarray
    c_coef: np.ndarray
    d_slot: np.ndarray
    d_col: np.ndarray
    d_coef: np.ndarray


def edge_type_tables(N: int, fd: int, pd: int, tu: bool, tv: bool) -> EdgeTypeTables:
    M, B, _w, _t0, _t1 = elements.tables(fd, pd)
    loc = []
    qv = [None] * (N + 1)
    qi = [[None] * (fd - 1) for _ in range(N)]
    pv = [None] * (N + 1)  # local index, or ("z", slot) for a nodal pressure
    pi = [[] for _ in range(N)]
    for j in range(N + 1):
        qv[j] = len(loc)
        loc.append((K_FLUX, j))
        if pd >= 1:
            if 0 < j < N:
                pv[j] = len(loc)
                loc.append((K_PVERT, j - 1))
            elif j == 0:
                if tu:
                    pv[0] = ("z", Z_PU)
                else:
                    pv[0] = len(loc)
                    loc.append((K_PU, 0))
            else:
                if tv:
                    pv[N] = ("z", Z_PV)
                else:
                    pv[N] = len(loc)
                    loc.append((K_PV, 0))
        if j < N:
            for i in range(fd - 1):
                qi[j][i] = len(loc)
                loc.append((K_FLUX, (N + 1) + j * (fd - 1) + i))
            if pd == 0:
                pi[j] = [len(loc)]
                loc.append((K_PCELL, j))
            else:
                for i in range(pd - 1):
                    pi[j].append(len(loc))
                    loc.append((K_PCELL, j * (pd - 1) + i))
    K, Cc, Dd = [], [], []
    for j in range(N):
        lq = [qv[j], qv[j + 1], *qi[j]]  # element dof order [X=0, X=1, interior]
        lp = pi[j] if pd == 0 else [pv[j], pv[j + 1], *pi[j]]
        for a in range(fd + 1):
            for b in range(fd + 1):
                K.append((lq[a], lq[b], j, M[a, b]))  # assembly.py:253
        for r, pr in enumerate(lp):
            for a in range(fd + 1):
                if isinstance(pr, tuple):
                    Dd.append((pr[1], lq[a], B[r, a]))  # assembly.py:254, nodal pressure row
                    Cc.append((lq[a], pr[1], -B[r, a]))  # assembly.py:255
                else:
                    K.append((pr, lq[a], -1, B[r, a]))
                    K.append((lq[a], pr, -1, -B[r, a]))
    if tv:  # in-edge of v: +1 at the last vertex (assembly.py:271-277)
        Dd.append((Z_LV, qv[N], 1.0))
        Cc.append((qv[N], Z_LV, 1.0))
    if tu:  # out-edge of u: -1 at the first vertex
        Dd.append((Z_LU, qv[0], -1.0))
        Cc.append((qv[0], Z_LU, -1.0))
    i32, f64 = np.int32, np.float64
    col = lambda rows, k, dt: np.asarray([r[k] for r in rows], dtype=dt)  # noqa: E731
    return EdgeTypeTables(
        len(loc), col(loc, 0, i32), col(loc, 1, i32),
        col(K, 0, i32), col(K, 1, i32), col(K, 2, i32), col(K, 3, f64),
        col(Cc, 0, i32), col(Cc, 1, i32), col(Cc, 2, f64),
        col(Dd, 0, i32), col(Dd, 1, i32), col(Dd, 2, f64),
    )


@dataclasses.dataclass
class Condensation:
    fd: int
    pd: int
    N: int
    n_max: int
    kl: int  # half bandwidth of K_e (sub- = super-diagonals before pivoting)
    pcell_base: int
    pcell_stride: int
    types: list  # 4 EdgeTypeTables, index = tu + 2 tv
    bif_node: np.ndarray

    def packed(self):
        """Concatenated tables in the argument order of ``nxfx_set_condensation``."""
        i32 = np.int32
        ptr = lambda key: np.concatenate([[0], np.cumsum([getattr(t, key).size for t in self.types])]).astype(i32)  # noqa: E731
        cat = lambda key: np.ascontiguousarray(np.concatenate([getattr(t, key) for t in self.types]))  # noqa: E731
        return dict(
            type_n=np.asarray([t.n for t in self.types], dtype=i32),
            loc_ptr=ptr("loc_kind"), loc_kind=cat("loc_kind"), loc_off=cat("loc_off"),
            k_ptr=ptr("k_row"), k_row=cat("k_row"), k_col=cat("k_col"), k_cell=cat("k_cell"), k_coef=cat("k_coef"),
            c_ptr=ptr("c_row"), c_row=cat("c_row"), c_slot=cat("c_slot"), c_coef=cat("c_coef"),
            d_ptr=ptr("d_slot"), d_slot=cat("d_slot"), d_col=cat("d_col"), d_coef=cat("d_coef"),
        )


def build_condensation(nm, flux_degree: int, pressure_degree: int) -> Condensation:
    fd, pd, N = int(flux_degree), int(pressure_degree), int(nm.cells_per_edge)
    types = [edge_type_tables(N, fd, pd, bool(t & 1), bool(t & 2)) for t in range(4)]
    kl = max(int(np.abs(t.k_row - t.k_col).max()) for t in types)
    E = nm.graph_edges.shape[0]
    n_nodes = nm._n_nodes
    nq = E * (fd * N + 1)
    nv = n_nodes + (N - 1) * E
    if pd == 0:
        base, stride = nq, N
    else:
        base, stride = nq + nv, N * (pd - 1)
    return Condensation(fd, pd, N, max(t.n for t in types), kl, base, stride, types,
                        np.ascontiguousarray(nm.bifurcation_values, dtype=np.int32))
