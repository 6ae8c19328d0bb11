"""CPU check of the host tables of the exact higher-order condensation (networks_fenicsx_b200/condense.py):
a NumPy emulation of what the device kernels do with them (per-edge solve, 4 x 4 Schur contribution,
2 x 2 block elimination over the schedule, back-substitution) must reproduce a direct solve of the
oracle's matrix.  Test infrastructure only -- the product has no CPU path."""

import numpy as np
import pytest
import scipy.sparse.linalg as spla

import networks_fenicsx_b200 as nxfx
from networks_fenicsx_b200 import condense
from networks_fenicsx_b200 import network_generation as ng
from networks_fenicsx_b200.schedule import build_tree_schedule
from oracle import reference_port as rp
from tests import helpers


def emulate(nm, cond, sched, cell_rh, r):
    """z = P^{-1} r exactly as condense.cuh / the block tree sweeps do it (dense per-edge algebra)."""
    edges, N, fd, pd = nm.graph_edges, cond.N, cond.fd, cond.pd
    E = edges.shape[0]
    per_edge = fd * N + 1
    nq = E * per_edge
    n_nodes = nm._n_nodes
    lm = nm.node_multiplier_index
    n_bif = nm.bifurcation_values.size
    loff = r.size - n_bif
    slot = nm.edge_slot

    def glob(e, kind, off):
        u, v = edges[e]
        if kind == condense.K_FLUX:
            return slot[e] * per_edge + off
        if kind == condense.K_PCELL:
            return cond.pcell_base + e * cond.pcell_stride + off
        if kind == condense.K_PVERT:
            return nq + n_nodes + e * (N - 1) + off
        return nq + (u if kind == condense.K_PU else v)

    S = np.zeros((E, 4, 4))
    Y, y0, gl = [], [], []
    h = np.zeros((E, 4))
    for e in range(E):
        u, v = edges[e]
        t = cond.types[int(lm[u] >= 0) + 2 * int(lm[v] >= 0)]
        K = np.zeros((t.n, t.n))
        scale = np.where(t.k_cell >= 0, cell_rh[e * N + np.maximum(t.k_cell, 0)], 1.0)
        np.add.at(K, (t.k_row, t.k_col), t.k_coef * scale)
        Cm = np.zeros((t.n, 4))
        np.add.at(Cm, (t.c_row, t.c_slot), t.c_coef)
        Dm = np.zeros((4, t.n))
        np.add.at(Dm, (t.d_slot, t.d_col), t.d_coef)
        g = np.array([glob(e, k, o) for k, o in zip(t.loc_kind, t.loc_off)])
        assert np.unique(g).size == g.size
        Ye = np.linalg.solve(K, Cm)
        ye = np.linalg.solve(K, r[g])
        S[e] = -Dm @ Ye
        h[e] = Dm @ ye
        Y.append(Ye); y0.append(ye); gl.append(g)
    # nodal blocks in schedule order
    tob = sched.t_of_bif
    D0 = np.zeros((n_bif, 2, 2))
    rz = np.zeros((n_bif, 2))
    for b in range(n_bif):
        node = nm.bifurcation_values[b]
        rz[tob[b]] = [r[nq + node] if pd >= 1 else 0.0, r[loff + b]]
    for e in range(E):
        u, v = edges[e]
        if lm[u] >= 0:
            D0[tob[lm[u]]] += S[e][0:2, 0:2]
            rz[tob[lm[u]]] -= h[e][0:2]
        if lm[v] >= 0:
            D0[tob[lm[v]]] += S[e][2:4, 2:4]
            rz[tob[lm[v]]] -= h[e][2:4]
    if pd == 0:
        D0[:, 0, 0] = 1.0
    U = np.zeros((n_bif, 2, 2))
    L = np.zeros((n_bif, 2, 2))
    bif_of_t = np.argsort(tob)
    for t_ in range(n_bif):
        pe = sched.t_pedge[t_]
        if pe < 0:
            continue
        if lm[edges[pe][0]] == bif_of_t[t_]:
            U[t_], L[t_] = S[pe][0:2, 2:4], S[pe][2:4, 0:2]
        else:
            U[t_], L[t_] = S[pe][2:4, 0:2], S[pe][0:2, 2:4]
    Dinv = np.zeros_like(D0); G = np.zeros_like(D0); H = np.zeros_like(D0)
    z = np.zeros((n_bif, 2))

    def levels(c):
        return range(sched.chunk_lptr[c], sched.chunk_lptr[c + 1])

    def up(c):
        for lv in reversed(levels(c)):
            for n in range(sched.lvl_ptr[lv], sched.lvl_ptr[lv + 1]):
                Dn = D0[n].copy()
                for k in range(sched.t_cptr[n], sched.t_cptr[n + 1]):
                    ch = sched.t_cidx[k]
                    Dn -= H[ch] @ U[ch]
                    rz[n] -= H[ch] @ rz[ch]
                Dinv[n] = np.linalg.inv(Dn)
                G[n] = Dinv[n] @ U[n]
                H[n] = L[n] @ Dinv[n]

    def down(c):
        for lv in levels(c):
            for n in range(sched.lvl_ptr[lv], sched.lvl_ptr[lv + 1]):
                z[n] = Dinv[n] @ rz[n]
                if sched.t_parent[n] >= 0:
                    z[n] -= G[n] @ z[sched.t_parent[n]]

    for c in range(sched.n_chunks - 1):
        up(c)
    up(sched.n_chunks - 1)
    down(sched.n_chunks - 1)
    for c in range(sched.n_chunks - 1):
        down(c)
    x = np.zeros_like(r)
    for e in range(E):
        u, v = edges[e]
        ze = np.zeros(4)
        if lm[u] >= 0:
            ze[0:2] = z[tob[lm[u]]]
        if lm[v] >= 0:
            ze[2:4] = z[tob[lm[v]]]
        x[gl[e]] = y0[e] - Y[e] @ ze
    for b in range(n_bif):
        if pd >= 1:
            x[nq + nm.bifurcation_values[b]] = z[tob[b], 0]
        x[loff + b] = z[tob[b], 1]
    return x


@pytest.mark.parametrize("fd,pd", [(1, 0), (2, 0), (2, 1), (3, 2), (4, 3)])
def test_condensation_is_an_exact_solve_on_trees(fd, pd):
    rng = np.random.default_rng(100 * fd + pd)
    for G, N, strategy in ((ng.make_tree(2, 1, 3), 4, None), (helpers.random_tree(40, 3), 2, "smallest_last"),
                           (helpers.random_tree(25, 7), 1, "largest_first"), (helpers.linear_graph(2, 3), 3, None)):
        nm = nxfx.NetworkMesh(G, N=N, color_strategy=strategy)
        net = rp.OracleNetworkHO(nm._node_pos, nm.graph_edges, nm.edge_colors, N, fd, pd)
        nc = N * nm.graph_edges.shape[0]
        R, f = rng.uniform(0.5, 2.0, nc), rng.normal(size=nc)
        A, b = net.assemble(net.eval_pbc(lambda x: x[1] + 0.5 * x[2]), R=R, f=f)
        cond = condense.build_condensation(nm, fd, pd)
        assert cond.n_max == max(t.n for t in cond.types) and cond.kl <= fd + pd + 2
        sched = build_tree_schedule(nm.graph_edges, nm.node_multiplier_index, nm.bifurcation_values.size,
                                    root_hint_nodes=nm._boundary_out_nodes, chunk_nodes=8)
        assert sched.is_forest
        r = rng.normal(size=net.n_dofs)
        x = emulate(nm, cond, sched, R * net.cell_lengths(), r)
        x_ref = spla.spsolve(A.tocsc(), r)
        assert helpers.rel_l2(x, x_ref) < 1e-9, (fd, pd, N, helpers.rel_l2(x, x_ref))
        packed = cond.packed()
        assert packed["k_ptr"][-1] == packed["k_row"].size and packed["loc_ptr"][-1] == packed["loc_kind"].size
