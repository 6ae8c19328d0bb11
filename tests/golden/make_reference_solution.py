"""Dump what the UNMODIFIED reference (DOLFINx + PETSc + MUMPS) assembles and solves, keyed canonically,
so that ``tests/test_reference_solution.py`` can pin the oracle -- and through it the CUDA path -- against
numbers this repository did not produce.

Run it where the reference runs, e.g. in its CI image (``.github/workflows/test_package.yml``):

    docker run --rm -v $PWD:/work -w /work ghcr.io/fenics/dolfinx/dolfinx:stable bash -c \
        "python3 -m pip install /path/to/networks_fenicsx && python3 tests/golden/make_reference_solution.py"

and commit the resulting ``tests/golden/reference_solution.npz``.  It CANNOT run in the build container of
this repository (no DOLFINx / PETSc / MPI, no network): until somebody produces the file, the assembled
values and the solution stay "parity unpinned by the reference"  and the consuming
test skips.  This script has therefore not been executed here; it only uses documented DOLFINx 0.10 API
(``tabulate_dof_coordinates``, ``Mat.getValuesCSR``, ``Vec.array``) and geometric matching, no numbering
internals.

Configs (BASELINE.json ``configs[0..3]``, the reference's own demos):
  y            demos/demo_Y_bifurcation.py          make_tree(2,1,3), N=4, uncoloured, p_bc = x[1]
  double_y     demos/demo_double_Y_bifurcation.py   make_tree(2,3.1,7.3), N=5, uncoloured, p_bc = x[0]
  tree         demos/demo_tree.py-style             make_tree(5,5,5), N=4, smallest_last, p_bc = x[1]
  arterial     demos/demo_arterial_tree.py          make_arterial_tree(5, [0.1,1,0]), N=40, largest_first

Canonical keying (the numbering the oracle and the CUDA path use): every DOLFINx dof is
located geometrically -- flux dof -> (graph edge e, fine index a in 0..fd*N along u -> v), pressure dof ->
(cell e*N + j) for DG0, multiplier dof -> bifurcation node -- and mapped to its canonical global index;
the matrix is permuted accordingly and stored as sorted CSR with its explicit zeros.
"""

import pathlib

import networkx as nx
import numpy as np

HERE = pathlib.Path(__file__).parent


def canonical_tables(pos, edges, colors, N, fd=1):
    """Canonical dof offsets: flux slots per colour block (ascending edge id inside a colour),
    pressure = cell id, multipliers = ascending bifurcation node id."""
    E = edges.shape[0]
    C = int(colors.max()) + 1
    per_edge = fd * N + 1
    count = np.bincount(colors, minlength=C)
    qoff = np.concatenate([[0], np.cumsum(count * per_edge)])
    order = np.argsort(colors, kind="stable")
    rank = np.empty(E, dtype=np.int64)
    rank[order] = np.arange(E) - np.concatenate([[0], np.cumsum(count)])[colors[order]]
    fb = qoff[colors] + rank * per_edge
    poff = int(qoff[-1])
    deg = np.bincount(edges.ravel(), minlength=pos.shape[0])
    bif = np.flatnonzero(deg > 1)
    return fb, poff, poff + E * N, bif, qoff


def locate_on_edges(X, pos3, edges, candidates, tol=1e-9):
    """For every point the (edge, parameter t in [0,1]) of the candidate edge it lies on."""
    out_e = np.full(X.shape[0], -1, dtype=np.int64)
    out_t = np.zeros(X.shape[0])
    for e in candidates:
        a, b = pos3[edges[e, 0]], pos3[edges[e, 1]]
        d = b - a
        L2 = float(d @ d)
        t = (X - a) @ d / L2
        dist = np.linalg.norm(X - (a + np.outer(t, d)), axis=1)
        hit = (dist < tol * max(1.0, np.sqrt(L2))) & (t > -1e-9) & (t < 1 + 1e-9)
        assert not np.any(hit & (out_e >= 0)), "a dof lies on two edges of one colour"
        out_e[hit] = e
        out_t[hit] = t[hit]
    assert np.all(out_e >= 0), "a dof could not be located on any candidate edge"
    return out_e, out_t


def dump_case(name, G, N, strategy, p_bc, out):
    from networks_fenicsx import HydraulicNetworkAssembler, NetworkMesh, Solver

    nm = NetworkMesh(G, N=N, color_strategy=strategy)
    assembler = HydraulicNetworkAssembler(nm, flux_degree=1, pressure_degree=0)
    assembler.compute_forms(p_bc_ex=p_bc)
    solver = Solver(assembler)  # kind=None: monolithic AIJ, block order [flux colours, pressure, lm]
    solver.assemble()
    sol = solver.solve()
    nodes = list(G.nodes())
    assert nodes == list(range(len(nodes)))
    pos = np.asarray([G.nodes[v]["pos"] for v in nodes], dtype=np.float64)
    pos3 = np.zeros((pos.shape[0], 3))
    pos3[:, : pos.shape[1]] = pos
    edges = np.asarray(list(G.edges()), dtype=np.int64)
    E = edges.shape[0]
    if strategy is None:
        colors = np.arange(E, dtype=np.int64)
    else:
        col = nx.coloring.greedy_color(nx.line_graph(G.to_undirected()), strategy=strategy)
        colors = np.asarray([col[(u, v)] if (u, v) in col else col[(v, u)] for u, v in edges.tolist()], dtype=np.int64)
    fb, poff, loff, bif, qoff = canonical_tables(pos, edges, colors, N)
    spaces = assembler.function_spaces
    sizes = [V.dofmap.index_map.size_local * V.dofmap.index_map_bs for V in spaces]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    n = int(offs[-1])
    perm = np.full(n, -1, dtype=np.int64)  # reference global index -> canonical global index
    C = len(spaces) - 2
    for c in range(C):
        X = spaces[c].tabulate_dof_coordinates()[: sizes[c]]
        e, t = locate_on_edges(X, pos3, edges, np.flatnonzero(colors == c))
        a = np.rint(t * N).astype(np.int64)
        assert np.allclose(t * N, a, atol=1e-7)
        perm[offs[c] + np.arange(sizes[c])] = fb[e] + a
    Xp = spaces[C].tabulate_dof_coordinates()[: sizes[C]]
    e, t = locate_on_edges(Xp, pos3, edges, np.arange(E))
    perm[offs[C] + np.arange(sizes[C])] = poff + e * N + np.minimum(np.floor(t * N).astype(np.int64), N - 1)
    Xl = spaces[C + 1].tabulate_dof_coordinates()[: sizes[C + 1]]
    for i, x in enumerate(Xl):
        node = int(np.argmin(np.linalg.norm(pos3 - x, axis=1)))
        assert np.linalg.norm(pos3[node] - x) < 1e-9
        perm[offs[C + 1] + i] = loff + int(np.searchsorted(bif, node))
    assert sorted(perm.tolist()) == list(range(n)), "the geometric matching is not a permutation"
    indptr, indices, data = solver.A.getValuesCSR()
    import scipy.sparse as sp

    A = sp.csr_matrix((data, indices, indptr), shape=(n, n)).tocoo()
    Acan = sp.coo_matrix((np.ones_like(A.data), (perm[A.row], perm[A.col])), shape=(n, n)).tocsr()  # pattern incl. explicit zeros
    Aval = sp.coo_matrix((A.data, (perm[A.row], perm[A.col])), shape=(n, n)).tocsr()
    Acan.sort_indices()
    vals = np.asarray(Aval[Acan.nonzero()]).ravel()  # values on the stored pattern (explicit zeros stay zeros)
    b = np.empty(n)
    b[perm] = solver.b.array[:n]
    x = np.empty(n)
    x[perm] = np.concatenate([f.x.array[: sizes[i]] for i, f in enumerate(sol)])
    for key, val in dict(pos=pos, edges=edges, colors=colors, N=np.int64(N), indptr=Acan.indptr, indices=Acan.indices,
                         values=vals, b=b, x=x).items():
        out[f"{name}/{key}"] = val
    print(f"{name}: {n} dofs, nnz {Acan.nnz}")


def main():
    import ufl  # noqa: F401  (the reference's demos pass UFL expressions; callables are used here)
    from networks_fenicsx import network_generation as ng

    out = {}
    dump_case("y", ng.make_tree(2, 1, 3), 4, None, lambda x: x[1], out)
    dump_case("double_y", ng.make_tree(2, 3.1, 7.3), 5, None, lambda x: x[0], out)
    dump_case("tree", ng.make_tree(5, 5, 5), 4, "smallest_last", lambda x: x[1], out)
    dump_case("arterial", ng.make_arterial_tree(N=5, direction=np.array([0.1, 1, 0])), 40, "largest_first", lambda x: x[1], out)
    np.savez_compressed(HERE / "reference_solution.npz", **out)
    print("wrote", HERE / "reference_solution.npz")


if __name__ == "__main__":
    main()
