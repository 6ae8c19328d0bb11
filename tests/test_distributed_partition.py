"""Single-tree edge partition (distributed.py): the partitioned assembly must sum to the global
system, the cut multipliers are replicated consistently, and the local elimination schedules share
an identical top chunk.  CPU only (the per-rank systems are assembled with the oracle)."""

import numpy as np
import pytest
import scipy.sparse as sp

from networks_fenicsx_b200 import network_generation as ng
from networks_fenicsx_b200.distributed import partition_tree
from networks_fenicsx_b200.mesh import _greedy_edge_coloring_arrays
from oracle import reference_port as rp
from tests import helpers


def global_oracle(G, N):
    colors = _greedy_edge_coloring_arrays(G.number_of_nodes(), G.edges)
    return rp.OracleNetwork(G.pos, G.edges, colors, N)


def local_to_global_dofs(part, lnet, gnet, N):
    """Global dof of every local dof (flux dofs follow their edge, multipliers their node)."""
    idx = np.empty(lnet.n_dofs, dtype=np.int64)
    for k, ge in enumerate(part.global_edges):
        idx[lnet.fb[k]:lnet.fb[k] + N + 1] = gnet.fb[ge] + np.arange(N + 1)
        idx[lnet.pb[k]:lnet.pb[k] + N] = gnet.pb[ge] + np.arange(N)
    idx[lnet.loff:] = gnet.loff + part.global_bif
    return idx


@pytest.mark.parametrize("case,world,N,chunk", [
    ("tree8", 2, 1, 16), ("tree8", 4, 3, 16), ("tree9", 8, 1, 8), ("random", 3, 2, 12), ("arterial", 2, 2, 4),
])
def test_partitioned_assembly_sums_to_global(case, world, N, chunk):
    if case.startswith("tree"):
        G = ng.make_tree(int(case[4:]), 3, 5, as_arrays=True)
    elif case == "arterial":
        G = ng.make_arterial_tree(6, as_arrays=True)
    else:
        H = helpers.random_tree(300, 7)
        G = ng.ArrayGraph(np.asarray([H.nodes[i]["pos"] for i in H.nodes()]), np.asarray(list(H.edges()), dtype=np.int64))
    p_bc = lambda x: x[1] - 0.5 * x[0]  # noqa: E731
    gnet = global_oracle(G, N)
    A, b = gnet.assemble(gnet.eval_pbc(p_bc), R=1.3, f=0.2)
    S = sp.csr_matrix(A.shape)
    bsum = np.zeros_like(b)
    edge_seen = np.zeros(G.number_of_edges(), dtype=int)
    top_orders, private = [], []
    for rank in range(world):
        part = partition_tree(G, world, rank, chunk_nodes=chunk)
        edge_seen[part.global_edges] += 1
        colors = _greedy_edge_coloring_arrays(part.graph.number_of_nodes(), part.graph.edges)
        lnet = rp.OracleNetwork(part.graph.pos, part.graph.edges, colors, N, degree=part.node_degree)
        np.testing.assert_array_equal(part.global_nodes[lnet.bifurcation_values],
                                      gnet.bifurcation_values[part.global_bif])
        Al, bl = lnet.assemble(lnet.eval_pbc(p_bc), R=1.3, f=0.2)
        idx = local_to_global_dofs(part, lnet, gnet, N)
        P = sp.csr_matrix((np.ones(idx.size), (np.arange(idx.size), idx)), shape=(idx.size, gnet.n_dofs))
        S = S + P.T @ Al @ P
        bsum += P.T @ bl
        # schedule sanity: top chunk last, identical (in global ids) on every rank
        s = part.schedule
        t0 = s.lvl_ptr[s.chunk_lptr[-2]]
        bif_of_t = np.argsort(s.t_of_bif)
        top_nodes = part.global_bif[bif_of_t[t0:]].tolist()
        shared_global = part.global_bif[part.shared_lm].tolist()
        top_orders.append(shared_global)  # exchanged in THIS order: must be the same on every rank
        private.append(sorted(set(top_nodes) - set(shared_global)))
        assert part.n_top == part.shared_lm.size and set(shared_global) <= set(top_nodes)
        assert np.all(part.lam_weight[part.shared_lm] == (1.0 if rank == 0 else 0.0))
        assert np.all(np.delete(part.lam_weight, part.shared_lm) == 1.0)
        for t in range(s.t_parent.size):  # tree links stay inside the part
            if s.t_parent[t] >= 0 and s.t_pedge[t] >= 0:
                u, v = part.graph.edges[s.t_pedge[t]]
                gu, gv = part.global_nodes[u], part.global_nodes[v]
                assert {gnet.lm_index[gu], gnet.lm_index[gv]} == {part.global_bif[bif_of_t[t]], part.global_bif[bif_of_t[s.t_parent[t]]]}
    assert np.all(edge_seen == 1), "every graph edge belongs to exactly one rank"
    assert all(o == top_orders[0] for o in top_orders), "the shared multipliers (or their order) differ between ranks"
    # heavy nodes whose subtree stays on one rank are private to it: nobody else holds them
    allp = [n for p in private for n in p]
    assert len(allp) == len(set(allp)) and not set(allp) & set(top_orders[0])
    if case.startswith("tree"):  # balanced binary tree, chunks dealt in order: only the nodes above the rank subtrees are shared
        assert len(top_orders[0]) == world - 1
    assert abs(S - A).max() < 1e-14
    S.eliminate_zeros()
    np.testing.assert_allclose(bsum, b, rtol=0, atol=1e-14)
    # balance: bottom chunks are dealt evenly
    counts = np.bincount([r for r in range(world) for _ in partition_tree(G, world, r, chunk).global_edges], minlength=world)
    assert counts.max() <= 2.5 * max(counts.min(), 1)


def test_partition_rejects_cycles_and_tiny_graphs():
    H = helpers.edge_info_graph()
    G = ng.ArrayGraph(np.asarray([H.nodes[i]["pos"] for i in H.nodes()]), np.asarray(list(H.edges()), dtype=np.int64))
    with pytest.raises(NotImplementedError):
        partition_tree(G, 2, 0)
    with pytest.raises(ValueError):
        partition_tree(ng.make_tree(2, 1, 1, as_arrays=True), 4, 0)


def test_network_mesh_partitions_transparently_under_a_multi_rank_comm():
    """NetworkMesh(G, N, comm=<size 2>) keeps this rank's part (reference: mesh.py:84-96 under mpiexec);
    too-small / cyclic networks are replicated with a warning."""
    import types

    import networkx as nx
    import pytest

    import networks_fenicsx_b200 as nxfx
    from networks_fenicsx_b200.distributed import partition_tree

    G = ng.make_tree(9, 9, 9, as_arrays=True)
    sizes = []
    for rank in range(2):
        comm = types.SimpleNamespace(size=2, rank=rank)
        nm = nxfx.NetworkMesh(G, N=2, color_strategy="smallest_last", comm=comm, device=0)
        part = partition_tree(G, 2, rank)
        assert nm._partition is not None and np.array_equal(nm._partition.global_edges, part.global_edges)
        assert np.array_equal(nm.graph_edges, part.graph.edges)
        assert np.array_equal(nm.bifurcation_values, np.flatnonzero(part.node_degree > 1))
        sizes.append(nm.graph_edges.shape[0])
    assert sum(sizes) == G.number_of_edges()
    # a networkx graph with edge attributes is cut the same way
    Gx = ng.make_arterial_tree(N=8, direction=np.array([0.1, 1.0, 0.0]))
    nm = nxfx.NetworkMesh(Gx, N=1, comm=types.SimpleNamespace(size=2, rank=1), device=0)
    assert nm._partition is not None and "radius" in nm._global_graph.edge_attrs
    # too small to cut -> replicated, with a warning on rank 0
    with pytest.warns(UserWarning, match="every rank solves all of it"):
        nm = nxfx.NetworkMesh(ng.make_tree(2, 1, 3), N=4, comm=types.SimpleNamespace(size=2, rank=0), device=0)
    assert nm._partition is None and nm.graph_edges.shape[0] == 3
    Gc = nx.DiGraph()
    for i, p in enumerate([(0, 0), (0, 1), (1, 2), (-1, 2), (0, 3), (0, 4)]):
        Gc.add_node(i, pos=np.array(p, dtype=float))
    for e in [(0, 1), (1, 2), (1, 3), (2, 4), (3, 4), (4, 5)]:
        Gc.add_edge(*e)
    with pytest.warns(UserWarning):
        nm = nxfx.NetworkMesh(Gc, N=2, comm=types.SimpleNamespace(size=2, rank=0), device=0)
    assert nm._partition is None
