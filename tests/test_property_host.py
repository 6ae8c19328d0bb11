"""Property tests (hypothesis) of the host logic on random directed graphs -- trees, forests and
graphs with cycles, arbitrary edge directions: the vectorised NetworkMesh analysis equals the
literal restatement of mesh.py, the dof layout equals the oracle's, schedules are valid, the
single-tree partition sums to the global system."""

import networkx as nx
import numpy as np
import scipy.sparse as sp
from hypothesis import given, settings
from hypothesis import strategies as st

import networks_fenicsx_b200 as nxfx
from networks_fenicsx_b200 import network_generation as ng
from networks_fenicsx_b200.distributed import partition_tree
from networks_fenicsx_b200.mesh import _greedy_edge_coloring_arrays
from networks_fenicsx_b200.schedule import build_tree_schedule
from oracle import reference_port as rp
from tests.test_host_logic import check_schedule


@st.composite
def random_graphs(draw, cycles=True):
    n = draw(st.integers(3, 24))
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    G = nx.DiGraph()
    for i in range(n):
        G.add_node(i, pos=rng.normal(size=3))
    und = set()
    for k in range(1, n):  # random recursive tree, random orientation; node 0 and 1 always joined
        p = 0 if k == 1 else int(rng.integers(0, k))
        und.add((p, k))
        G.add_edge(*((p, k) if rng.random() < 0.6 else (k, p)))
    if cycles:
        for _ in range(draw(st.integers(0, 4))):
            a, b = (int(v) for v in rng.integers(0, n, 2))
            if a != b and (min(a, b), max(a, b)) not in und:
                und.add((min(a, b), max(a, b)))
                G.add_edge(a, b)
    return G


@settings(max_examples=40, deadline=None)
@given(G=random_graphs(), N=st.integers(1, 4), colored=st.booleans())
def test_mesh_analysis_matches_literal(G, N, colored):
    strategy = "largest_first" if colored else None
    nm = nxfx.NetworkMesh(G, N=N, color_strategy=strategy)
    col = rp.color_graph_literal(G, strategy)
    info = rp.analyse_graph_literal(G, col)
    np.testing.assert_array_equal(nm.bifurcation_values, info.bifurcation_values)
    np.testing.assert_array_equal(nm.boundary_values, info.boundary_values)
    for i in range(len(info.bifurcation_values)):
        np.testing.assert_array_equal(nm.in_edges(i), info.in_color[info.in_offsets[i]:info.in_offsets[i + 1]])
        np.testing.assert_array_equal(nm.out_edges(i), info.out_color[info.out_offsets[i]:info.out_offsets[i + 1]])
    nodes, cells, markers, _ = rp.mesh_arrays_literal(G, N, col)
    np.testing.assert_array_equal(nm._cells(), cells)
    np.testing.assert_array_equal(nm.subdomains.values, markers)
    net = rp.OracleNetwork(*rp.graph_to_arrays(G, col), N)
    np.testing.assert_array_equal(nm.edge_slot * (N + 1), net.fb)
    # the pattern size formula of SURVEY Appendix B holds for any graph
    A, _ = net.assemble(net.eval_pbc(lambda x: x[0]))
    assert A.nnz == net.expected_nnz()
    s = build_tree_schedule(nm.graph_edges, nm.node_multiplier_index, nm.bifurcation_values.size,
                            root_hint_nodes=nm._boundary_out_nodes, chunk_nodes=5)
    if nm.bifurcation_values.size:
        check_schedule(nm, s, 5)


@settings(max_examples=25, deadline=None)
@given(G=random_graphs(cycles=False), N=st.integers(1, 3), world=st.integers(2, 4))
def test_tree_partition_sums_to_global(G, N, world):
    AG = ng.ArrayGraph(np.asarray([G.nodes[i]["pos"] for i in G.nodes()]), np.asarray(list(G.edges()), dtype=np.int64))
    colors = _greedy_edge_coloring_arrays(AG.number_of_nodes(), AG.edges)
    gnet = rp.OracleNetwork(AG.pos, AG.edges, colors, N)
    try:
        parts = [partition_tree(AG, world, r, chunk_nodes=3) for r in range(world)]
    except ValueError:
        return  # network too small for that many parts
    A, b = gnet.assemble(gnet.eval_pbc(lambda x: x[2]))
    S = sp.csr_matrix(A.shape)
    bsum = np.zeros_like(b)
    for part in parts:
        lc = _greedy_edge_coloring_arrays(part.graph.number_of_nodes(), part.graph.edges)
        lnet = rp.OracleNetwork(part.graph.pos, part.graph.edges, lc, N, degree=part.node_degree)
        Al, bl = lnet.assemble(lnet.eval_pbc(lambda x: x[2]))
        idx = np.empty(lnet.n_dofs, dtype=np.int64)
        for k, ge in enumerate(part.global_edges):
            idx[lnet.fb[k]:lnet.fb[k] + N + 1] = gnet.fb[ge] + np.arange(N + 1)
            idx[lnet.pb[k]:lnet.pb[k] + N] = gnet.pb[ge] + np.arange(N)
        idx[lnet.loff:] = gnet.loff + part.global_bif
        P = sp.csr_matrix((np.ones(idx.size), (np.arange(idx.size), idx)), shape=(idx.size, gnet.n_dofs))
        S = S + P.T @ Al @ P
        bsum += P.T @ bl
    assert abs(S - A).max() < 1e-13
    np.testing.assert_allclose(bsum, b, atol=1e-13)
    assert sum(p.global_edges.size for p in parts) == AG.number_of_edges()
