"""Host-side logic of the product (no GPU): generators against the reference fixtures, graph
analysis / dof layout against the oracle's literal restatement, the reference's own structural
tests on the NetworkMesh duck types, the elimination schedule."""

import pathlib

import networkx as nx
import numpy as np
import pytest

import networks_fenicsx_b200 as nxfx
from networks_fenicsx_b200 import network_generation as ng
from networks_fenicsx_b200.mesh import _greedy_edge_coloring_arrays, color_graph
from networks_fenicsx_b200.schedule import build_tree_schedule
from oracle import reference_port as rp
from tests import helpers

GOLDEN = np.load(pathlib.Path(__file__).parent / "golden" / "reference_graphs.npz")


def graph_arrays(G):
    pos = np.asarray([G.nodes[v]["pos"] for v in G.nodes()], dtype=float)
    return pos, np.asarray(list(G.edges()), dtype=np.int64)


# ---- generators: bit-for-bit against the reference's make_tree / make_arterial_tree --------------
def test_native_greedy_colouring_on_long_chains():
    """Path-like graphs colour an edge or two per round: after 64 rounds the native greedy colouring finishes
    sequentially (round-1 advice) -- same visiting order, same colours as the plain edge-by-edge loop."""
    from networks_fenicsx_b200 import mesh as nxmesh

    def plain(n_nodes, edges):
        used = [set() for _ in range(n_nodes)]
        out = []
        for u, v in edges.tolist():
            c = 0
            while c in used[u] or c in used[v]:
                c += 1
            out.append(c)
            used[u].add(c)
            used[v].add(c)
        return np.asarray(out)

    E = 5000
    path = np.stack([np.arange(E), np.arange(1, E + 1)], axis=1)
    rng = np.random.default_rng(1)
    chains = np.stack([np.maximum(np.arange(1, E + 1) - rng.integers(1, 3, E), 0), np.arange(1, E + 1)], axis=1)
    for edges in (path, chains):
        col = nxmesh._greedy_edge_coloring_arrays(E + 1, edges)
        assert np.array_equal(col, plain(E + 1, edges))


def test_boundary_vertex_runs_for_the_pbc_upload():
    """compute_forms uploads only the contiguous runs of boundary vertex ids (the forms read p_bc nowhere else,
    assembly.py:258-260); scattered boundary ids fall back to the full array."""
    from networks_fenicsx_b200.assembly import HydraulicNetworkAssembler

    def runs_of(nm):
        asm = HydraulicNetworkAssembler.__new__(HydraulicNetworkAssembler)
        asm._network_mesh = nm
        return asm._boundary_runs()

    nm = nxfx.NetworkMesh(ng.make_tree(8, 8, 8, as_arrays=True), N=1, color_strategy="smallest_last")
    runs = runs_of(nm)
    assert runs == [(0, 1), (128, 256)]
    covered = np.concatenate([np.arange(a, b) for a, b in runs])
    assert np.array_equal(covered, np.sort(nm.boundary_values))
    assert runs_of(nxfx.NetworkMesh(helpers.random_tree(80, 5), N=2, color_strategy="smallest_last")) is None


def test_bench_colouring_note_follows_the_mesh_thresholds():
    import bench
    from networks_fenicsx_b200 import mesh as nxmesh

    assert "reference's own call" in bench.colouring_note(nxmesh.NETWORKX_COLORING_MAX_EDGES, 1)
    assert "networkx-identical" in bench.colouring_note(nxmesh.NETWORKX_COLORING_MAX_EDGES + 1, 1)
    assert "sub-network" in bench.colouring_note(1_048_575, 8)
    assert "native greedy" in bench.colouring_note(nxmesh.NETWORKX_IDENTICAL_MAX_EDGES + 1, 1)


def test_tree_edges_matches_reference_generator():
    """network_generation.py:18-38 restated literally (parent stack) against the closed form."""
    def literal(n, r):
        if n == 0:
            return
        yield 0, 1
        nodes = iter(range(1, n))
        parents = [next(nodes)]
        while parents:
            source = parents.pop(0)
            for _ in range(r):
                try:
                    target = next(nodes)
                    parents.append(target)
                    yield source, target
                except StopIteration:
                    break

    for n in (0, 2, 3, 7, 8, 30, 100):  # (n = 1 raises inside the reference generator)
        for r in (1, 2, 3, 5):
            assert list(ng.tree_edges(n, r)) == list(literal(n, r)), (n, r)


def test_make_tree_matches_reference_fixtures():
    keys = sorted({k.split("/")[0] for k in GOLDEN.files if k.startswith("tree_")})
    assert len(keys) >= 30
    for key in keys:
        _, n, H, W, d = key.split("_")
        n, H, W, d = int(n[1:]), float(H[1:]), float(W[1:]), int(d[1:])
        H, W = (int(H) if H == int(H) else H), (int(W) if W == int(W) else W)
        G = ng.make_tree(n, H, W, d)
        pos, edges = graph_arrays(G)
        assert np.array_equal(pos, GOLDEN[key + "/pos"]) and np.array_equal(edges, GOLDEN[key + "/edges"]), key
        assert list(G.nodes()) == GOLDEN[key + "/nodes"].tolist()
        A = ng.make_tree(n, H, W, d, as_arrays=True)
        assert np.array_equal(A.pos, pos) and np.array_equal(A.edges, edges)


def test_make_arterial_tree_matches_reference_fixtures():
    keys = sorted({k.split("/")[0] for k in GOLDEN.files if k.startswith("arterial_")})
    assert len(keys) == 4
    for key in keys:
        parts = key.split("_")
        N, gam = int(parts[1][1:]), float(parts[2][1:])
        direction = np.array([float(parts[3][1:])] + [float(x) for x in parts[4:]])
        G = ng.make_arterial_tree(N=N, direction=direction, gamma=gam)
        pos, edges = graph_arrays(G)
        radius = np.asarray([G.edges[e]["radius"] for e in G.edges()])
        assert np.array_equal(pos, GOLDEN[key + "/pos"]), key
        assert np.array_equal(edges, GOLDEN[key + "/edges"])
        assert np.array_equal(radius, GOLDEN[key + "/radius"])
        A = ng.make_arterial_tree(N=N, direction=direction, gamma=gam, as_arrays=True)
        assert np.array_equal(A.pos, pos) and np.array_equal(A.edge_attrs["radius"], radius)
    with pytest.raises(ValueError):
        ng.make_arterial_tree(3, gamma=1.5)


def test_vectorised_arterial_tree_equals_the_per_vessel_loop():
    """Generation-wise construction vs the reference-shaped per-vessel loop at a size the fixtures do not
    reach (4095 vessels; the vectorised form has to use the scalar power and the per-vessel BLAS product to
    stay bit-identical)."""
    for N, gamma, direction in ((12, 0.8, [0.1, 1.0, 0.0]), (11, 0.5, [1.0, 1.0, 0.0])):
        A = ng.make_arterial_tree(N=N, direction=np.array(direction), gamma=gamma, as_arrays=True)
        pos = np.empty((2**N, 3))
        rad = np.empty(2**N - 1)
        ed = np.empty((2**N - 1, 2), dtype=np.int64)
        pos[:2], ed[0], rad[0] = A.pos[:2], (0, 1), A.edge_attrs["radius"][0]
        ng._arterial_generations_loop(N, pos, rad, ed, 8.0, gamma, ng._default_normal, False)
        assert np.array_equal(pos, A.pos) and np.array_equal(ed, A.edges) and np.array_equal(rad, A.edge_attrs["radius"])
    # a custom surface normal takes the loop form
    B = ng.make_arterial_tree(N=4, normal=lambda x: np.array([0.0, 0.2, 1.0]), as_arrays=True)
    assert B.pos.shape == (16, 3) and np.abs(B.pos[:, 2]).max() > 0


def test_large_tree_generation_is_fast():
    A = ng.make_tree(18, 18, 18, as_arrays=True)
    assert A.number_of_edges() == 2**18 - 1 and A.pos.shape == (2**18, 3)


# ---- colouring ---------------------------------------------------------------------------------
def test_color_graph_matches_reference_call_sequence():
    for n in (3, 5, 7):
        G = ng.make_tree(n, 1, 1)
        for strat in ("smallest_last", "largest_first"):
            nm = nxfx.NetworkMesh(G, N=1, color_strategy=strat)
            assert nm.edge_colors.tolist() == GOLDEN[f"color_tree_n{n}_{strat}"].tolist()
    G = ng.make_tree(3, 1, 1)
    assert color_graph(G, None) == {e: i for i, e in enumerate(G.edges)}
    nm = nxfx.NetworkMesh(G, N=2, color_strategy=nx.coloring.strategy_largest_first)
    assert nm.num_edge_colors == 3


def test_reversed_edge_keys_do_not_raise():
    """the reference raises KeyError for edges keyed (v,u) by the line graph."""
    G = nx.DiGraph()
    for i in range(4):
        G.add_node(i, pos=np.array([float(i), 0.0]))
    for e in [(1, 0), (2, 1), (3, 2)]:
        G.add_edge(*e)
    nm = nxfx.NetworkMesh(G, N=2, color_strategy="largest_first")
    assert nm.num_edge_colors == 2


def _networkx_colors(G, strategy):
    """mesh.py:38-39 verbatim, per edge in graph.edges() order."""
    col = nx.coloring.greedy_color(nx.line_graph(G.to_undirected()), strategy=strategy)
    return np.asarray([col[(u, v)] if (u, v) in col else col[(v, u)] for u, v in G.edges()], dtype=np.int32)


@pytest.mark.parametrize("strategy", ["smallest_last", "largest_first"])
def test_networkx_identical_coloring_equals_networkx(strategy):
    """The set-order emulation reproduces networkx edge by edge: binary trees, arterial trees, random
    trees with random orientation, graphs with cycles, a star and a chain."""
    from networks_fenicsx_b200.mesh import _networkx_identical_edge_coloring

    graphs = [ng.make_tree(n, 1, 1) for n in (2, 3, 6, 9, 12)]
    graphs += [ng.make_arterial_tree(N=N, direction=np.array([0.1, 1.0, 0.0])) for N in (3, 7, 10)]
    graphs += [helpers.random_tree(n, seed) for n, seed in ((40, 0), (300, 1), (2000, 2))]
    graphs += [helpers.edge_info_graph(), helpers.double_junction_graph(), helpers.linear_graph(50)]
    rng = np.random.default_rng(7)
    for _ in range(4):  # random connected graphs with chords and mixed orientation
        n = int(rng.integers(10, 120))
        G = nx.DiGraph()
        G.add_nodes_from(range(n))
        und = set()
        for k in range(1, n):
            p = int(rng.integers(0, k))
            und.add((p, k))
            G.add_edge(*((p, k) if rng.random() < 0.5 else (k, p)))
        for _ in range(n // 3):
            a, b = (int(v) for v in rng.integers(0, n, 2))
            if a != b and (min(a, b), max(a, b)) not in und:
                und.add((min(a, b), max(a, b)))
                G.add_edge(a, b)
        graphs.append(G)
    for G in graphs:
        edges = np.asarray(list(G.edges()), dtype=np.int64)
        mine = _networkx_identical_edge_coloring(G.number_of_nodes(), edges, strategy)
        assert np.array_equal(mine, _networkx_colors(G, strategy)), (G.number_of_nodes(), strategy)


def test_large_tree_coloring_is_the_reference_coloring():
    """Above NETWORKX_COLORING_MAX_EDGES (every headline run) NetworkMesh still gives networkx's
    ``smallest_last`` colouring: compared with fixtures that networkx itself produced for the 16- and the
    20-generation tree (tests/golden/make_golden.py --large; 84 s / 3 GB for n = 20)."""
    large = np.load(pathlib.Path(__file__).parent / "golden" / "reference_colors_large.npz")
    A = ng.make_tree(16, 16, 16, as_arrays=True)
    assert A.number_of_edges() > 60000
    import networks_fenicsx_b200.mesh as mesh_mod

    old = mesh_mod.NETWORKX_COLORING_MAX_EDGES
    mesh_mod.NETWORKX_COLORING_MAX_EDGES = 1000  # force the large-graph path at a size the suite affords
    try:
        nm = nxfx.NetworkMesh(A, N=1, color_strategy="smallest_last")
    finally:
        mesh_mod.NETWORKX_COLORING_MAX_EDGES = old
    assert np.array_equal(nm.edge_colors, large["color_tree_n16_smallest_last"])


@pytest.mark.slow
def test_headline_tree_coloring_is_the_reference_coloring():
    """The 20-generation headline graph (1,048,575 edges): 26 s."""
    large = np.load(pathlib.Path(__file__).parent / "golden" / "reference_colors_large.npz")
    A = ng.make_tree(20, 20, 20, as_arrays=True)
    nm = nxfx.NetworkMesh(A, N=1, color_strategy="smallest_last")
    assert np.array_equal(nm.edge_colors, large["color_tree_n20_smallest_last"])


def test_unreproducible_strategy_at_scale_warns():
    from networks_fenicsx_b200.mesh import _large_graph_colors

    A = ng.make_tree(9, 1, 1, as_arrays=True)
    with pytest.warns(UserWarning, match="native greedy"):
        col = _large_graph_colors(A.number_of_nodes(), A.edges, "saturation_largest_first")
    assert col.max() == 2
    with pytest.warns(UserWarning, match="native greedy"):
        _large_graph_colors(A.number_of_nodes(), A.edges, lambda G, colors: list(G))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_native_greedy_coloring_is_proper(seed):
    G = helpers.random_tree(300, seed)
    edges = np.asarray(list(G.edges()), dtype=np.int64)
    col = _greedy_edge_coloring_arrays(300, edges)
    deg = np.bincount(edges.ravel())
    assert col.min() == 0 and col.max() + 1 <= 2 * deg.max() - 1
    for node in range(300):
        inc = col[(edges[:, 0] == node) | (edges[:, 1] == node)]
        assert len(inc) == len(set(inc.tolist()))
    A = ng.make_tree(12, 1, 1, as_arrays=True)
    assert _greedy_edge_coloring_arrays(A.number_of_nodes(), A.edges).max() == 2  # 3 colours on a binary tree


# ---- NetworkMesh host analysis vs the oracle's literal restatement of mesh.py ----------------------
CASES = [
    ("edge_info", helpers.edge_info_graph, None, 10),
    ("edge_info_col", helpers.edge_info_graph, "largest_first", 3),
    ("tree5", lambda: ng.make_tree(5, 1, 1), "smallest_last", 4),
    ("tree3_uncol", lambda: ng.make_tree(3, 2, 1, 2), None, 1),
    ("line_alt", lambda: helpers.linear_graph(12, ordered=lambda k: k % 2), None, 2),
    ("random", lambda: helpers.random_tree(80, 3), "smallest_last", 2),
    ("two_junctions", helpers.double_junction_graph, "largest_first", 5),
]


@pytest.mark.parametrize("name,make,strategy,N", CASES)
def test_network_mesh_matches_literal_reference_logic(name, make, strategy, N):
    G = make()
    nm = nxfx.NetworkMesh(G, N=N, color_strategy=strategy)
    col = rp.color_graph_literal(G, strategy)
    info = rp.analyse_graph_literal(G, col)
    assert nm.num_edge_colors == info.num_edge_colors
    np.testing.assert_array_equal(nm.bifurcation_values, info.bifurcation_values)
    np.testing.assert_array_equal(nm.boundary_values, info.boundary_values)
    for i in range(len(info.bifurcation_values)):
        np.testing.assert_array_equal(nm.in_edges(i), info.in_color[info.in_offsets[i]:info.in_offsets[i + 1]])
        np.testing.assert_array_equal(nm.out_edges(i), info.out_color[info.out_offsets[i]:info.out_offsets[i + 1]])
    tags, im, om = rp.vertex_markers_literal(info)
    assert (nm.in_marker, nm.out_marker) == (im, om)
    np.testing.assert_array_equal(nm.boundaries.values, tags)
    nodes, cells, markers, orient = rp.mesh_arrays_literal(G, N, col)
    np.testing.assert_array_equal(nm._cells(), cells)
    np.testing.assert_array_equal(nm.subdomains.values, markers)
    np.testing.assert_array_equal(nm.orientation.x.array, rp.orientation_net_effect(cells))
    assert nm.mesh.topology.dim == 1 and nm.mesh.geometry.dim == nodes.shape[1]
    assert nm.mesh.topology.index_map(1).size_global == cells.shape[0]
    assert nm.mesh.topology.index_map(0).size_global == nodes.shape[0]
    # dof layout against the oracle
    net = rp.OracleNetwork(*rp.graph_to_arrays(G, col), N)
    np.testing.assert_array_equal(nm.edge_slot * (N + 1), net.fb)
    asm = nxfx.HydraulicNetworkAssembler(nm)
    assert asm.block_sizes == net.block_sizes and asm.num_dofs == net.n_dofs
    # integration entities (assembly.py:28-92)
    infl, outfl = rp.integration_entities_literal(G, N, col, info)
    data = dict(asm._integration_data)
    for c in range(nm.num_edge_colors):
        np.testing.assert_array_equal(data[asm._in_idx + c], infl[c])
        np.testing.assert_array_equal(data[asm._out_idx + c], outfl[c])
    # entity maps / submeshes
    for c in range(min(nm.num_edge_colors, 4)):
        cells_c = nm.entity_maps[c].sub_topology_to_topology(np.arange(nm.submeshes[c].topology.index_map(1).size_local))
        np.testing.assert_array_equal(cells_c, np.flatnonzero(markers == c))
        np.testing.assert_array_equal(nm.entity_maps[c].sub_topology_to_topology(cells_c, inverse=True), np.arange(cells_c.size))
    assert nm.lm_mesh.topology.index_map(0).size_global == len(info.bifurcation_values)
    # incidence table: sorted by slot within a bifurcation, signs = in/out
    for i, b in enumerate(nm.bifurcation_values):
        inc = nm._bif_inc[nm._bif_ptr[i]:nm._bif_ptr[i + 1]]
        e = inc >> 1
        assert np.all(np.diff(nm.edge_slot[e]) > 0)
        for code in inc:
            u, v = nm.graph_edges[code >> 1]
            assert (v == b) if (code & 1) else (u == b)


@pytest.mark.parametrize("N", [10, 50])
def test_edge_info_reference_test(N):
    """tests/test_edge_info.py of the reference, verbatim assertions."""
    network_mesh = nxfx.NetworkMesh(helpers.edge_info_graph(), N=N)
    assert len(network_mesh.bifurcation_values) == 6
    np.testing.assert_allclose([1, 2, 3, 4, 5, 7], network_mesh.bifurcation_values)
    expected = [(1, 1), (1, 1), (1, 1), (2, 1), (2, 1), (1, 3)]
    for i, (n_in, n_out) in enumerate(expected):
        assert len(network_mesh.in_edges(i)) == n_in
        assert len(network_mesh.out_edges(i)) == n_out


@pytest.mark.parametrize("gdim", [2, 3])
@pytest.mark.parametrize("N", [1, 4, 10])
@pytest.mark.parametrize("n", [2, 5, 7])
@pytest.mark.parametrize("H", [1, 2])
def test_make_tree_reference_test(n, H, gdim, N):
    """tests/test_make_tree.py of the reference, verbatim assertions."""
    G = ng.make_tree(n=n, H=H, W=1, dim=gdim)
    network_mesh = nxfx.NetworkMesh(G, N=N)
    domain = network_mesh.mesh
    tdim = domain.topology.dim
    assert tdim == 1
    assert domain.geometry.dim == gdim
    num_segments = sum(2**i for i in range(n))
    assert domain.topology.index_map(tdim).size_global == N * num_segments
    assert domain.topology.index_map(0).size_global == N + 1 + (num_segments - 1) * N


def test_invalid_inputs():
    G = nx.DiGraph()
    G.add_node(5, pos=np.zeros(2))
    G.add_node(0, pos=np.ones(2))
    G.add_edge(5, 0)
    with pytest.raises(ValueError):
        nxfx.NetworkMesh(G, N=1)
    with pytest.raises(ValueError):
        nxfx.NetworkMesh(ng.make_tree(2, 1, 1), N=0)
    nm = nxfx.NetworkMesh(ng.make_tree(2, 1, 1), N=1)
    with pytest.raises(ValueError):
        nxfx.HydraulicNetworkAssembler(nm, flux_degree=0, pressure_degree=0)
    assert nxfx.__version__ is not None  # tests/test_version.py


def test_api_surface_matches_reference():
    """__init__.py:19-25 exports and the public method names of the three classes."""
    for name in ["HydraulicNetworkAssembler", "NetworkMesh", "post_processing", "Solver", "network_generation"]:
        assert hasattr(nxfx, name)
    for name in ["lm_mesh", "lm_map", "comm", "submesh_facet_markers", "mesh", "subdomains", "boundaries", "submeshes",
                 "entity_maps", "orientation", "bifurcation_values", "boundary_values", "in_edges", "out_edges",
                 "num_edge_colors", "in_marker", "out_marker"]:
        assert hasattr(nxfx.NetworkMesh, name), name
    for name in ["compute_forms", "lm_space", "pressure_space", "flux_spaces", "function_spaces", "network", "assemble",
                 "bilinear_forms", "bilinear_form", "linear_forms", "linear_form"]:
        assert hasattr(nxfx.HydraulicNetworkAssembler, name), name
    for name in ["assembler", "A", "b", "assemble", "ksp", "solve"]:
        assert hasattr(nxfx.Solver, name), name
    for name in ["extract_global_flux", "export_functions", "export_submeshes"]:
        assert hasattr(nxfx.post_processing, name)
    from networks_fenicsx_b200.common import timed, timing

    @timed("nxfx:test")
    def f():
        return 1

    f()
    assert timing("nxfx:test")[0] == 1


# ---- elimination schedule -----------------------------------------------------------------------
def check_schedule(nm, s, chunk_nodes):
    n_bif = nm.bifurcation_values.size
    assert sorted(s.t_of_bif.tolist()) == list(range(n_bif))
    assert s.lvl_ptr[0] == 0 and s.lvl_ptr[-1] == n_bif and s.chunk_lptr[-1] == s.lvl_ptr.size - 1
    lm = nm.node_multiplier_index
    bif_of_t = np.argsort(s.t_of_bif)
    level_of = np.repeat(np.arange(s.lvl_ptr.size - 1), np.diff(s.lvl_ptr))
    chunk_of_level = np.repeat(np.arange(s.n_chunks), np.diff(s.chunk_lptr))
    n_tree_edges = 0
    for t in range(n_bif):
        p, e = s.t_parent[t], s.t_pedge[t]
        assert (p >= 0) == (e >= 0)
        if p >= 0:
            n_tree_edges += 1
            u, v = nm.graph_edges[e]
            assert {int(lm[u]), int(lm[v])} == {int(bif_of_t[t]), int(bif_of_t[p])}
            ct, cp = chunk_of_level[level_of[t]], chunk_of_level[level_of[p]]
            # the parent is in the same chunk at a shallower level, or in the top (last) chunk
            assert (ct == cp and level_of[p] < level_of[t]) or (cp == s.n_chunks - 1 and ct != cp)
            assert t in s.t_cidx[s.t_cptr[p]:s.t_cptr[p + 1]]
    assert s.t_cptr[-1] == n_tree_edges
    for c in range(s.n_chunks - 1):
        lo, hi = s.lvl_ptr[s.chunk_lptr[c]], s.lvl_ptr[s.chunk_lptr[c + 1]]
        assert 0 < hi - lo <= chunk_nodes
    links = [(u, v) for u, v in nm.graph_edges if lm[u] >= 0 and lm[v] >= 0]
    assert len(links) == n_tree_edges + s.chord_edge.size


@pytest.mark.parametrize("name,make,chunk", [
    ("tree9", lambda: ng.make_tree(9, 1, 1), 16), ("tree9_big", lambda: ng.make_tree(9, 1, 1), 2048),
    ("random", lambda: helpers.random_tree(400, 5), 32), ("cyclic", helpers.edge_info_graph, 4),
    ("line", lambda: helpers.linear_graph(30), 8), ("y", lambda: ng.make_tree(2, 1, 3), 2048),
])
def test_tree_schedule(name, make, chunk):
    nm = nxfx.NetworkMesh(make(), N=1)
    s = build_tree_schedule(nm.graph_edges, nm.node_multiplier_index, nm.bifurcation_values.size,
                            root_hint_nodes=nm._boundary_out_nodes, chunk_nodes=chunk)
    check_schedule(nm, s, chunk)
    assert s.is_forest == (name != "cyclic")
    if name == "cyclic":
        assert s.chord_edge.size == 2  # 7 links among 6 bifurcations, spanning tree has 5


def test_single_edge_graph_has_empty_schedule():
    G = nx.DiGraph()
    G.add_node(0, pos=np.zeros(3))
    G.add_node(1, pos=np.ones(3))
    G.add_edge(0, 1)
    nm = nxfx.NetworkMesh(G, N=3)
    assert nm.bifurcation_values.size == 0
    s = build_tree_schedule(nm.graph_edges, nm.node_multiplier_index, 0)
    assert s.n_chunks == 0


# ---- higher-order (table-driven) path: host symbolic vs the oracle ---------------------------------
@pytest.mark.parametrize("fd,pd", [(2, 1), (2, 0), (1, 1), (3, 2)])
@pytest.mark.parametrize("case", ["y", "random", "cyclic"])
def test_generic_symbolic_matches_oracle(fd, pd, case):
    from networks_fenicsx_b200 import elements
    from networks_fenicsx_b200.generic import RH_FLAG, VERTEX_FLAG, build_generic_system

    G = {"y": lambda: ng.make_tree(2, 1, 3), "random": lambda: helpers.random_tree(40, 2),
         "cyclic": helpers.edge_info_graph}[case]()
    N = 3
    nm = nxfx.NetworkMesh(G, N=N, color_strategy="largest_first")
    gs = build_generic_system(nm, fd, pd)
    col = rp.color_graph_literal(G, "largest_first")
    net = rp.OracleNetworkHO(*rp.graph_to_arrays(G, col), N, fd, pd)
    for mine, ref in zip(elements.tables(fd, pd), net.tables):
        np.testing.assert_allclose(mine, ref, rtol=0, atol=1e-14)
    rng = np.random.default_rng(0)
    R = rng.uniform(0.5, 2.0, N * G.number_of_edges())
    f = rng.normal(size=R.size)
    pbc = net.eval_pbc(lambda x: x[0] + 2 * x[1])
    A, b = net.assemble(pbc, R=R, f=f)
    assert gs.n_dofs == net.n_dofs and gs.block_sizes == net.block_sizes
    np.testing.assert_array_equal(gs.rowptr, A.indptr)
    np.testing.assert_array_equal(gs.colidx, A.indices)
    # evaluate the contribution lists on the host exactly as the device kernels do
    h = net.cell_lengths()
    vals = np.zeros(A.nnz)
    for s in range(2):
        sid, co = gs.src_id[:, s], gs.src_coef[:, s]
        on = sid >= 0
        cellid = np.where(on, sid & (RH_FLAG - 1), 0)
        term = np.where(on & ((sid & RH_FLAG) != 0), (R[cellid] * h[cellid]) * co, co)
        vals += np.where(on, term, 0.0)
    np.testing.assert_allclose(vals, A.data, rtol=1e-13, atol=5e-14)  # cancelling pairs: tables differ in the last bit
    bb = np.zeros(net.n_dofs)
    rows = np.repeat(np.arange(net.n_dofs), np.diff(gs.bsrc_ptr))
    isv = (gs.bsrc_id & VERTEX_FLAG) != 0
    idx = gs.bsrc_id & (VERTEX_FLAG - 1)
    term = np.where(isv, gs.bsrc_coef * pbc[np.minimum(idx, pbc.size - 1)],
                    (f[np.minimum(idx, f.size - 1)] * h[np.minimum(idx, h.size - 1)]) * gs.bsrc_coef)
    np.add.at(bb, rows, term)
    np.testing.assert_allclose(bb, b, rtol=1e-13, atol=5e-14)


def test_higher_order_known_answers():
    """SURVEY A.3 (sympy-checked): P2 mass (1/30)[[4,-1,2],[-1,4,2],[2,2,16]],
    B_ref = [[-5/6,1/6,2/3],[-1/6,5/6,-2/3]], int psi = [1/2,1/2]; and with f = 0 the P2/P1 flux
    coincides with the P1/DG0 flux (the exact flux is constant per edge)."""
    from networks_fenicsx_b200 import elements

    M, B, w, t0, t1 = elements.tables(2, 1)
    np.testing.assert_allclose(M * 30, [[4, -1, 2], [-1, 4, 2], [2, 2, 16]], atol=1e-13)
    np.testing.assert_allclose(B * 6, [[-5, 1, 4], [-1, 5, -4]], atol=1e-13)
    np.testing.assert_allclose(w, [0.5, 0.5], atol=1e-15)
    M1, B1, w1, _, _ = elements.tables(1, 0)
    np.testing.assert_allclose(M1 * 6, [[2, 1], [1, 2]], atol=1e-14)
    np.testing.assert_allclose(B1, [[-1, 1]], atol=1e-15)
    G = helpers.random_tree(30, 4)
    col = rp.color_graph_literal(G, "smallest_last")
    arrays = rp.graph_to_arrays(G, col)
    lo, ho = rp.OracleNetwork(*arrays, 2), rp.OracleNetworkHO(*arrays, 2, 2, 1)
    pbc = lo.eval_pbc(lambda x: x[1] - x[0])
    x_lo = lo.solve(*lo.assemble(pbc, R=1.7))
    x_ho = ho.solve(*ho.assemble(pbc, R=1.7))
    np.testing.assert_allclose(x_ho[ho.fb], x_lo[lo.fb], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(x_ho[ho.loff:], x_lo[lo.loff:], rtol=1e-10, atol=1e-13)


def test_schedule_of_tree_with_many_inlets_is_a_forest():
    """Regression (found by the property tests): a component with several inlets must be grown
    from ONE root, otherwise the edge where two BFS fronts meet is mistaken for a cycle."""
    A = ng.make_tree(7, 1, 1, as_arrays=True)
    rev = ng.ArrayGraph(A.pos, A.edges[:, ::-1].copy())  # every leaf becomes an inlet
    nm = nxfx.NetworkMesh(rev, N=2)
    assert nm._boundary_out_nodes.size == 64
    s = build_tree_schedule(nm.graph_edges, nm.node_multiplier_index, nm.bifurcation_values.size,
                            root_hint_nodes=nm._boundary_out_nodes, chunk_nodes=16)
    assert s.is_forest
    check_schedule(nm, s, 16)
