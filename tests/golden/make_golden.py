"""Generate golden fixtures by importing the REFERENCE's own code (run in the build container,
where /root/reference exists; the GPU box only reads the committed .npz files).

Only ``network_generation.py`` of the reference is importable here (DOLFINx/PETSc are absent);
``dolfinx.common.timed`` is stubbed with an identity decorator.  The fixtures pin
``make_tree`` / ``make_arterial_tree`` (node positions, edge lists in ``graph.edges()`` order, edge
radii) and networkx's greedy edge colouring as called by mesh.py:38-39.

    python tests/golden/make_golden.py
"""

import importlib.util
import pathlib
import sys
import types

import networkx as nx
import numpy as np

HERE = pathlib.Path(__file__).parent
REF = pathlib.Path("/root/reference/src/networks_fenicsx/network_generation.py")


def load_reference_generators():
    dolfinx = types.ModuleType("dolfinx")
    common = types.ModuleType("dolfinx.common")
    common.timed = lambda name: (lambda fn: fn)
    dolfinx.common = common
    sys.modules.setdefault("dolfinx", dolfinx)
    sys.modules.setdefault("dolfinx.common", common)
    spec = importlib.util.spec_from_file_location("reference_network_generation", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def graph_arrays(G, with_radius=False):
    pos = np.asarray([G.nodes[v]["pos"] for v in G.nodes()], dtype=np.float64)
    edges = np.asarray(list(G.edges()), dtype=np.int64)
    out = {"pos": pos, "edges": edges, "nodes": np.asarray(list(G.nodes()), dtype=np.int64)}
    if with_radius:
        out["radius"] = np.asarray([G.edges[e]["radius"] for e in G.edges()], dtype=np.float64)
    return out


def reference_coloring(G, strategy):
    """mesh.py:38-39 verbatim call sequence; returns colour per edge in graph.edges() order."""
    line = nx.line_graph(G.to_undirected())
    col = nx.coloring.greedy_color(line, strategy=strategy)
    return np.asarray(
        [col[(u, v)] if (u, v) in col else col[(v, u)] for u, v in G.edges()], dtype=np.int32
    )


def main():
    ref = load_reference_generators()
    data = {}
    for n in (1, 2, 3, 5, 7, 10):
        for H, W in ((1, 1), (2, 1), (1, 3), (3.1, 7.3)):
            for dim in (2, 3):
                if n == 1:
                    continue  # reference divides by zero for n == 1 (x_offset), not a valid input
                G = ref.make_tree(n, H, W, dim)
                key = f"tree_n{n}_H{H}_W{W}_d{dim}"
                for k, v in graph_arrays(G).items():
                    data[f"{key}/{k}"] = v
    for n in (3, 5, 7):
        G = ref.make_tree(n, 1, 1)
        for strat in ("smallest_last", "largest_first"):
            data[f"color_tree_n{n}_{strat}"] = reference_coloring(G, strat)
    for N, direction, gamma in ((1, [0, 1, 0], 0.8), (3, [0, 1, 0], 0.8), (5, [0.1, 1, 0], 0.8), (6, [1, 1, 0], 0.5)):
        G = ref.make_arterial_tree(N=N, direction=np.array(direction, dtype=np.float64), gamma=gamma)
        key = f"arterial_N{N}_g{gamma}_d{'_'.join(str(x) for x in direction)}"
        for k, v in graph_arrays(G, with_radius=True).items():
            data[f"{key}/{k}"] = v
        if N >= 3:
            data[f"color_{key}_largest_first"] = reference_coloring(G, "largest_first")
    np.savez_compressed(HERE / "reference_graphs.npz", **data)
    print(f"wrote {len(data)} arrays to {HERE / 'reference_graphs.npz'}")


def main_large():
    """``--large``: networkx's own ``smallest_last`` colouring (mesh.py:38-39) of the 16- and the
    20-generation ``make_tree`` (the headline graph; 84 s and ~3 GB), one int8 per edge in edge order ->
    reference_colors_large.npz.  Pins the networkx-identical colouring NetworkMesh uses at scale."""
    ref = load_reference_generators()
    out = {}
    for n in (16, 20):
        G = ref.make_tree(n, n, n)
        out[f"color_tree_n{n}_smallest_last"] = reference_coloring(G, "smallest_last").astype(np.int8)
    np.savez_compressed(HERE / "reference_colors_large.npz", **out)


if __name__ == "__main__":
    main_large() if "--large" in sys.argv else main()
