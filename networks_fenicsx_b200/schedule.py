"""Elimination schedule of the bifurcation graph for the network Schur preconditioner.

Host-side symbolic analysis (NumPy, level-synchronous BFS): a spanning forest of the graph whose
nodes are the bifurcations and whose links are the graph edges joining two bifurcations, cut into
chunks that one thread block eliminates level by level (``precond.cuh``), plus a top chunk.
Graph edges that close a cycle are reported as chords.  The tables are uploaded once through
``nxfx_set_tree_schedule``.
"""

from __future__ import annotations

import dataclasses

import numpy as np

CHUNK_NODES = 2048  # nodes per bottom chunk (one thread block, 1024 threads)


@dataclasses.dataclass
class TreeSchedule:
    t_of_bif: np.ndarray
    t_parent: np.ndarray
    t_pedge: np.ndarray
    t_cptr: np.ndarray
    t_cidx: np.ndarray
    chunk_lptr: np.ndarray
    lvl_ptr: np.ndarray
    chord_edge: np.ndarray
    depth: np.ndarray  # per bifurcation (bif order)

    @property
    def n_chunks(self) -> int:
        return self.chunk_lptr.size - 1

    @property
    def is_forest(self) -> bool:
        return self.chord_edge.size == 0


def spanning_forest(edges: np.ndarray, node_lm: np.ndarray, n_bif: int,
                    root_hint_nodes: np.ndarray | None = None):
    """Level-synchronous BFS over the bifurcation graph.  Returns ``(parent, pedge, depth, chord)``:
    parent bifurcation (-1 for roots), graph edge to the parent, depth, and the graph edges between
    bifurcations that are not in the forest."""
    u, v = edges[:, 0], edges[:, 1]
    a, b = node_lm[u], node_lm[v]
    link = np.flatnonzero((a >= 0) & (b >= 0))
    la, lb = a[link].astype(np.int64), b[link].astype(np.int64)
    # symmetric adjacency CSR over bifurcation indices
    src = np.concatenate([la, lb])
    dst = np.concatenate([lb, la])
    eid = np.concatenate([link, link])
    order = np.argsort(src, kind="stable")
    src, dst, eid = src[order], dst[order], eid[order]
    ptr = np.concatenate([[0], np.cumsum(np.bincount(src, minlength=n_bif))])

    parent = np.full(n_bif, -1, dtype=np.int64)
    pedge = np.full(n_bif, -1, dtype=np.int64)
    depth = np.full(n_bif, -1, dtype=np.int64)

    def bfs(frontier, d0):
        depth[frontier] = d0
        d = d0
        while frontier.size:
            starts = ptr[frontier]
            counts = ptr[frontier + 1] - starts
            total = int(counts.sum())
            if total == 0:
                break
            idx = np.repeat(starts - (np.cumsum(counts) - counts), counts) + np.arange(total)
            nb, sr, ed = dst[idx], np.repeat(frontier, counts), eid[idx]
            m = depth[nb] < 0
            nb, sr, ed = nb[m], sr[m], ed[m]
            if nb.size == 0:
                break
            uniq, first = np.unique(nb, return_index=True)
            parent[uniq] = sr[first]
            pedge[uniq] = ed[first]
            d += 1
            depth[uniq] = d
            frontier = uniq

    if root_hint_nodes is not None and len(root_hint_nodes):
        hint = np.zeros(node_lm.size, dtype=bool)
        hint[root_hint_nodes] = True
        roots = np.unique(np.concatenate([b[hint[u] & (b >= 0)], a[hint[v] & (a >= 0)]])).astype(np.int64)
        # one root per connected component: a component with several inlets must not be grown from
        # several roots at once (the edge where two fronts meet would be mistaken for a chord)
        while roots.size:
            bfs(roots[:1], 0)
            roots = roots[depth[roots] < 0]
    while True:
        rest = np.flatnonzero(depth < 0)
        if rest.size == 0:
            break
        bfs(rest[:1], 0)

    tree_edge = np.zeros(edges.shape[0], dtype=bool)
    tree_edge[pedge[pedge >= 0]] = True
    chord = link[~tree_edge[link]].astype(np.int32)
    return parent, pedge, depth, chord


def assign_chunks(parent: np.ndarray, depth: np.ndarray, chunk_nodes: int = CHUNK_NODES):
    """Cut the forest into bottom chunks (bin-packed complete subtrees of <= ``chunk_nodes``
    nodes) and one top chunk (the nodes whose subtree is larger).  Returns ``(chunk, n_chunks)``
    with the top chunk last."""
    n_bif = parent.size
    size = np.ones(n_bif, dtype=np.int64)
    by_depth = np.argsort(depth, kind="stable")
    dsorted = depth[by_depth]
    bounds = np.flatnonzero(np.diff(dsorted)) + 1
    levels = np.split(by_depth, bounds)
    for lv in reversed(levels[1:]):
        np.add.at(size, parent[lv], size[lv])
    heavy = size > chunk_nodes
    if not heavy.any():
        return np.zeros(n_bif, dtype=np.int64), 1  # everything in the single (top) chunk
    # roots of the light subtrees hanging below heavy nodes (or light whole trees)
    par_heavy = np.where(parent >= 0, heavy[np.maximum(parent, 0)], True)
    croots = np.flatnonzero(~heavy & par_heavy)
    # bin-pack consecutive subtree roots into chunks of <= chunk_nodes nodes
    cum = np.cumsum(size[croots])
    bin_of_root = np.zeros(croots.size, dtype=np.int64)
    start, base, k = 0, 0, 0
    while start < croots.size:
        end = int(np.searchsorted(cum, base + chunk_nodes, side="right"))
        end = max(end, start + 1)
        bin_of_root[start:end] = k
        base = cum[end - 1]
        start = end
        k += 1
    n_bottom = k
    chunk = np.full(n_bif, n_bottom, dtype=np.int64)  # heavy nodes -> top chunk (last)
    chunk[croots] = bin_of_root
    for lv in levels:  # propagate chunk ids root -> leaf
        m = (~heavy[lv]) & (parent[lv] >= 0)
        sel = lv[m]
        inherit = ~heavy[parent[sel]]
        chunk[sel[inherit]] = chunk[parent[sel[inherit]]]
    return chunk, n_bottom + 1


def assemble_schedule(parent, pedge, depth, chunk, n_chunks, chord) -> TreeSchedule:
    """Order the nodes by (chunk, depth) and build the tables ``nxfx_set_tree_schedule`` takes."""
    i32 = np.int32
    n_bif = parent.size
    t_order = np.lexsort((depth, chunk))
    t_of_bif = np.empty(n_bif, dtype=np.int64)
    t_of_bif[t_order] = np.arange(n_bif)
    ck, dp = chunk[t_order], depth[t_order]
    change = np.flatnonzero((np.diff(ck) != 0) | (np.diff(dp) != 0)) + 1
    lvl_ptr = np.concatenate([[0], change, [n_bif]])
    lvl_chunk = ck[lvl_ptr[:-1]]
    chunk_lptr = np.concatenate([[0], np.cumsum(np.bincount(lvl_chunk, minlength=n_chunks))])
    t_parent = np.where(parent[t_order] >= 0, t_of_bif[np.maximum(parent[t_order], 0)], -1)
    t_pedge = pedge[t_order]
    has_p = np.flatnonzero(t_parent >= 0)
    corder = has_p[np.argsort(t_parent[has_p], kind="stable")]
    t_cptr = np.concatenate([[0], np.cumsum(np.bincount(t_parent[has_p], minlength=n_bif))])
    return TreeSchedule(
        t_of_bif.astype(i32), t_parent.astype(i32), t_pedge.astype(i32), t_cptr.astype(i32),
        corder.astype(i32), chunk_lptr.astype(i32), lvl_ptr.astype(i32), np.asarray(chord, dtype=i32),
        depth.astype(i32),
    )


def build_tree_schedule(edges: np.ndarray, node_lm: np.ndarray, n_bif: int,
                        root_hint_nodes: np.ndarray | None = None,
                        chunk_nodes: int = CHUNK_NODES) -> TreeSchedule:
    """``edges[E,2]`` graph edges, ``node_lm[n_nodes]`` multiplier index or -1.
    ``root_hint_nodes``: graph nodes (inlets) whose neighbouring bifurcations become roots."""
    i32 = np.int32
    if n_bif == 0:
        z = np.zeros(0, dtype=i32)
        return TreeSchedule(z, z, z, np.zeros(1, dtype=i32), z, np.zeros(1, dtype=i32), np.zeros(1, dtype=i32), z, z)
    parent, pedge, depth, chord = spanning_forest(edges, node_lm, n_bif, root_hint_nodes)
    chunk, n_chunks = assign_chunks(parent, depth, chunk_nodes)
    sched = assemble_schedule(parent, pedge, depth, chunk, n_chunks, chord)
    if chunk_nodes == CHUNK_NODES and _top_size(sched) > CHUNK_NODES:
        # very large trees: the top chunk outgrows a 2048-node block; 4096-node chunks halve it
        # (the device kernels accept both capacities)
        chunk, n_chunks = assign_chunks(parent, depth, 2 * CHUNK_NODES)
        bigger = assemble_schedule(parent, pedge, depth, chunk, n_chunks, chord)
        if _top_size(bigger) <= 2 * CHUNK_NODES:
            sched = bigger
    return sched


def _top_size(s: TreeSchedule) -> int:
    return int(s.lvl_ptr[s.chunk_lptr[-1]] - s.lvl_ptr[s.chunk_lptr[-2]])
