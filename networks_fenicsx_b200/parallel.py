"""Multi-GPU layer: one process per GPU (torchrun), edge partition of the network.

Round-1 scope (DESIGN.md "Multi-GPU"): the partition unit is a connected component -- a network
that is a forest is split into whole trees, one group of trees per GPU, so the edge partition has
zero cut bifurcations and the data path needs no collective; ``torch.distributed`` is only used for
metadata (sizes, timing, diagnostics).  Cutting a single tree across GPUs (subtree partition with a
halo of cut multipliers and an all-gathered top of the elimination tree) is the next step and is
described in DESIGN.md.
"""

from __future__ import annotations

import dataclasses

import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components

from .network_generation import ArrayGraph


class TorchDistComm:
    """``MPI.Comm`` look-alike on top of ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests)
    for the few collective calls user scripts make (demo_tree.py:64-71, test_orientation.py:50)."""

    def __init__(self, device=None):
        import torch.distributed as dist

        self._dist = dist
        self._device = device
        self.rank = dist.get_rank()
        self.size = dist.get_world_size()

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def allreduce(self, value, op="sum"):
        import torch

        name = getattr(op, "__name__", str(op)).lower()
        red = {"sum": self._dist.ReduceOp.SUM, "max": self._dist.ReduceOp.MAX, "min": self._dist.ReduceOp.MIN}[
            "max" if "max" in name else ("min" in name and "min" or "sum")]
        t = torch.tensor([float(value)], dtype=torch.float64, device=self._device)
        self._dist.all_reduce(t, op=red)
        return float(t.item())

    def bcast(self, obj, root=0):
        box = [obj]
        self._dist.broadcast_object_list(box, src=root)
        return box[0]

    def barrier(self):
        self._dist.barrier()


@dataclasses.dataclass
class LocalPart:
    graph: ArrayGraph
    global_nodes: np.ndarray  # local node -> global node
    global_edges: np.ndarray  # local edge -> global edge
    n_components: int


def partition_components(edges: np.ndarray, n_nodes: int, world_size: int) -> np.ndarray:
    """Rank of every graph edge: connected components are distributed over the ranks by greedy
    longest-processing-time bin packing on their edge counts (deterministic)."""
    adj = sp.coo_matrix((np.ones(edges.shape[0]), (edges[:, 0], edges[:, 1])), shape=(n_nodes, n_nodes))
    _, label = connected_components(adj, directed=False)
    comp_of_edge = label[edges[:, 0]]
    sizes = np.bincount(comp_of_edge)
    load = np.zeros(world_size, dtype=np.int64)
    rank_of_comp = np.zeros(sizes.size, dtype=np.int64)
    for comp in np.argsort(-sizes, kind="stable"):
        r = int(np.argmin(load))
        rank_of_comp[comp] = r
        load[r] += sizes[comp]
    return rank_of_comp[comp_of_edge]


def local_part(graph: ArrayGraph, rank_of_edge: np.ndarray, rank: int) -> LocalPart:
    """Sub-network owned by ``rank``: its edges in global order, nodes renumbered in ascending
    global order (so that every local array is a slice of the global canonical numbering)."""
    ge = np.flatnonzero(rank_of_edge == rank)
    edges = graph.edges[ge]
    gn = np.unique(edges)
    local_of = np.full(graph.number_of_nodes(), -1, dtype=np.int64)
    local_of[gn] = np.arange(gn.size)
    attrs = {k: v[ge] for k, v in graph.edge_attrs.items()}
    sub = ArrayGraph(graph.pos[gn], local_of[edges], attrs)
    adj = sp.coo_matrix((np.ones(ge.size), (sub.edges[:, 0], sub.edges[:, 1])), shape=(gn.size, gn.size))
    ncomp = connected_components(adj, directed=False)[0] if ge.size else 0
    return LocalPart(sub, gn, ge, ncomp)


def forest(trees: list[ArrayGraph]) -> ArrayGraph:
    """Disjoint union of networks (node and edge numbering concatenated)."""
    off = np.cumsum([0] + [t.number_of_nodes() for t in trees])
    pos = np.vstack([t.pos for t in trees])
    edges = np.vstack([t.edges + o for t, o in zip(trees, off[:-1])])
    keys = set.intersection(*[set(t.edge_attrs) for t in trees]) if trees else set()
    attrs = {k: np.concatenate([t.edge_attrs[k] for t in trees]) for k in keys}
    return ArrayGraph(pos, edges, attrs)
