"""Network generators with the interface of the reference's ``network_generation`` module
(network_generation.py:41-100 ``make_tree``, :157-283 ``make_arterial_tree``).

They return the same ``networkx.DiGraph`` (node ids, ``"pos"``, edge order, ``"radius"``) as the
reference -- pinned bit for bit by tests/golden/reference_graphs.npz -- but are built from array
formulas, so a 20+ generation tree takes milliseconds instead of minutes.  With
``as_arrays=True`` they return an :class:`ArrayGraph`, a networkx-free container that
``NetworkMesh`` accepts directly (building a ``DiGraph`` with millions of nodes costs seconds and
gigabytes for no benefit).
"""

from __future__ import annotations

import dataclasses
from typing import Callable

import numpy as np
import numpy.typing as npt

from .common import timed

__all__ = ["make_tree", "make_arterial_tree", "tree_edges", "ArrayGraph"]


@dataclasses.dataclass
class ArrayGraph:
    """Directed graph as arrays: ``pos[n_nodes, gdim]`` (node ``i`` = row ``i``), ``edges[E, 2]``
    in ``graph.edges()`` order, optional per-edge attributes (e.g. ``"radius"``)."""

    pos: npt.NDArray[np.float64]
    edges: npt.NDArray[np.int64]
    edge_attrs: dict[str, npt.NDArray] = dataclasses.field(default_factory=dict)

    def number_of_nodes(self) -> int:
        return int(self.pos.shape[0])

    def number_of_edges(self) -> int:
        return int(self.edges.shape[0])

    def to_networkx(self):
        import networkx as nx

        G = nx.DiGraph()
        G.add_nodes_from(range(self.number_of_nodes()))
        for i, p in enumerate(self.pos):
            G.nodes[i]["pos"] = p
        G.add_edges_from(map(tuple, self.edges.tolist()))
        for name, values in self.edge_attrs.items():
            for (u, v), val in zip(self.edges.tolist(), values.tolist()):
                G.edges[(u, v)][name] = val
        return G


def _tree_arrays(n: int, H: float, W: float, dim: int):
    """Heap-numbered symmetric binary tree: node 0 is the inlet, node 1 the first junction, node
    k >= 2 hangs below node k // 2; generation g (1-based depth) occupies nodes 2**g .. 2**(g+1)-1
    (network_generation.py:18-38, 55-99)."""
    assert n >= 1, "Number of generations must be at least 1"
    nb_nodes = 2**n
    nb_last = 2 ** (n - 1)
    x_offset = W / (2 * (nb_last - 1))  # n == 1 divides by zero exactly like the reference
    y_offset = H / n
    pos = np.zeros((nb_nodes, dim), dtype=np.float64)
    pos[1, 1] = y_offset
    for gen in range(1, n):
        factor = 2 ** (n - gen)
        half = 2**gen // 2
        # the reference accumulates x += x_offset * factor; cumsum reproduces that rounding
        inc = np.full(half, x_offset * factor)
        inc[0] = x_offset * (factor / 2)
        xs = np.cumsum(inc)
        first = 2**gen
        pos[first : first + half, 0] = -xs[::-1]
        pos[first + half : first + 2 * half, 0] = xs
        pos[first : first + 2 * half, 1] = y_offset * (gen + 1)
    child = np.arange(1, nb_nodes, dtype=np.int64)
    parent = child // 2
    edges = np.stack([parent, child], axis=1)
    return pos, edges


def tree_edges(n: int, r: int):
    """Edges of the rooted tree at 0 with ``n`` nodes and branching ratio ``r`` in the order the reference's
    generator yields them (network_generation.py:18-38): the root branch ``(0, 1)``, then every node ``k >= 2``
    hangs below ``1 + (k - 2) // r`` -- closed form instead of the parent stack."""
    if n == 0:
        return
    yield 0, 1
    if n > 2 and r > 0:
        k = np.arange(2, n)
        for s, t in zip((1 + (k - 2) // r).tolist(), k.tolist()):
            yield s, t


@timed("nxfx:make_tree")
def make_tree(n: int, H: float, W: float, dim=3, as_arrays: bool = False):
    """Symmetric binary tree with ``n`` generations, height ``H``, width ``W``
    (network_generation.py:41-100)."""
    pos, edges = _tree_arrays(n, H, W, dim)
    if as_arrays:
        return ArrayGraph(pos, edges)
    import networkx as nx

    G = nx.DiGraph()
    G.add_nodes_from(range(pos.shape[0]))
    # the reference stores python lists whose entries are ints where no arithmetic happened
    for i, p in enumerate(pos.tolist()):
        G.nodes[i]["pos"] = p
    G.add_edges_from(map(tuple, edges.tolist()))
    return G


def _default_normal(x: npt.NDArray[np.floating]) -> npt.NDArray[np.floating]:
    """Normal of the xy-plane (network_generation.py:103-107)."""
    out = np.zeros_like(x)
    out[2] = 1
    return out


def _endpoint(pm1, p0, normal, angle_deg, length):
    """End point of a daughter vessel: project the parent direction onto the plane with normal
    ``normal``, rotate it by ``angle_deg`` about the normal (Rodrigues) and advance ``length``
    (network_generation.py:110-154)."""
    prev = p0 - pm1
    nn = np.linalg.norm(normal)
    d = np.dot(prev, normal) / nn
    in_plane = prev - d * normal / nn
    theta = np.radians(angle_deg)
    k = normal / np.linalg.norm(normal)
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    rot = np.eye(3) + np.sin(theta) * K + (1 - np.cos(theta)) * np.dot(K, K)
    newdir = np.dot(rot, in_plane)
    return p0 + length * newdir / np.linalg.norm(newdir, axis=-1)


def _arterial_generations_loop(N, pos, radius, edges, lmbda, gamma, normal, random):
    """Vessel by vessel, as the reference does (any ``normal`` callable, random branching sides)."""
    inode = 1
    previous = [0]  # edge indices of the previous generation
    for _ in range(1, N):
        current = []
        for e in previous:
            a, b = edges[e]
            Dp = radius[e] * 2
            D2 = Dp * (gamma**3 + 1) ** (-1 / 3)
            D1 = gamma * D2
            cos1 = (Dp**4 + D1**4 - (Dp**3 - D1**3) ** (4 / 3)) / (2 * Dp**2 * D1**2)
            cos2 = (Dp**4 + D2**4 - (Dp**3 - D2**3) ** (4 / 3)) / (2 * Dp**2 * D2**2)
            ang1, ang2 = np.degrees(np.arccos(cos1)), np.degrees(np.arccos(cos2))
            sign1 = 1 if not random else np.random.choice([-1, 1])
            nrm = normal(pos[b])
            for sign, ang, D in ((sign1, ang1, D1), (-sign1, ang2, D2)):
                inode += 1
                edges[inode - 1] = (b, inode)
                pos[inode] = _endpoint(pos[a], pos[b], nrm, sign * ang, lmbda * D)
                radius[inode - 1] = D / 2
                current.append(inode - 1)
        previous = current


def _scalar_power(x: np.ndarray, p: float) -> np.ndarray:
    """``x ** p`` element by element with the SCALAR power (what the reference's per-vessel code runs);
    a generation holds few distinct diameters, so the scalar work is done once per distinct value."""
    uniq, inv = np.unique(x, return_inverse=True)
    return np.array([v**p for v in uniq], dtype=np.float64)[inv]


def _arterial_generations_vectorised(N, pos, radius, edges, lmbda, gamma):
    """All vessels of a generation at once (default surface normal e_z, deterministic sides).  Generation
    g >= 2 holds the edges 2**(g-1)-1 .. 2**g-2; the daughters of the i-th edge of generation g-1 are the
    nodes 2**(g-1)+2i and 2**(g-1)+2i+1 (creation order of the reference's loop).

    Everything elementwise is done on whole generations with the reference's operations in the
    reference's order (with the normal (0, 0, 1) the projection and ``K = [[0,-1,0],[1,0,0],[0,0,0]]``
    are exact).  The one BLAS call of ``_endpoint`` -- the 3x3 matrix-vector product ``np.dot(rot, in_plane)``,
    whose rounding depends on the BLAS kernel (FMA or not) -- is still issued per vessel, so the positions
    are bit-identical to the loop form on any machine (tests/test_host_logic.py; 1 M vessels in ~3 s
    instead of 35 s)."""
    for g in range(2, N + 1):
        cnt = 2 ** (g - 2)
        par = np.arange(cnt) + (cnt - 1)
        a, b = edges[par, 0], edges[par, 1]
        Dp = radius[par] * 2
        D2 = Dp * (gamma**3 + 1) ** (-1 / 3)
        D1 = gamma * D2
        P = _scalar_power  # numpy rounds ``array ** p`` (SIMD pow) and ``scalar ** p`` (libm pow) differently
        cos1 = (P(Dp, 4) + P(D1, 4) - P(P(Dp, 3) - P(D1, 3), 4 / 3)) / (2 * P(Dp, 2) * P(D1, 2))
        cos2 = (P(Dp, 4) + P(D2, 4) - P(P(Dp, 3) - P(D2, 3), 4 / 3)) / (2 * P(Dp, 2) * P(D2, 2))
        ang1, ang2 = np.degrees(np.arccos(cos1)), np.degrees(np.arccos(cos2))
        prev = pos[b] - pos[a]
        in_plane = prev.copy()
        in_plane[:, 2] = prev[:, 2] - prev[:, 2]  # x - (x . n) n with n = (0, 0, 1)
        node0 = 2 ** (g - 1) + 2 * np.arange(cnt)
        for k, (sgn, ang, D) in enumerate(((1.0, ang1, D1), (-1.0, ang2, D2))):
            theta = np.radians(sgn * ang)
            s_, c_ = np.sin(theta), 1 - np.cos(theta)
            rot = np.zeros((cnt, 3, 3))  # I + s K + c K^2, K^2 = diag(-1, -1, 0)
            rot[:, 0, 0] = 1.0 + c_ * -1.0
            rot[:, 1, 1] = 1.0 + c_ * -1.0
            rot[:, 2, 2] = 1.0
            rot[:, 0, 1] = s_ * -1.0
            rot[:, 1, 0] = s_
            newdir = np.empty((cnt, 3))
            dot = np.dot
            for i in range(cnt):
                newdir[i] = dot(rot[i], in_plane[i])
            nn = np.sqrt(newdir[:, 0] * newdir[:, 0] + newdir[:, 1] * newdir[:, 1] + newdir[:, 2] * newdir[:, 2])
            length = lmbda * D
            node = node0 + k
            pos[node] = pos[b] + length[:, None] * newdir / nn[:, None]
            edges[node - 1, 0] = b
            edges[node - 1, 1] = node
            radius[node - 1] = D / 2


@timed("nxfx:make_arterial_tree")
def make_arterial_tree(
    N: int,
    p0: npt.NDArray[np.floating] = np.zeros(3, dtype=np.float64),
    direction: npt.NDArray[np.floating] = np.array([0, 1, 0], dtype=np.float64),
    D0: float = 2.0,
    lmbda: float = 8.0,
    gamma: float = 0.8,
    normal: Callable[[npt.NDArray[np.floating]], npt.NDArray[np.floating]] = _default_normal,
    random: bool = False,
    as_arrays: bool = False,
):
    """Murray-law arterial tree with ``N`` generations (network_generation.py:157-283): daughter
    diameters ``D2 = D0 (gamma^3+1)^(-1/3)``, ``D1 = gamma D2``, lengths ``lmbda * D``, bifurcation
    angles from the minimum-energy relation, per-edge ``"radius"`` attribute."""
    if gamma > 1:
        raise ValueError("Please choose a gamma lower or equal to 1")
    n_nodes = 2**N
    pos = np.empty((n_nodes, 3), dtype=np.float64)
    radius = np.empty(n_nodes - 1, dtype=np.float64)
    edges = np.empty((n_nodes - 1, 2), dtype=np.int64)
    pos[0] = p0
    pos[1] = p0 + (D0 * lmbda) * direction / np.linalg.norm(direction, axis=-1)
    edges[0] = (0, 1)
    radius[0] = D0 / 2
    if normal is _default_normal and not random:
        _arterial_generations_vectorised(N, pos, radius, edges, lmbda, gamma)
    else:
        _arterial_generations_loop(N, pos, radius, edges, lmbda, gamma, normal, random)
    if as_arrays:
        return ArrayGraph(pos, edges, {"radius": radius})
    import networkx as nx

    G = nx.DiGraph()
    for (u, v), r in zip(edges.tolist(), radius.tolist()):
        G.add_edge(u, v, radius=r)
    for i in range(n_nodes):
        G.nodes[i]["pos"] = pos[i].copy()
    return G
