"""Edge partition of ONE network over several GPUs (north_star "each GPU owns an edge
partition; NCCL carries the bifurcation-node halo exchange and the all-reduces").

Partition = the elimination schedule's own structure (``schedule.py``): the bottom chunks
(complete subtrees) are dealt to the ranks in order, every graph edge follows its deeper
bifurcation, and the multipliers of the top chunk -- the cut bifurcations -- are REPLICATED on all
ranks.  Everything that couples ranks is additive and small (n_top doubles):

* ``y = A x``: the shared multiplier rows are partial sums -> one all-reduce;
* preconditioner: the top chunk's partial diagonal / right-hand side (contributions of each rank's
  chunk roots and incident edges) -> one all-reduce, then every rank eliminates the identical top
  chunk redundantly and back-substitutes its own chunks;
* norms: owned entries only, one all-reduce of the scalars.

Host code (NumPy) in this file; collectives through ``torch.distributed`` (NCCL on GPUs, gloo in
the CPU tests).
"""

from __future__ import annotations

import dataclasses

import numpy as np

from .network_generation import ArrayGraph
from .schedule import CHUNK_NODES, TreeSchedule, assemble_schedule, assign_chunks, spanning_forest


@dataclasses.dataclass
class TreePartition:
    rank: int
    world: int
    graph: ArrayGraph  # local sub-network (nodes: endpoints of the local edges + all cut nodes)
    node_degree: np.ndarray  # GLOBAL degree of every local node
    global_nodes: np.ndarray  # local node -> global node
    global_edges: np.ndarray  # local edge -> global edge
    global_bif: np.ndarray  # local multiplier -> global multiplier index
    shared_lm: np.ndarray  # local multiplier indices of the replicated (shared) multipliers
    lam_weight: np.ndarray  # 1.0 where this rank counts the multiplier row in norms / -r_lambda
    schedule: TreeSchedule  # local elimination schedule; top chunk (this rank's private heavy nodes + the shared ones) last
    n_top: int  # number of SHARED multipliers (the exchanged part of the top chunk), identical on all ranks
    n_global_bif: int


def partition_tree(graph: ArrayGraph, world: int, rank: int, chunk_nodes: int = CHUNK_NODES) -> TreePartition:
    """Deterministic: every rank calls this with the same graph and gets its own part."""
    edges = np.asarray(graph.edges, dtype=np.int64)
    n_nodes = graph.number_of_nodes()
    u, v = edges[:, 0], edges[:, 1]
    deg_out = np.bincount(u, minlength=n_nodes)
    deg_in = np.bincount(v, minlength=n_nodes)
    degree = deg_in + deg_out
    bif = np.flatnonzero(degree > 1)
    n_bif = bif.size
    lm = np.full(n_nodes, -1, dtype=np.int64)
    lm[bif] = np.arange(n_bif)
    inlets = np.flatnonzero((degree == 1) & (deg_out == 1))
    parent, pedge, depth, chord = spanning_forest(edges, lm, n_bif, inlets)
    if chord.size:
        raise NotImplementedError("the multi-GPU partition needs a forest (no cycles among the bifurcations)")
    cn = chunk_nodes
    while True:
        chunk, n_chunks = assign_chunks(parent, depth, cn)
        if n_chunks - 1 >= world or cn <= 2:
            break
        cn //= 2
    n_bottom = n_chunks - 1
    if n_bottom < world:
        raise ValueError(f"network too small to cut into {world} parts")
    top = chunk == n_bottom
    rank_of_chunk = (np.arange(n_bottom) * world) // n_bottom
    # Heavy (top-chunk) nodes come in two kinds.  A heavy node all of whose bottom chunks went to ONE rank is
    # PRIVATE to that rank: it joins that rank's top chunk and nobody else sees it.  Only the heavy nodes
    # whose subtree spans several ranks are SHARED (replicated, exchanged): world - 1 nodes for a balanced
    # binary tree instead of the whole top of the elimination tree (7 instead of 2047 on 8 GPUs), so the
    # redundant top-chunk work and the exchanged payload stay constant as the number of GPUs grows.
    rmin = np.full(n_bif, world, dtype=np.int64)
    rmax = np.full(n_bif, -1, dtype=np.int64)
    rmin[~top] = rmax[~top] = rank_of_chunk[chunk[~top]]
    by_depth = np.argsort(depth, kind="stable")
    bounds = np.flatnonzero(np.diff(depth[by_depth])) + 1
    for lv in reversed(np.split(by_depth, bounds)):
        lv = lv[parent[lv] >= 0]
        np.minimum.at(rmin, parent[lv], rmin[lv])
        np.maximum.at(rmax, parent[lv], rmax[lv])
    shared_bif = top & (rmin != rmax)
    # shared nodes (and the edges between two of them) -> rank 0; private heavy nodes -> their rank
    owner_bif = np.where(shared_bif, 0, np.where(top, rmin, rank_of_chunk[np.minimum(chunk, n_bottom - 1)]))
    # every edge follows its deeper bifurcation
    a, b = lm[u], lm[v]
    da = np.where(a >= 0, depth[np.maximum(a, 0)], -1)
    db = np.where(b >= 0, depth[np.maximum(b, 0)], -1)
    deeper = np.where(db >= da, b, a)
    owner_edge = np.where(deeper >= 0, owner_bif[np.maximum(deeper, 0)], 0)
    ge = np.flatnonzero(owner_edge == rank)
    # local nodes: endpoints of the local edges and ALL shared nodes, ascending global id
    gn = np.unique(np.concatenate([edges[ge].ravel(), bif[shared_bif]]))
    local_of = np.full(n_nodes, -1, dtype=np.int64)
    local_of[gn] = np.arange(gn.size)
    attrs = {k: np.asarray(val)[ge] for k, val in graph.edge_attrs.items()}
    sub = ArrayGraph(np.asarray(graph.pos)[gn], local_of[edges[ge]], attrs)
    # local multipliers (ascending global node id, as NetworkMesh numbers them)
    lbif_nodes = gn[degree[gn] > 1]
    gb = lm[lbif_nodes]  # global multiplier index of every local multiplier
    lb_of_gb = np.full(n_bif, -1, dtype=np.int64)
    lb_of_gb[gb] = np.arange(gb.size)
    mine = shared_bif[gb] | (owner_bif[gb] == rank)
    if not np.all(mine):
        # a bifurcation of another rank's chunk can only appear here as the far end of one of our
        # edges, which the "deeper bifurcation" rule excludes on forests
        raise AssertionError("partition invariant violated")
    lpar = np.where(parent[gb] >= 0, lb_of_gb[np.maximum(parent[gb], 0)], -1)
    assert np.all((parent[gb] < 0) | (lpar >= 0)), "parent of a local multiplier is not local"
    ledge_of = np.full(edges.shape[0], -1, dtype=np.int64)
    ledge_of[ge] = np.arange(ge.size)
    lpedge = np.where(pedge[gb] >= 0, ledge_of[np.maximum(pedge[gb], 0)], -1)
    # my bottom chunks renumbered 0..k-1, the top chunk last
    my_chunks = np.flatnonzero(rank_of_chunk == rank)
    lchunk_of = np.full(n_chunks, -1, dtype=np.int64)
    lchunk_of[my_chunks] = np.arange(my_chunks.size)
    lchunk_of[n_bottom] = my_chunks.size
    lchunk = lchunk_of[chunk[gb]]
    assert np.all(lchunk >= 0)
    sched = assemble_schedule(lpar, lpedge, depth[gb], lchunk, my_chunks.size + 1, np.zeros(0, dtype=np.int32))
    shared = np.flatnonzero(shared_bif[gb])
    weight = np.where(shared_bif[gb] & (rank != 0), 0.0, 1.0)
    return TreePartition(rank, world, sub, degree[gn], gn, ge, gb, shared.astype(np.int32), weight, sched,
                         int(shared_bif.sum()), n_bif)


class DistributedSolver:
    """Assemble + solve of ONE network cut over the ranks of ``torch.distributed`` (one process per
    GPU).  Mirrors ``Solver`` for the default options (network-Schur direct solve + iterative
    refinement); the collectives are three kinds of small SUM all-reduces (see module docstring).

    Args:
        graph: the GLOBAL network (every rank passes the same :class:`ArrayGraph`).
        N: cells per graph edge.
        p_bc_ex, f, R: as :meth:`HydraulicNetworkAssembler.compute_forms`; ``R``/``f`` arrays are given for
            the GLOBAL network (per graph edge or per cell) and restricted to the local edges here.
        exchange: ``"peer"``: the kernels exchange the partial sums themselves over NVLink (CUDA IPC peer
            memory, ``peer.cuh``) -- the single-GPU launch sequence, no NCCL call, one host sync per solve
            (needs every chunk of this rank co-resident in the cooperative tree kernel: up to ~21 generations
            per GPU); ``"nccl"``: split phases around ``torch.distributed`` all-reduces;
            ``"auto"``: peer when available.
        device: CUDA ordinal of this rank.
    """

    def __init__(self, graph: ArrayGraph, N: int, p_bc_ex, f=None, R=None, device: int = 0,
                 color_strategy="smallest_last", chunk_nodes: int = CHUNK_NODES, group=None,
                 exchange: str = "auto"):
        import torch.distributed as dist

        from .assembly import HydraulicNetworkAssembler
        from .mesh import NetworkMesh, SerialComm
        from .solver import Solver

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        self.part = part = partition_tree(graph, world, rank, chunk_nodes)
        # the sub-network is handed over explicitly (comm = serial): no second partitioning inside
        mesh = NetworkMesh(part.graph, N=N, color_strategy=color_strategy, device=device,
                           node_degree=part.node_degree, comm=SerialComm())
        assembler = HydraulicNetworkAssembler(mesh)
        assembler.compute_forms(p_bc_ex=p_bc_ex, f=self._restrict(f, N, graph), R=self._restrict(R, N, graph))
        solver = Solver(assembler, schedule=part.schedule)
        self._attach(part, mesh, assembler, solver, group, exchange)

    @classmethod
    def from_solver(cls, solver, part: TreePartition, group=None, exchange: str = "auto") -> "DistributedSolver":
        """Distributed layer for an existing ``Solver`` on one rank's part of a partitioned network
        (``NetworkMesh(G, N, comm=TorchDistComm())`` under torchrun: the reference-API path)."""
        self = cls.__new__(cls)
        self._attach(part, solver.assembler.network, solver.assembler, solver, group, exchange)
        return self

    def _attach(self, part: TreePartition, mesh, assembler, solver, group, exchange: str) -> None:
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib

        self._C, self._torch, self._dist, self._group = C, torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.part = part
        self.mesh, self.assembler, self.solver = mesh, assembler, solver
        N = mesh.cells_per_edge
        device = mesh.device.index
        assert np.array_equal(part.global_nodes[self.mesh.bifurcation_values], part.global_nodes[
            np.flatnonzero(part.node_degree > 1)])
        self.dev = self.mesh.device
        shared = np.ascontiguousarray(part.shared_lm, dtype=np.int32)
        weight = np.ascontiguousarray(part.lam_weight, dtype=np.float64)
        self.dev.call("nxfx_set_shared", shared.size, _lib.as_i32p(shared), _lib.as_f64p(weight))
        cuda = torch.device("cuda", device)
        self._top = torch.zeros(3 * max(part.n_top, 1), dtype=torch.float64, device=cuda)  # [setup 2n | apply n]
        self._sh = torch.zeros(shared.size + 2, dtype=torch.float64, device=cuda)  # halo rows + 2 norm partials
        self._nrm = torch.zeros(2 * 40, dtype=torch.float64, device=cuda)
        self._r = self.dev.empty(self.assembler.num_dofs)
        self.n_dofs_global = int(self._allreduce_scalar(self._owned_dofs()))
        self.history: list[float] = []
        self.rhs_norm = float("nan")
        self.corrections = 0
        self.exchange = "nccl"
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError("exchange must be 'auto', 'peer' or 'nccl'")
        if exchange != "nccl":
            self._connect_peers(required=exchange == "peer")

    def _connect_peers(self, required: bool) -> None:
        """Peer exchange over NVLink (``peer.cuh``): every rank exports its exchange buffer as a CUDA IPC
        handle, the handles are all-gathered with torch.distributed (the only use of the process group
        on this path) and mapped.  All ranks agree on the outcome; any failure falls back to the
        host-driven NCCL all-reduces unless ``required``."""
        C, torch, dist = self._C, self._torch, self._dist
        handle = (C.c_ubyte * 64)()
        slot = C.c_int32(0)
        ok = 1
        try:
            self.dev.call("nxfx_comm_create", self.rank, self.world, C.cast(handle, C.c_void_p), C.byref(slot))
        except RuntimeError as exc:  # e.g. the schedule does not fit the cooperative kernel
            ok, err = 0, str(exc)
        cuda = self._top.device
        mine = torch.tensor(list(bytes(handle)) + [slot.value & 0xFF, (slot.value >> 8) & 0xFF, (slot.value >> 16) & 0xFF, ok],
                            dtype=torch.uint8, device=cuda)
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=self._group)
        table = np.stack([t.cpu().numpy() for t in allh])
        if ok and table[:, 67].all() and (table[:, 64:67] == table[0, 64:67]).all():
            blob = np.ascontiguousarray(table[:, :64]).tobytes()
            try:
                self.dev.call("nxfx_comm_connect", C.c_char_p(blob))
            except RuntimeError as exc:
                ok, err = 0, str(exc)
        else:
            ok, err = 0, "a rank could not create its exchange buffer (or the slot sizes differ)"
        if self._allreduce_scalar(float(ok)) == float(self.world):
            self.exchange = "peer"
            dist.barrier(group=self._group)  # every buffer is mapped and zeroed before the first flag is written
            return
        self.dev.call("nxfx_comm_destroy")
        if required:
            raise RuntimeError(f"peer exchange unavailable: {err if not ok else 'another rank failed'}")

    def _restrict(self, coef, N: int, graph: ArrayGraph):
        """Per-edge / per-cell coefficient arrays of the GLOBAL network -> this rank's edges (arrays that
        already have the local size, scalars and None pass through)."""
        if coef is None or np.isscalar(coef):
            return coef
        arr = np.asarray(coef, dtype=np.float64)
        E = graph.number_of_edges()
        ge = self.part.global_edges
        if arr.shape == (E,) and ge.size != E:
            return arr[ge]
        if arr.shape == (E * N,) and ge.size != E:
            return arr.reshape(E, N)[ge].ravel()
        return arr

    def _owned_dofs(self) -> int:
        return int(self.assembler.num_dofs - self.part.n_top + (self.part.n_top if self.rank == 0 else 0))

    def _allreduce_scalar(self, v: float) -> float:
        t = self._torch.tensor([float(v)], dtype=self._torch.float64, device=self._top.device)
        self._dist.all_reduce(t, group=self._group)
        return float(t.item())

    def _ptr(self, t, offset: int = 0):
        return self._C.c_void_p(t.data_ptr() + 8 * offset)

    def assemble(self) -> None:
        self.solver.assemble()

    def _apply(self, r_ptr, z_ptr, add: int) -> None:
        top = self._ptr(self._top)
        self.dev.call("nxfx_pc_apply_begin", r_ptr, top)
        self._dist.all_reduce(self._top[: self.part.n_top], group=self._group)
        self.dev.call("nxfx_pc_apply_end", r_ptr, z_ptr, top, add)

    def solve(self, refine_steps: int = 1, final_residual: bool = False, refine_rtol: float = 1e-13):
        """x = P^{-1} b, then iterative refinement while ``||b - A x|| > refine_rtol ||b||`` (at most
        ``refine_steps`` corrections; the all-reduced norms are identical on every rank, so all
        ranks take the same decision).  Returns the global relative residual norms that were
        evaluated.  Collectives: 1 (setup fused with the first application) + 1 per residual (halo rows
        and norm partials in one buffer) + 1 per correction."""
        dev = self.dev
        b = self.solver.b.device_ptr()
        x = self.solver.x.device_ptr_overwrite()
        if self.exchange == "peer":
            return self._solve_peer(b, x, refine_steps, final_residual, refine_rtol)
        # factorisation and first application share one all-reduce: the forward sweep of the bottom
        # chunks only needs their own factors
        top = self._ptr(self._top)  # [partial pivots | link conductances | partial rhs], 3 n_top
        dev.call("nxfx_pc_setup_apply_begin", b, top)
        self._dist.all_reduce(self._top, group=self._group)
        dev.call("nxfx_pc_setup_apply_end", b, x, top)
        self.history = []
        applied = 0
        while refine_steps > 0 or final_residual:
            self._residual(b, x, 0)
            vals = self._nrm[:2].cpu().numpy()  # synchronises; identical on all ranks
            self.rhs_norm = float(np.sqrt(vals[1]))
            rel = float(np.sqrt(vals[0]) / self.rhs_norm) if self.rhs_norm > 0 else 0.0
            self.history.append(rel)
            if applied >= refine_steps or (refine_rtol > 0 and rel <= refine_rtol) or not np.isfinite(rel):
                break
            self._apply(self._r.c_ptr, x, 1)
            applied += 1
            if applied >= refine_steps and not final_residual:
                break
        if not self.history:
            dev.sync()
        self.solver.x.mark_device_modified()
        self.corrections = applied
        return self.history

    def _solve_peer(self, b, x, refine_steps, final_residual, refine_rtol):
        """The single-GPU launch sequence (fused factor+solve tree kernel, back-substitution, residual)
        with the exchanges done inside the kernels over NVLink: one library call, one host
        synchronisation, no NCCL."""
        from . import _lib

        C = self._C
        key = (int(refine_steps), bool(final_residual), float(refine_rtol))
        cache = getattr(self, "_peer_opts", None)
        if cache is None or cache[0] != key:  # the option struct is built once: the step loop stays off the Python heap
            opts = self.solver.solve_options()
            opts.ksp_type, opts.pc_type = _lib.KSP_PREONLY, _lib.PC_NETWORK_SCHUR
            opts.refine_steps, opts.final_residual, opts.refine_rtol = key[0], int(key[1]), key[2]
            opts.error_if_not_converged = 0
            self._peer_opts = cache = (key, opts, _lib.SolveInfo())
        _, opts, info = cache
        self.solver.A._materialise_zero()
        self.solver.A.bind()
        self.dev.call("nxfx_solve", b, x, C.byref(opts), C.byref(info))
        self.solver.x.mark_device_modified()
        self.rhs_norm = float(info.rhs_norm)
        den = self.rhs_norm if self.rhs_norm > 0 else 1.0
        self.history = [info.history[i] / den for i in range(info.history_len)]
        self.corrections = int(info.iterations) - 1
        return self.history

    def _residual(self, b, x, k: int) -> None:
        """r = b - A x with consistent shared rows and the global [||r||^2, ||b||^2] in slot k:
        one all-reduce of n_shared + 2 doubles."""
        dev = self.dev
        sh = self._ptr(self._sh)
        dev.call("nxfx_residual_partial", b, x, self._r.c_ptr, sh)
        self._dist.all_reduce(self._sh, group=self._group)
        dev.call("nxfx_residual_finish", sh, self._r.c_ptr, self._ptr(self._nrm, 2 * k))

    # ---- results -------------------------------------------------------------------------------
    def local_solution(self) -> np.ndarray:
        return self.solver.x.array_r

    def edge_values(self):
        """(global edge ids, flux[E_loc, N+1], pressure[E_loc, N]) of the local edges and
        (global multiplier ids, values) of the multipliers this rank owns."""
        x = self.local_solution()
        N = self.mesh.cells_per_edge
        E = self.part.global_edges.size
        fb = self.mesh.edge_slot.astype(np.int64) * (N + 1)
        q = x[fb[:, None] + np.arange(N + 1)[None, :]]
        nq = E * (N + 1)
        p = x[nq: nq + E * N].reshape(E, N)
        lam = x[nq + E * N:]
        own = self.part.lam_weight > 0
        return self.part.global_edges, q, p, self.part.global_bif[own], lam[own]
