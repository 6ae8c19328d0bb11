"""Multi-GPU check of the single-network partition (run under torchrun on >= 2 GPUs):

    torchrun --nproc-per-node 2 tests/dist_check.py [generations] [N] [chunk] [workload] [exchange]

workload ``tree``: ``make_tree(n, n, n)``, R = 1, f = 0 (demo_perf.py); ``arterial``:
``make_arterial_tree(n)`` with the radius-dependent resistance R_e = 8 mu / (pi r_e^4) from the graph's
``radius`` attribute and a source term f != 0 (BASELINE configs[3], demo_arterial_tree.py:16-27).
exchange ``peer`` (in-kernel NVLink exchange, N == 1), ``nccl`` (split phases around all-reduces) or
``auto``.  ``NXFX_DIST_ONE_GPU=1``: all ranks share cuda:0 and reduce through gloo (single-GPU boxes).  Every rank solves its part; rank 0 gathers the solution and compares it with the CPU
oracle's direct solve of the WHOLE network (bar: 1e-8 relative L2, BASELINE north_star)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import networks_fenicsx_b200 as nxfx  # noqa: E402
from networks_fenicsx_b200.distributed import DistributedSolver  # noqa: E402
from networks_fenicsx_b200.mesh import _edge_colors  # noqa: E402


def build_workload(workload, n, N):
    """(graph, p_bc, R per graph edge or None, f per cell or None) -- global, identical on all ranks."""
    if workload == "tree":
        G = nxfx.network_generation.make_tree(n, n, n, as_arrays=True)
        return G, (lambda x: x[1] + 0.1 * x[0]), None, None
    G = nxfx.network_generation.make_arterial_tree(N=n, direction=np.array([0.1, 1.0, 0.0]), as_arrays=True)
    radius = np.asarray(G.edge_attrs["radius"], dtype=np.float64)
    R = 8.0 * 1.0 / (np.pi * radius**4)  # Poiseuille resistance per unit length, mu = 1
    E = G.number_of_edges()
    f = 1e-3 * np.sin(np.arange(E * N, dtype=np.float64))  # deterministic, != 0
    return G, (lambda x: x[1]), R, f


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
    workload = sys.argv[4] if len(sys.argv) > 4 else "tree"
    exchange = sys.argv[5] if len(sys.argv) > 5 else "auto"
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("NXFX_DIST_ONE_GPU") == "1":
        # every rank on cuda:0 (time-sliced contexts), collectives through gloo: the partition, the split-phase
        # kernels and the host-driven all-reduces of the partitioned solve on a single-GPU box (NCCL refuses two
        # ranks on one device; the in-kernel peer exchange would spin against a kernel that is not running)
        local_rank, exchange = 0, "nccl"
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    G, p_bc, R, f = build_workload(workload, n, N)
    ds = DistributedSolver(G, N, p_bc, R=R, f=f, device=local_rank, chunk_nodes=chunk, exchange=exchange)
    ds.assemble()
    hist = ds.solve(refine_steps=1, final_residual=True)
    ok = True
    if workload == "tree":
        assert ds.corrections == 0, f"the distributed direct solve needed {ds.corrections} correction(s) on a tree: {hist}"
    # a second step on the same objects (epochs / parities of the peer exchange advance)
    ds.assemble()
    hist2 = ds.solve(refine_steps=1, final_residual=False)
    ge, q, p, gl, lam = ds.edge_values()
    gathered = [None] * world
    dist.gather_object((ge, q, p, gl, lam), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        from oracle import reference_port as rp

        colors = _edge_colors(G, "smallest_last", np.asarray(G.edges, dtype=np.int64))
        net = rp.OracleNetwork(G.pos, G.edges, colors, N)
        Rc = 1.0 if R is None else np.repeat(R, N)
        A, b = net.assemble(net.eval_pbc(p_bc), R=Rc, f=0.0 if f is None else f)
        x_ref = net.solve(A, b)
        x = np.full(net.n_dofs, np.nan)
        for ge_, q_, p_, gl_, lam_ in gathered:
            x[net.fb[ge_][:, None] + np.arange(N + 1)[None, :]] = q_
            x[net.pb[ge_][:, None] + np.arange(N)[None, :]] = p_
            x[net.loff + gl_] = lam_
        assert not np.isnan(x).any(), "some dofs were not owned by any rank"
        err = np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref)
        res = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
        # the reported (exchanged) norm must be the true one
        ok = err < 1e-8 and 0.2 * res <= hist[-1] <= 5 * res + 1e-17 and abs(hist2[0] - hist[0]) <= 1e-3 * hist[0] + 1e-17
        print(f"dist_check world={world} workload={workload} n={n} N={N} exchange={ds.exchange}: dofs={net.n_dofs} "
              f"(allreduce {ds.n_dofs_global}) n_top={ds.part.n_top} corrections={ds.corrections} "
              f"rel L2 error vs direct solve {err:.2e}, true residual {res:.2e}, reported residuals {hist} / {hist2} "
              f"-> {'OK' if ok else 'FAIL'}", flush=True)
        assert ds.n_dofs_global == net.n_dofs
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
