"""Multi-GPU check of the single-tree partition (run under torchrun on >= 2 GPUs):

    torchrun --nproc-per-node 2 tests/dist_check.py [generations] [N]

Every rank solves its part; rank 0 gathers the solution and compares it with the CPU oracle's
direct solve of the WHOLE network (bar: 1e-8 relative L2, BASELINE north_star)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import networks_fenicsx_b200 as nxfx  # noqa: E402
from networks_fenicsx_b200.distributed import DistributedSolver  # noqa: E402
from networks_fenicsx_b200.mesh import _greedy_edge_coloring_arrays  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    G = nxfx.network_generation.make_tree(n, n, n, as_arrays=True)
    p_bc = lambda x: x[1] + 0.1 * x[0]  # noqa: E731
    ds = DistributedSolver(G, N, p_bc, device=local_rank, chunk_nodes=chunk)
    ds.assemble()
    hist = ds.solve(refine_steps=1, final_residual=True)
    assert ds.corrections == 0, f"the distributed direct solve needed {ds.corrections} correction(s) on a tree: {hist}"
    ge, q, p, gl, lam = ds.edge_values()
    gathered = [None] * world
    dist.gather_object((ge, q, p, gl, lam), gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        from oracle import reference_port as rp

        colors = _greedy_edge_coloring_arrays(G.number_of_nodes(), G.edges)
        net = rp.OracleNetwork(G.pos, G.edges, colors, N)
        A, b = net.assemble(net.eval_pbc(p_bc))
        x_ref = net.solve(A, b)
        x = np.full(net.n_dofs, np.nan)
        for ge_, q_, p_, gl_, lam_ in gathered:
            x[net.fb[ge_][:, None] + np.arange(N + 1)[None, :]] = q_
            x[net.pb[ge_][:, None] + np.arange(N)[None, :]] = p_
            x[net.loff + gl_] = lam_
        assert not np.isnan(x).any(), "some dofs were not owned by any rank"
        err = np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref)
        res = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
        ok = err < 1e-8 and 0.2 * res <= hist[-1] <= 5 * res + 1e-17  # reported norm must match the true one
        print(f"dist_check world={world} n={n} N={N}: dofs={net.n_dofs} (allreduce {ds.n_dofs_global}) n_top={ds.part.n_top} "
              f"rel L2 error vs direct solve {err:.2e}, true residual {res:.2e}, reported residuals {hist} -> {'OK' if ok else 'FAIL'}",
              flush=True)
        assert ds.n_dofs_global == net.n_dofs
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
