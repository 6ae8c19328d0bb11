"""Phase timing of the fused tree kernel in the partitioned (multi-GPU) solve (development helper; needs the
-DNXFX_TREE_STAMPS build):

    NXFX_LIB=.../libnxfx_b200_stamps.so torchrun --nproc-per-node 2 scripts/dist_stamps.py [generations]
"""
import ctypes as C
import os
import pathlib
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from networks_fenicsx_b200 import _lib  # noqa: E402

_lib.LIB_PATH = pathlib.Path(os.environ["NXFX_LIB"]).resolve()
import networks_fenicsx_b200 as nxfx  # noqa: E402
from networks_fenicsx_b200.distributed import DistributedSolver  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 21
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
G = nxfx.network_generation.make_tree(n, n, n, as_arrays=True)
ds = DistributedSolver(G, 1, lambda x: x[1], device=lr)
dev = ds.dev
names = {0: "start", 1: "tables", 2: "gdc.wait", 3: "staged", 4: "ticket", 5: "fold", 12: "pre-exchange", 13: "exchanged",
         6: "sweep_up", 7: "store d/gd", 8: "published", 9: "flag", 10: "solve_down", 11: "flag raised"}
acc, host = [], []
for it in range(12):
    t0 = time.perf_counter()
    ds.assemble()
    ds.solve()
    host.append(time.perf_counter() - t0)
    st = (C.c_ulonglong * 32)()
    dev.lib.nxfx_debug_tree_stamps(dev.handle, st)
    acc.append(np.array(list(st), dtype=np.int64))
a = np.array(acc[4:])
t0 = np.minimum(a[:, 0], a[:, 16])
for r in range(dist.get_world_size()):
    if r == dist.get_rank():
        print(f"rank {r}: exchange {ds.exchange}, n_shared {ds.part.n_top}, host step (assemble+solve wall) median {np.median(host[4:]) * 1e6:.1f} us", flush=True)
        for blk, off in (("block 0", 0), ("top block", 16)):
            print(" ", blk)
            for k in (0, 1, 2, 3, 4, 5, 12, 13, 6, 7, 8, 9, 10, 11):
                v = a[:, off + k]
                if (v == 0).all():
                    continue
                print(f"     {names[k]:12s} +{np.median(v - t0) / 1e3:7.2f} us", flush=True)
    dist.barrier()
dist.destroy_process_group()
