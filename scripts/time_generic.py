"""Time the exact condensation of the higher-order path with CUDA events (development helper).

    python scripts/time_generic.py [generations] [cells_per_edge]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import networks_fenicsx_b200 as nxfx  # noqa: E402
from networks_fenicsx_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for fd, pd in ((2, 1), (3, 2), (2, 0)):
    G = nxfx.network_generation.make_tree(n, n, n, as_arrays=True)
    nm = nxfx.NetworkMesh(G, N=N, color_strategy="smallest_last")
    asm = nxfx.HydraulicNetworkAssembler(nm, flux_degree=fd, pressure_degree=pd)
    asm.compute_forms(p_bc_ex=lambda x: x[1])
    solver = nxfx.Solver(asm)
    solver.assemble()
    solver.solve()
    dev = nm.device

    def timeit(fn, reps=10):
        fn(); dev.sync(); dev.timer_start()
        for _ in range(reps):
            fn()
        return dev.timer_stop() / reps * 1e3

    y = solver.b.duplicate()
    t_set = timeit(lambda: dev.call("nxfx_pc_setup"))
    t_pc = timeit(lambda: dev.call("nxfx_pc_apply", solver.b.d.c_ptr, y.d.c_ptr))
    t_spmv = timeit(lambda: dev.call("nxfx_spmv", solver.x.d.c_ptr, y.d.c_ptr))
    opts = solver.solve_options(); info = _lib.SolveInfo()
    t_solve = timeit(lambda: (dev.call("nxfx_pc_setup"), dev.call("nxfx_solve", solver.b.d.c_ptr, solver.x.d.c_ptr, C.byref(opts), C.byref(info))))
    print(f"P{fd}/P{pd} make_tree({n}) N={N}: {asm.num_dofs} dofs, nnz {solver.A.nnz}: factor {t_set:7.1f} us  apply {t_pc:7.1f} us  "
          f"spmv {t_spmv:6.1f} us  factor+solve {t_solve:7.1f} us = {asm.num_dofs / t_solve * 1e6:.3e} DOF/s  its {info.iterations}  "
          f"res {info.residual_norm / info.rhs_norm:.1e}")
