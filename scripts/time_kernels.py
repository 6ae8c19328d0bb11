"""Time the hot kernels of the headline workload with CUDA events (development helper).

    NXFX_LIB=/path/to/variant.so python scripts/time_kernels.py [generations]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from networks_fenicsx_b200 import _lib  # noqa: E402

if os.environ.get("NXFX_LIB"):
    import pathlib

    _lib.LIB_PATH = pathlib.Path(os.environ["NXFX_LIB"])
import networks_fenicsx_b200 as nxfx  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
G = nxfx.network_generation.make_tree(n, n, n, as_arrays=True)
nm = nxfx.NetworkMesh(G, N=1, color_strategy="smallest_last")
asm = nxfx.HydraulicNetworkAssembler(nm)
asm.compute_forms(p_bc_ex=lambda x: x[1])
solver = nxfx.Solver(asm)
solver.assemble()
solver.solve()
dev = nm.device
nnz, nd, nv = solver.A.nnz, asm.num_dofs, nm.mesh.topology.index_map(0).size_local


def timeit(fn, reps=30):
    fn(); dev.sync(); dev.timer_start()
    for _ in range(reps):
        fn()
    return dev.timer_stop() / reps * 1e3


y = solver.b.duplicate(); bb = solver.b.duplicate()
t_asm = timeit(lambda: dev.call("nxfx_assemble", None, C.c_double(1.0), None, C.c_double(0.0), 1, 1, 0, bb.d.c_ptr))
t_spmv = timeit(lambda: dev.call("nxfx_spmv", solver.x.d.c_ptr, y.d.c_ptr))
t_set = timeit(lambda: dev.call("nxfx_pc_setup"))
t_pc = timeit(lambda: dev.call("nxfx_pc_apply", solver.b.d.c_ptr, y.d.c_ptr))
opts = solver.solve_options(); info = _lib.SolveInfo()
t_solve = timeit(lambda: (dev.call("nxfx_pc_setup"), dev.call("nxfx_solve", solver.b.d.c_ptr, solver.x.d.c_ptr, C.byref(opts), C.byref(info))), 10)
# the step as the bench runs it: a fresh matrix, factorisation fused with the first solve
t_step = timeit(lambda: (dev.call("nxfx_assemble", None, C.c_double(1.0), None, C.c_double(0.0), 1, 1, 0, bb.d.c_ptr),
                         dev.call("nxfx_solve", solver.b.d.c_ptr, solver.x.d.c_ptr, C.byref(opts), C.byref(info))), 10)
b_asm = 24 * nv + 8 * nnz + 8 * nd + 8 * nm.boundary_values.size
b_spmv = 12 * nnz + 4 * (nd + 1) + 16 * nd
print(f"{os.environ.get('NXFX_LIB', 'default'):40s} asm {t_asm:6.1f} us ({b_asm / t_asm / 1e3 / 6543.1:.3f})  spmv {t_spmv:6.1f} us "
      f"({b_spmv / t_spmv / 1e3 / 6543.1:.3f})  pc_setup {t_set:5.1f}  pc_apply {t_pc:5.1f}  setup+solve {t_solve:6.1f}  asm+solve {t_step:6.1f} us  res {info.residual_norm / info.rhs_norm:.1e}")
if os.environ.get("NXFX_ASM_SPLIT"):
    for lhs, rhs in ((1, 0), (0, 1), (1, 1)):
        t = timeit(lambda: dev.call("nxfx_assemble", None, C.c_double(1.0), None, C.c_double(0.0), lhs, rhs, 0, bb.d.c_ptr))
        print(f"   assemble lhs={lhs} rhs={rhs}: {t:6.1f} us")
    import numpy as np
    big = dev.empty(nnz)
    t = timeit(lambda: dev.call("nxfx_memset", big.c_ptr, 0, C.c_size_t(8 * nnz)))
    print(f"   memset {8*nnz/1e6:.0f} MB: {t:6.1f} us = {8*nnz/t/1e3:.0f} GB/s")
    big2 = dev.empty(nnz)
    t = timeit(lambda: dev.call("nxfx_memcpy_d2d", big2.c_ptr, big.c_ptr, C.c_size_t(8 * nnz)))
    print(f"   d2d copy {8*nnz/1e6:.0f} MB: {t:6.1f} us = {16*nnz/t/1e3:.0f} GB/s (r+w)")
