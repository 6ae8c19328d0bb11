"""A few assemble+solve steps of a higher-order workload (short target for ncu captures of the condensation)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import networks_fenicsx_b200 as nxfx  # noqa: E402
from networks_fenicsx_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fd, pd = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (2, 1)
G = nxfx.network_generation.make_tree(n, n, n, as_arrays=True)
nm = nxfx.NetworkMesh(G, N=4, color_strategy="smallest_last")
asm = nxfx.HydraulicNetworkAssembler(nm, flux_degree=fd, pressure_degree=pd)
asm.compute_forms(p_bc_ex=lambda x: x[1])
solver = nxfx.Solver(asm)
opts, info = solver.solve_options(), _lib.SolveInfo()
for _ in range(steps):
    solver.assemble()
    nm.device.call("nxfx_solve", solver.b.device_ptr(), solver.x.device_ptr_overwrite(), C.byref(opts), C.byref(info))
print("dofs", asm.num_dofs, "residual", info.residual_norm / info.rhs_norm, "iterations", info.iterations)
