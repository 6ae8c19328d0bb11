"""A few assemble+solve steps of the headline workload (short target for ncu captures)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import networks_fenicsx_b200 as nxfx  # noqa: E402
from networks_fenicsx_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
G = nxfx.network_generation.make_tree(n, n, n, as_arrays=True)
nm = nxfx.NetworkMesh(G, N=1, color_strategy="smallest_last")
asm = nxfx.HydraulicNetworkAssembler(nm)
asm.compute_forms(p_bc_ex=lambda x: x[1])
solver = nxfx.Solver(asm)
opts, info = solver.solve_options(), _lib.SolveInfo()
for _ in range(steps):
    solver.assemble()
    nm.device.call("nxfx_solve", solver.b.device_ptr(), solver.x.device_ptr_overwrite(), C.byref(opts), C.byref(info))
print("residual", info.residual_norm / info.rhs_norm)
