"""Phase timing of the fused tree kernel (development helper; needs the -DNXFX_TREE_STAMPS build).

    NXFX_LIB=networks_fenicsx_b200/csrc/libnxfx_b200_stamps.so python scripts/tree_stamps.py [generations]
"""
import ctypes as C
import os
import pathlib
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from networks_fenicsx_b200 import _lib  # noqa: E402

_lib.LIB_PATH = pathlib.Path(os.environ["NXFX_LIB"]).resolve()
import networks_fenicsx_b200 as nxfx  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
G = nxfx.network_generation.make_tree(n, n, n, as_arrays=True)
nm = nxfx.NetworkMesh(G, N=1, color_strategy="smallest_last")
asm = nxfx.HydraulicNetworkAssembler(nm)
asm.compute_forms(p_bc_ex=lambda x: x[1])
solver = nxfx.Solver(asm)
dev = nm.device
lib = dev.lib
attrs = (C.c_int * 8)()
lib.nxfx_debug_device_attrs(dev.handle, attrs)
print("attrs [L2, maxPersistL2, maxWindow, smemPerSM, SMs, hostPtrReg, cluster, pools]:", list(attrs))
opts = solver.solve_options()
info = _lib.SolveInfo()
names = {0: "start", 1: "tables", 2: "gdc.wait", 3: "staged", 4: "ticket", 5: "fold", 6: "sweep_up", 7: "store d/gd",
         8: "published", 9: "flag", 10: "solve_down", 11: "flag raised"}
acc = []
for it in range(8):
    solver.assemble()
    dev.call("nxfx_solve", solver.b.device_ptr(), solver.x.device_ptr_overwrite(), C.byref(opts), C.byref(info))
    st = (C.c_ulonglong * 32)()
    lib.nxfx_debug_tree_stamps(dev.handle, st)
    acc.append(np.array(list(st), dtype=np.int64))
a = np.array(acc[2:])
t0 = np.minimum(a[:, 0], a[:, 16])
for blk, off in (("block 0", 0), ("top block", 16)):
    print(blk)
    for k in sorted(names):
        v = a[:, off + k]
        if (v == 0).all():
            continue
        print(f"   {names[k]:12s} +{np.median(v - t0) / 1e3:7.2f} us")
print("residual", info.residual_norm / info.rhs_norm)
