"""Timing of the stages for trees with n generations (BASELINE config 4; reference demos/demo_perf.py:
same stages and timer names, read with ``common.timing``).  The reference stops solving at n >= 20;
here the 20-generation tree (3.67 M DOFs) is part of the default list."""
import sys

from networks_fenicsx_b200 import HydraulicNetworkAssembler, NetworkMesh, Solver, network_generation
from networks_fenicsx_b200.common import timing
from networks_fenicsx_b200.post_processing import extract_global_flux


def p_bc(x):
    return x[1]


ns = [int(a) for a in sys.argv[1:]] or [3, 6, 12, 16, 20]
tracked = ["nxfx:NetworkMesh:build_mesh", "nxfx:NetworkMesh:build_network_submeshes",
           "nxfx:NetworkMesh:create_lm_submesh", "nxfx:HydraulicNetworkAssembler:__init__",
           "nxfx:HydraulicNetworkAssembler:compute_forms", "nxfx:HydraulicNetworkAssembler:assemble",
           "nxfx:Solver:solve"]
previous = {k: 0.0 for k in tracked}
print(f"{'n':>3} {'segments':>9} " + " ".join(f"{k.split(':')[-1]:>24}" for k in tracked))
for n in ns:
    G = network_generation.make_tree(n=n, H=n, W=n, as_arrays=n > 14)
    network_mesh = NetworkMesh(G, N=1, color_strategy="smallest_last")
    assembler = HydraulicNetworkAssembler(network_mesh, flux_degree=1, pressure_degree=0)
    assembler.compute_forms(p_bc_ex=p_bc)
    solver = Solver(assembler)
    solver.assemble()
    sol = solver.solve()
    if n <= 16:
        extract_global_flux(network_mesh, sol)
    row = []
    for k in tracked:
        total = timing(k)[1].total_seconds()
        row.append(total - previous[k])
        previous[k] = total
    print(f"{n:>3} {2**n - 1:>9} " + " ".join(f"{v:>24.6f}" for v in row))
    del solver, assembler, network_mesh
