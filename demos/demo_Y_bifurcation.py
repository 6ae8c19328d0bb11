"""Single Y bifurcation (BASELINE config 0; mirrors the reference's demos/demo_Y_bifurcation.py)."""
from pathlib import Path

from networks_fenicsx_b200 import HydraulicNetworkAssembler, NetworkMesh, Solver, fem, network_generation
from networks_fenicsx_b200.post_processing import export_functions, extract_global_flux

outdir = Path(__file__).parent / "results_Y_bifurcation"

G = network_generation.make_tree(2, 1, 3)
network_mesh = NetworkMesh(G, N=4)

x = fem.SpatialCoordinate(network_mesh.mesh)
assembler = HydraulicNetworkAssembler(network_mesh)
assembler.compute_forms(p_bc_ex=x[1])

solver = Solver(assembler)
solver.assemble()
sol = solver.solve()

global_flux = extract_global_flux(network_mesh, sol)
export_functions(functions=sol, outpath=outdir)
print("Y bifurcation: fluxes", [float(f.x.array[0]) for f in sol[:-2]], "multiplier", sol[-1].x.array)
