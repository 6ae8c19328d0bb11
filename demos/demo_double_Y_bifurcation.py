"""Wider Y bifurcation with p_bc = x (BASELINE config 1; reference demos/demo_double_Y_bifurcation.py)."""
from pathlib import Path

from networks_fenicsx_b200 import HydraulicNetworkAssembler, NetworkMesh, Solver, fem, network_generation
from networks_fenicsx_b200.post_processing import export_functions, extract_global_flux

G = network_generation.make_tree(2, 3.1, 7.3)
network_mesh = NetworkMesh(G, N=5)
x = fem.SpatialCoordinate(network_mesh.mesh)

assembler = HydraulicNetworkAssembler(network_mesh)
assembler.compute_forms(p_bc_ex=x[0])

solver = Solver(assembler)
solver.assemble()
sol = solver.solve()

global_flux = extract_global_flux(network_mesh, sol)
export_functions(sol, outpath=Path(__file__).parent / "results_double_Y_bifurcation")
print("double Y: fluxes", [float(f.x.array[0]) for f in sol[:-2]])
