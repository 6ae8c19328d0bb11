"""Murray-law arterial tree, 40 cells per vessel, nest matrix (BASELINE config 3; reference
demos/demo_arterial_tree.py) -- plus the radius-dependent Poiseuille resistance R = 8 mu / (pi r^4)."""
import networkx as nx
import numpy as np

from networks_fenicsx_b200 import HydraulicNetworkAssembler, NetworkMesh, Solver
from networks_fenicsx_b200.network_generation import make_arterial_tree
from networks_fenicsx_b200.post_processing import extract_global_flux


def p_bc_expr(x):
    return x[1]


n = 5
G = make_arterial_tree(N=n, direction=np.array([0.1, 1, 0]))
network_mesh = NetworkMesh(G, N=40, color_strategy=nx.coloring.strategy_largest_first)
assembler = HydraulicNetworkAssembler(network_mesh, flux_degree=1, pressure_degree=0)
for label, R in (("R = 1", None),
                 ("R = 8 mu / (pi r^4)", np.repeat(8.0 / (np.pi * np.array([G.edges[e]["radius"] for e in G.edges()]) ** 4), 40))):
    assembler.compute_forms(p_bc_ex=p_bc_expr, R=R)
    solver = Solver(assembler, kind="nest")
    solver.assemble()
    sol = solver.solve()
    global_flux = extract_global_flux(network_mesh, sol)
    print(f"{label}: {assembler.num_dofs} dofs, inlet flux {float(global_flux.x.array[0]):.6e}, "
          f"|residual| {solver.ksp.getResidualNorm():.2e}")
