"""Mesh-refinement study on a two-generation tree (BASELINE config 2; reference demos/demo_tree.py):
min / max / mean of the global flux for N = 2 .. 1024 cells per edge."""
import numpy as np

from networks_fenicsx_b200 import HydraulicNetworkAssembler, NetworkMesh, Solver, network_generation
from networks_fenicsx_b200.post_processing import extract_global_flux


def p_bc(x):
    return x[1]


G = network_generation.make_tree(n=2, H=1, W=1)
N = 1
for i in range(10):
    N *= 2
    network_mesh = NetworkMesh(G, N=N)
    assembler = HydraulicNetworkAssembler(network_mesh)
    assembler.compute_forms(p_bc_ex=p_bc)
    solver = Solver(assembler, petsc_options={"ksp_type": "preonly", "pc_type": "lu",
                                              "pc_factor_mat_solver_type": "mumps"}, kind="mpi")
    solver.assemble()
    sol = solver.solve()
    global_flux = extract_global_flux(network_mesh, sol)
    q = global_flux.x.array.reshape(-1, 2)
    # mean flux = int q dx / int 1 dx with the DG1 representation (cell average x cell length)
    xg = network_mesh.mesh.geometry.x
    cells = network_mesh.mesh.topology.connectivity(1, 0).array.reshape(-1, 2)
    h = np.linalg.norm(xg[cells[:, 1]] - xg[cells[:, 0]], axis=1)
    mean_q = float(np.sum(q.mean(axis=1) * h) / np.sum(h))
    print(f"N = {N:5d}  min {q.min():.12f}  max {q.max():.12f}  mean {mean_q:.12f}")
