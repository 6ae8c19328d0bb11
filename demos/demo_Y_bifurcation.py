"""Single Y bifurcation -- BASELINE config 0, the workload of the reference's
demos/demo_Y_bifurcation.py (make_tree(2, 1, 3), 4 cells per edge, one flux space per edge,
boundary pressure p = y)."""
import pathlib

import networks_fenicsx_b200 as nxfx


def main(cells_per_edge: int = 4) -> list:
    graph = nxfx.network_generation.make_tree(2, 1, 3)
    net = nxfx.NetworkMesh(graph, N=cells_per_edge)
    y = nxfx.fem.SpatialCoordinate(net.mesh)[1]

    problem = nxfx.HydraulicNetworkAssembler(net)
    problem.compute_forms(p_bc_ex=y)
    ksp = nxfx.Solver(problem)
    ksp.assemble()
    fields = ksp.solve()

    nxfx.post_processing.extract_global_flux(net, fields)
    nxfx.post_processing.export_functions(fields, pathlib.Path(__file__).parent / "results_Y_bifurcation")
    print("Y bifurcation: fluxes", [float(f.x.array[0]) for f in fields[:-2]], "multiplier", fields[-1].x.array)
    return fields


if __name__ == "__main__":
    main()
