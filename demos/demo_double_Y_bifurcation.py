"""Wide Y bifurcation driven by p = x -- BASELINE config 1, the workload of the reference's
demos/demo_double_Y_bifurcation.py (make_tree(2, 3.1, 7.3), 5 cells per edge).  By symmetry the
stem carries no flux and the two branches carry +-0.9204."""
import pathlib

import networks_fenicsx_b200 as nxfx


def main() -> list:
    net = nxfx.NetworkMesh(nxfx.network_generation.make_tree(2, 3.1, 7.3), N=5)
    problem = nxfx.HydraulicNetworkAssembler(net)
    problem.compute_forms(p_bc_ex=nxfx.fem.SpatialCoordinate(net.mesh)[0])
    ksp = nxfx.Solver(problem)
    ksp.assemble()
    fields = ksp.solve()
    nxfx.post_processing.export_functions(fields, pathlib.Path(__file__).parent / "results_double_Y_bifurcation")
    flux = nxfx.post_processing.extract_global_flux(net, fields)
    print("double Y: fluxes", [float(f.x.array[0]) for f in fields[:-2]], "| global flux dofs:", flux.x.array.size)
    return fields


if __name__ == "__main__":
    main()
