"""Convenience functions for post-processing (drop-in for ``networks_fenicsx.post_processing``,
post_processing.py:19-97)."""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from .fem import Function, FunctionSpace
from .mesh import NetworkMesh

__all__ = ["extract_global_flux", "export_functions", "export_submeshes"]


def extract_global_flux(graph_mesh: NetworkMesh, functions: list[Function]) -> Function:
    """Gather the per-colour fluxes into one DG function on the parent mesh
    (post_processing.py:19-52): cell ``c`` holds the flux at its first and second vertex.

    Args:
        graph_mesh: The network mesh
        functions: The list of functions ``[flux_1, ..., flux_M, pressure, lm]``
    """
    flux_functions = functions[:-2]
    q_degree = flux_functions[0].function_space.element.basix_element.degree
    nc = graph_mesh.mesh.topology.index_map(1).size_local
    V = FunctionSpace(
        graph_mesh.mesh, q_degree, True, (q_degree + 1) * nc, 0,
        lambda: np.arange((q_degree + 1) * nc, dtype=np.int32).reshape(nc, q_degree + 1), "global_flux",
    )
    dev = graph_mesh.device
    if q_degree != 1:
        # higher order: host gather through the cell -> dof tables of the per-colour spaces
        global_q = Function(V, name="Global_Flux")
        out = global_q.x.array.reshape(nc, q_degree + 1)
        for i, flux in enumerate(flux_functions):
            flux.name = f"Flux_{i}"
            cells = graph_mesh.entity_maps[i].sub_topology_to_topology(
                np.arange(flux.function_space.dofmap.list.shape[0], dtype=np.int32))
            out[cells] = flux.x.array[flux.function_space.dofmap.list]
        return global_q
    global_q = Function(V, name="Global_Flux", array=dev.pinned(V.num_dofs))
    nq = sum(f.x.array.size for f in flux_functions)
    xq = dev.empty(nq)
    off = 0
    for i, flux in enumerate(flux_functions):
        flux.name = f"Flux_{i}"
        n = flux.x.array.size
        dev.call("nxfx_memcpy_h2d", C.c_void_p(xq.ptr + 8 * off), C.c_void_p(flux.x.array.ctypes.data), C.c_size_t(8 * n))
        off += n
    out = dev.empty(V.num_dofs)
    dev.call("nxfx_global_flux", xq.c_ptr, out.c_ptr)
    out.download(global_q.x.array)
    return global_q


def export_functions(functions: list[Function], outpath: Path | str):
    """Write the solution functions as ``.npy`` arrays (``flux_i``, ``pressure``, ``lm``) --
    the reference writes ADIOS2 VTX files (post_processing.py:55-78); ADIOS2 is not part of this
    build, see DESIGN.md "out of scope"."""
    export_path = Path(outpath)
    export_path.mkdir(parents=True, exist_ok=True)
    for i, q in enumerate(functions[:-2]):
        np.save(export_path / f"flux_{i}.npy", q.x.array)
    np.save(export_path / "pressure.npy", functions[-2].x.array)
    np.save(export_path / "lm.npy", functions[-1].x.array)


def export_submeshes(network_mesh: NetworkMesh, outpath: str | Path):
    """Write each colour submesh (vertices, cells, facet markers) as ``.npz``
    (reference: XDMF, post_processing.py:81-97)."""
    outpath = Path(outpath)
    outpath.mkdir(parents=True, exist_ok=True)
    for i in range(network_mesh.num_edge_colors):
        sm = network_mesh.submeshes[i]
        tags = network_mesh.submesh_facet_markers[i]
        np.savez(
            outpath / f"submesh_{i}.npz", x=sm.geometry.x,
            cells=sm.topology.connectivity(1, 0).array.reshape(-1, 2),
            facet_indices=tags.indices, facet_values=tags.values,
        )
