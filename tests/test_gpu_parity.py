"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): sparsity pattern and indexing bit-exact; matrix entries within
1e-12 relative (here: bit-exact, the kernels avoid FMA contraction); solution within 1e-8 relative
L2 of the direct solve (here: 1e-10).
"""

import networkx as nx
import numpy as np
import pytest

import networks_fenicsx_b200 as nxfx
from networks_fenicsx_b200 import network_generation as ng
from tests import helpers

pytestmark = pytest.mark.gpu

P_Y = lambda x: x[1]  # noqa: E731
P_X = lambda x: x[0]  # noqa: E731


def run_case(G, N, strategy, p_bc, R=None, f=None, kind=None, petsc_options=None):
    nm = nxfx.NetworkMesh(G, N=N, color_strategy=strategy)
    asm = nxfx.HydraulicNetworkAssembler(nm)
    asm.compute_forms(p_bc_ex=p_bc, R=R, f=f)
    solver = nxfx.Solver(asm, kind=kind, petsc_options=petsc_options)
    solver.assemble()
    sol = solver.solve()
    net = helpers.oracle_for(nm, N, G, strategy)
    A, b = net.assemble(net.eval_pbc(p_bc), R=1.0 if R is None else R, f=0.0 if f is None else f)
    return nm, asm, solver, sol, net, A, b


def check_system(solver, net, A, b, exact=True):
    rp_, ci, va = solver.A.getValuesCSR()
    assert np.array_equal(rp_, A.indptr), "row pointers differ"
    assert np.array_equal(ci, A.indices), "column indices differ"
    assert A.nnz == net.expected_nnz()
    if exact:
        assert np.array_equal(va, A.data)
        assert np.array_equal(solver.b.array_r, b)
    else:
        np.testing.assert_allclose(va, A.data, rtol=1e-12, atol=0)
        np.testing.assert_allclose(solver.b.array_r, b, rtol=1e-12, atol=1e-300)


def check_solution(sol, net, A, b, tol=1e-10):
    x_ref = net.solve(A, b)
    x = np.concatenate([f.x.array for f in sol])
    assert helpers.rel_l2(x, x_ref) < tol
    return x, x_ref


@pytest.mark.parametrize("N", [1, 2, 4, 7])
@pytest.mark.parametrize("strategy", [None, "smallest_last", "largest_first"])
def test_y_bifurcation(N, strategy):
    """configs[0]: demos/demo_Y_bifurcation.py (make_tree(2,1,3), p_bc = y)."""
    G = ng.make_tree(2, 1, 3)
    nm, asm, solver, sol, net, A, b = run_case(G, N, strategy, P_Y)
    check_system(solver, net, A, b)
    check_solution(sol, net, A, b)
    assert [f.name for f in sol] == [f"flux_color_{i}" for i in range(nm.num_edge_colors)] + ["pressure", "global_flux"]


def test_y_bifurcation_known_answer():
    """SURVEY A.6 KAT: q = (0.77485177.., 0.38742588.., 0.38742588..), lambda = -0.38742588.."""
    G = ng.make_tree(2, 1, 3)
    nm, asm, solver, sol, net, A, b = run_case(G, 4, None, P_Y)
    q = [f.x.array for f in sol[:3]]
    np.testing.assert_allclose(q[0], 0.7748517734455862, rtol=1e-12)
    np.testing.assert_allclose(q[1], 0.3874258867227931, rtol=1e-12)
    np.testing.assert_allclose(q[2], 0.3874258867227931, rtol=1e-12)
    np.testing.assert_allclose(sol[-1].x.array, [-0.3874258867227931], rtol=1e-12)
    np.testing.assert_allclose(
        sol[-2].x.array[:4], [-0.04842824, -0.14528471, -0.24214118, -0.33899765], atol=1e-8
    )


def test_double_y_demo():
    """configs[1]: demos/demo_double_Y_bifurcation.py (make_tree(2,3.1,7.3), N=5, p_bc = x)."""
    G = ng.make_tree(2, 3.1, 7.3)
    nm, asm, solver, sol, net, A, b = run_case(G, 5, None, P_X)
    check_system(solver, net, A, b)
    x, x_ref = check_solution(sol, net, A, b)
    np.testing.assert_allclose(sol[1].x.array, -0.920444352, atol=1e-8)
    np.testing.assert_allclose(sol[2].x.array, 0.920444352, atol=1e-8)
    np.testing.assert_allclose(sol[0].x.array, 0.0, atol=1e-12)


def test_two_junctions():
    G = helpers.double_junction_graph()
    nm, asm, solver, sol, net, A, b = run_case(G, 6, "largest_first", lambda x: x[1] + 0.3 * x[0])
    check_system(solver, net, A, b)
    check_solution(sol, net, A, b)


@pytest.mark.parametrize("n,N", [(10, 1), (10, 4), (10, 16), (7, 64)])
def test_tree(n, N):
    """configs[2]: binary tree, 10 generations, refined edges (demo_tree / demo_perf protocol)."""
    G = ng.make_tree(n, n, n)
    nm, asm, solver, sol, net, A, b = run_case(G, N, "smallest_last", P_Y)
    check_system(solver, net, A, b)
    x, x_ref = check_solution(sol, net, A, b)
    # on a tree the Schur preconditioner is an exact solve: no refinement correction may be needed
    # (a correction being applied here means the preconditioner is broken and merely being repaired)
    assert solver.ksp.getIterationNumber() == 1 and solver.info.residual_norm <= 1e-13 * solver.info.rhs_norm
    # closed form: resistor network with boundary pressures -p_bc
    q_edge, lam = net.resistor_network_solution(net.eval_pbc(P_Y))
    np.testing.assert_allclose(sol[-1].x.array, lam, rtol=1e-9, atol=1e-12)
    qs = np.concatenate([f.x.array for f in sol[:-2]])
    q_first = qs[net.fb]
    np.testing.assert_allclose(q_first, q_edge, rtol=1e-8, atol=1e-11)


def test_demo_tree_refinement():
    """demos/demo_tree.py: make_tree(2,1,1) refined to N=1024, kind='mpi'."""
    G = ng.make_tree(2, 1, 1)
    for N in (2, 64, 1024):
        nm, asm, solver, sol, net, A, b = run_case(G, N, None, P_Y, kind="mpi")
        check_system(solver, net, A, b)
        check_solution(sol, net, A, b)
        gq = nxfx.post_processing.extract_global_flux(nm, sol)
        x = np.concatenate([f.x.array for f in sol])
        assert np.array_equal(gq.x.array, net.global_flux(x))


def test_arterial_tree_nest():
    """configs[3]: demos/demo_arterial_tree.py (N=5 generations, 40 cells/edge, largest_first,
    kind='nest'), plus radius-dependent resistance R_e = 8 mu / (pi r_e^4)."""
    G = ng.make_arterial_tree(N=5, direction=np.array([0.1, 1, 0]))
    nm, asm, solver, sol, net, A, b = run_case(G, 40, nx.coloring.strategy_largest_first, P_Y, kind="nest")
    assert solver.A.getType() == "nest"
    check_system(solver, net, A, b)
    check_solution(sol, net, A, b)
    assert net.n_dofs == 2526 and A.nnz == 8891
    blk = solver.A.getNestSubMatrix(0, 0)
    assert blk.shape == (asm.block_sizes[0],) * 2
    radius = np.array([G.edges[e]["radius"] for e in G.edges()])
    R_edge = 8.0 * 1.0 / (np.pi * radius**4)
    nm, asm, solver, sol, net, A, b = run_case(G, 40, "largest_first", P_Y, R=np.repeat(R_edge, 40))
    check_system(solver, net, A, b)
    check_solution(sol, net, A, b)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_tree_with_source(seed):
    G = helpers.random_tree(200, seed)
    rng = np.random.default_rng(seed)
    N = 3
    nc = N * G.number_of_edges()
    R = rng.uniform(0.5, 2.0, nc)
    f = rng.normal(size=nc)
    nm, asm, solver, sol, net, A, b = run_case(G, N, "smallest_last", lambda x: x[0] - 2 * x[2], R=R, f=f)
    check_system(solver, net, A, b)
    check_solution(sol, net, A, b)


def test_cyclic_graph_solve():
    """tests/test_edge_info.py graph (cycles, degree-4 node): the Schur preconditioner is a
    spanning-tree approximation there; FGMRES must still reach the direct solution."""
    G = helpers.edge_info_graph()
    nm, asm, solver, sol, net, A, b = run_case(G, 10, None, lambda x: x[2])
    check_system(solver, net, A, b)
    check_solution(sol, net, A, b, tol=1e-9)
    assert not solver._schedule.is_forest


def test_orientation_and_counts():
    """tests/test_orientation.py + tests/test_make_tree.py of the reference on device vertices."""
    for order, ordered in (("in", lambda _: True), ("reverse", lambda _: False), ("alt", lambda k: k % 2)):
        for N in (1, 4, 8):
            G = helpers.linear_graph(30, ordered=ordered)
            nm = nxfx.NetworkMesh(G, N=N)
            val = nm.oriented_tangent_integral((1, 0))
            expected = {"in": 1.0, "reverse": -1.0, "alt": (29 % 2) * -1 / 29}[order]
            assert np.isclose(val, expected)
            net = helpers.oracle_for(nm, N)
            assert np.array_equal(nm.mesh.geometry.x, net.x3), "device vertices differ from mesh.py:290"


def test_spmv_and_krylov_options():
    G = ng.make_tree(8, 8, 8)
    nm, asm, solver, sol, net, A, b = run_case(G, 3, "smallest_last", P_Y)
    rng = np.random.default_rng(0)
    xv = solver.x.duplicate()
    yv = solver.x.duplicate()
    xh = rng.normal(size=net.n_dofs)
    xv.array[:] = xh
    solver.A.mult(xv, yv)
    assert np.array_equal(yv.array_r, A @ xh), "SpMV is not bit-identical to a sequential CSR product"
    x_ref = net.solve(A, b)
    for opts in (
        {"ksp_type": "gmres", "pc_type": "lu", "ksp_rtol": 1e-12},
        {"ksp_type": "gmres", "pc_type": "jacobi", "ksp_rtol": 1e-12, "ksp_gmres_restart": 100, "ksp_max_it": 3000},
    ):
        s2 = nxfx.Solver(asm, petsc_options=opts)
        s2.assemble()
        sol2 = s2.solve()
        x = np.concatenate([f.x.array for f in sol2])
        assert helpers.rel_l2(x, x_ref) < 1e-8, (opts, s2.ksp.getIterationNumber())


def test_accumulate_semantics_and_errors():
    G = ng.make_tree(4, 1, 1)
    nm = nxfx.NetworkMesh(G, N=2, color_strategy="smallest_last")
    asm = nxfx.HydraulicNetworkAssembler(nm)
    with pytest.raises(RuntimeError):
        asm.assemble()
    asm.compute_forms(p_bc_ex=P_Y)
    A, b = asm.assemble()
    va1 = A.getValuesCSR()[2].copy()
    b1 = b.array_r.copy()
    asm.assemble(A, b)  # ADD_VALUES into non-zeroed targets (assembly.py:355,362)
    assert np.array_equal(A.getValuesCSR()[2], 2 * va1)
    assert np.array_equal(b.array_r, 2 * b1)
    A.zeroEntries()
    b.zeroEntries()
    asm.assemble(A, b)
    assert np.array_equal(A.getValuesCSR()[2], va1) and np.array_equal(b.array_r, b1)
    solver = nxfx.Solver(asm, petsc_options={"ksp_type": "gmres", "pc_type": "none", "ksp_max_it": 2,
                                             "ksp_rtol": 1e-14, "ksp_error_if_not_converged": True})
    solver.assemble()
    with pytest.raises(RuntimeError, match="did not converge"):
        solver.solve()


def test_full_size_properties():
    """BASELINE headline size (make_tree(20,20,20), N=1, 3,670,012 DOFs): size-independent
    properties instead of an element-wise oracle comparison -- nnz formula, pattern symmetry of
    the coupling blocks, true residual, Kirchhoff at every bifurcation, q constant per edge,
    the DG0 pressure recurrence, and the resistor-network closed form."""
    import scipy.sparse as sp

    n = 20
    G = ng.make_tree(n, n, n, as_arrays=True)
    nm = nxfx.NetworkMesh(G, N=1, color_strategy="smallest_last")
    asm = nxfx.HydraulicNetworkAssembler(nm)
    asm.compute_forms(p_bc_ex=P_Y)
    solver = nxfx.Solver(asm, petsc_options={"ksp_type": "preonly", "pc_type": "lu", "nxfx_final_residual": True,
                                             "ksp_error_if_not_converged": True})
    solver.assemble()
    sol = solver.solve()
    E = G.number_of_edges()
    n_bif = nm.bifurcation_values.size
    assert asm.num_dofs == 3670012 and solver.A.nnz == 14680044 == E * 8 + 4 * (2 * E - nm.boundary_values.size)
    assert solver.info.residual_norm <= 1e-13 * solver.info.rhs_norm
    assert solver.ksp.getIterationNumber() == 1, "the direct solve needed a refinement correction on a tree"
    rp_, ci, va = solver.A.getValuesCSR()
    A = sp.csr_matrix((va, ci, rp_), shape=(asm.num_dofs,) * 2)
    x = np.concatenate([f.x.array for f in sol])
    b = solver.b.array_r
    assert np.linalg.norm(A @ x - b) <= 1e-13 * np.linalg.norm(b)
    nq, nc = 2 * E, E
    lam_rows = A[nq + nc:]
    assert np.abs(lam_rows @ x).max() < 1e-12  # flux conservation at every bifurcation
    # the multiplier blocks are transposes of each other, the pressure blocks negative transposes
    assert abs(A[nq + nc:, :nq] - A[:nq, nq + nc:].T).max() == 0.0
    assert abs(A[nq:nq + nc, :nq] + A[:nq, nq:nq + nc].T).max() == 0.0
    fb = 2 * nm.edge_slot.astype(np.int64)
    q0, q1 = x[fb], x[fb + 1]
    np.testing.assert_allclose(q0, q1, rtol=1e-9, atol=1e-13)  # f = 0: q constant along an edge
    # closed form: resistor network with boundary pressures -p_bc
    net = helpers.oracle_for(nm, 1)
    q_edge, lam = net.resistor_network_solution(net.eval_pbc(P_Y))
    assert helpers.rel_l2(x[nq + nc:], lam) < 1e-9
    assert helpers.rel_l2(q0, q_edge) < 1e-9
    # DG0 pressure of the single cell: p = lambda_u - R q h / 2 (inlet edge: -p_bc(x_u) - ...)
    h = net.cell_lengths()
    lu = net.lm_index[net.edges[:, 0]]
    pu = np.where(lu >= 0, lam[np.maximum(lu, 0)], -net.eval_pbc(P_Y)[net.edges[:, 0]])
    np.testing.assert_allclose(x[nq:nq + nc], pu - q_edge * h / 2, rtol=1e-8, atol=1e-11)
    # the headline step is four launches: assembly, one cooperative tree kernel (factorisation fused
    # with the first solve), back-substitution, residual.  A silently disabled cooperative path
    # (e.g. register growth -> one block per SM) shows up here as five or seven.
    fast = nxfx.Solver(asm, schedule=solver._schedule)
    fast.assemble()
    fast.solve()
    dev = nm.device
    l0 = dev.launch_count
    fast.assemble()
    fast.solve()
    assert dev.launch_count - l0 == 4
    x_fused = np.concatenate([f.x.array for f in fast.solve()])  # factors reused: separate solve kernel
    assert dev.launch_count - l0 == 7
    assert helpers.rel_l2(x_fused, x) < 1e-12


def _graph(points, edges):
    G = nx.DiGraph()
    for i, p in enumerate(points):
        G.add_node(i, pos=np.asarray(p, dtype=float))
    for e in edges:
        G.add_edge(*e)
    return G


def test_edge_cases_no_bifurcation_star_and_long_chain():
    """Degenerate inputs: a single edge (no multipliers at all), a degree-6 star, a 30-node line
    (tests/test_orientation.py graph) and a 200-node chain whose elimination tree has more levels
    than the shared-memory sweep kernels hold (global-memory fallback sweeps)."""
    cases = [
        (_graph([(0, 0, 0), (1, 2, 0.5)], [(0, 1)]), 5),
        (_graph([(0, 0, 0), (0, 1, 0)] + [(np.cos(k), 2 + np.sin(k), 0.1 * k) for k in range(5)],
                [(0, 1)] + [(1, 2 + k) for k in range(5)]), 3),
        (helpers.linear_graph(30, dim=3, ordered=lambda k: k % 3 != 0), 2),
        (helpers.linear_graph(200, dim=2), 1),
    ]
    for G, N in cases:
        for strategy in (None, "largest_first"):
            nm, asm, solver, sol, net, A, b = run_case(G, N, strategy, lambda x: 1.0 + x[0] + 2 * x[1])
            check_system(solver, net, A, b)
            check_solution(sol, net, A, b)
    assert not solver.assembler.network.device is None


def test_update_positions_and_reassemble():
    """Same topology, new coordinates / boundary data: re-assembly overwrites and the solve follows."""
    G = ng.make_tree(6, 2, 3)
    nm, asm, solver, sol, net, A, b = run_case(G, 2, "smallest_last", P_Y)
    x1 = np.concatenate([f.x.array for f in sol]).copy()
    asm.compute_forms(p_bc_ex=lambda x: 3.0 * x[1], R=2.0)
    solver.assemble()
    sol2 = solver.solve()
    A2, b2 = net.assemble(net.eval_pbc(lambda x: 3.0 * x[1]), R=2.0)
    check_system(solver, net, A2, b2)
    x2 = np.concatenate([f.x.array for f in sol2])
    assert helpers.rel_l2(x2, net.solve(A2, b2)) < 1e-10
    nq = net.poff
    assert helpers.rel_l2(x2[:nq], 1.5 * x1[:nq]) < 1e-10  # q ~ p_bc / R
    assert helpers.rel_l2(x2[nq:], 3.0 * x1[nq:]) < 1e-10  # p, lambda ~ p_bc


# ---- higher-order elements: table-driven assembly + exact condensation -----------------
def run_ho_case(G, N, strategy, fd, pd, p_bc, R=None, f=None, petsc_options=None):
    from oracle import reference_port as rp

    nm = nxfx.NetworkMesh(G, N=N, color_strategy=strategy)
    asm = nxfx.HydraulicNetworkAssembler(nm, flux_degree=fd, pressure_degree=pd)
    asm.compute_forms(p_bc_ex=p_bc, R=R, f=f)
    solver = nxfx.Solver(asm, petsc_options=petsc_options)
    solver.assemble()
    net = rp.OracleNetworkHO(nm._node_pos, nm.graph_edges, nm.edge_colors, N, fd, pd)
    A, b = net.assemble(net.eval_pbc(p_bc), R=1.0 if R is None else R, f=0.0 if f is None else f)
    return nm, asm, solver, net, A, b


@pytest.mark.parametrize("fd,pd", [(2, 1), (2, 0), (1, 1), (3, 2)])
def test_higher_order_assembly_and_solve(fd, pd):
    rng = np.random.default_rng(fd * 10 + pd)
    for G, N, strategy in ((ng.make_tree(2, 1, 3), 4, None), (helpers.random_tree(60, 3), 2, "smallest_last"),
                           (helpers.edge_info_graph(), 3, "largest_first")):
        nc = N * G.number_of_edges()
        R, f = rng.uniform(0.5, 2.0, nc), rng.normal(size=nc)
        nm, asm, solver, net, A, b = run_ho_case(G, N, strategy, fd, pd, lambda x: x[1] + 0.5 * x[2], R=R, f=f)
        assert asm.is_generic == ((fd, pd) != (1, 0)) and asm.block_sizes == net.block_sizes
        rp_, ci, va = solver.A.getValuesCSR()
        assert np.array_equal(rp_, A.indptr) and np.array_equal(ci, A.indices), "pattern differs"
        np.testing.assert_allclose(va, A.data, rtol=1e-12, atol=5e-14)
        np.testing.assert_allclose(solver.b.array_r, b, rtol=1e-12, atol=5e-14)
        # SpMV on the higher-order matrix
        xv, yv = solver.x.duplicate(), solver.x.duplicate()
        xh = rng.normal(size=net.n_dofs)
        xv.array[:] = xh
        solver.A.mult(xv, yv)
        np.testing.assert_allclose(yv.array_r, A @ xh, rtol=1e-12, atol=1e-12)
        if fd != pd + 1:
            continue  # e.g. P1/P1 is not inf-sup stable: the matrix is singular, only assembly is checked
        # default options (preonly + lu): the exact condensation; FGMRES around it on the graph with cycles
        sol = solver.solve()
        x = np.concatenate([fn.x.array for fn in sol])
        x_ref = net.solve(A, b)
        assert helpers.rel_l2(x, x_ref) < 1e-10, (fd, pd, solver.ksp.getIterationNumber(), helpers.rel_l2(x, x_ref))
        if solver._schedule.is_forest:
            assert solver.ksp.getIterationNumber() == 1  # one application, no correction needed
        else:
            assert solver.ksp.getIterationNumber() <= 12
        assert len(sol) == nm.num_edge_colors + 2
        gq = nxfx.post_processing.extract_global_flux(nm, sol)
        assert gq.x.array.size == (fd + 1) * nc
        np.testing.assert_array_equal(gq.x.array.reshape(nc, fd + 1), x[net.cell_flux_dofs()])


@pytest.mark.parametrize("fd,pd,n", [(2, 1, 16), (3, 2, 16), (2, 0, 14), (4, 3, 12)])
def test_higher_order_direct_solve_at_size(fd, pd, n):
    """``preonly + lu`` on P2/P1, P3/P2 (and DG0 / higher pairs) is ONE application of the exact
    condensation plus the residual check -- no Krylov iterations -- on make_tree(16) with 4 cells per edge, and
    agrees with SuperLU on the oracle's matrix to 1e-10 (bar: 1e-8)."""
    import scipy.sparse.linalg as spla

    G = ng.make_tree(n, float(n), float(n), as_arrays=True)
    N = 4
    nc = N * G.edges.shape[0]
    rng = np.random.default_rng(n)
    R, f = rng.uniform(0.5, 2.0, nc), rng.normal(size=nc)
    opts = {"ksp_type": "preonly", "pc_type": "lu", "ksp_error_if_not_converged": True}
    nm, asm, solver, net, A, b = run_ho_case(G, N, "smallest_last", fd, pd, P_Y, R=R, f=f, petsc_options=opts)
    launches0 = nm.device.launch_count
    sol = solver.solve()
    launches = nm.device.launch_count - launches0
    assert solver.ksp.getIterationNumber() == 1 and solver.ksp.reason > 0
    assert solver.info.residual_norm <= 1e-12 * solver.info.rhs_norm
    # factor (edge LU, node blocks, 2 sweeps) + apply (edge rhs, node rhs, 3 sweeps, back-substitution) + residual
    assert launches <= 11, launches
    x = np.concatenate([fn.x.array for fn in sol])
    x_ref = spla.spsolve(A.tocsc(), b)
    assert helpers.rel_l2(x, x_ref) < 1e-10, helpers.rel_l2(x, x_ref)
    # a second solve with the same matrix reuses the factors: apply + residual only
    solver.b.array[:] = rng.normal(size=net.n_dofs)
    launches0 = nm.device.launch_count
    sol = solver.solve()
    # (one refinement correction -- residual pass, application -- when the first iterate is not yet at 1e-13)
    assert nm.device.launch_count - launches0 <= (7 if solver.ksp.getIterationNumber() == 1 else 15)
    x = np.concatenate([fn.x.array for fn in sol])
    r = solver.b.array_r - A @ x
    assert np.linalg.norm(r) <= 1e-12 * np.linalg.norm(solver.b.array_r)


def test_condensation_tables_are_validated():
    """nxfx_set_condensation rejects tables that do not fit the network (bad bandwidth, entries outside the local
    block, unknown kinds) instead of factorising garbage; without the tables a direct solve is refused loudly."""
    import ctypes as C

    from networks_fenicsx_b200 import _lib
    from networks_fenicsx_b200.condense import build_condensation

    nm = nxfx.NetworkMesh(ng.make_tree(4, 1, 1), N=2, color_strategy="smallest_last")
    asm = nxfx.HydraulicNetworkAssembler(nm, flux_degree=2, pressure_degree=1)
    asm.compute_forms(p_bc_ex=P_Y)
    solver = nxfx.Solver(asm)
    solver.assemble()
    cd = build_condensation(nm, 2, 1)
    t = cd.packed()
    c32 = lambda a: _lib.as_i32p(np.ascontiguousarray(a, dtype=np.int32))  # noqa: E731
    c64 = lambda a: _lib.as_f64p(np.ascontiguousarray(a, dtype=np.float64))  # noqa: E731

    def call(kl=cd.kl, n_max=cd.n_max, **over):
        a = {**t, **over}
        nm.device.call(
            "nxfx_set_condensation", 1, cd.fd * cd.N + 1, n_max, kl, cd.pcell_base, cd.pcell_stride, c32(a["type_n"]),
            c32(a["loc_ptr"]), c32(a["loc_kind"]), c32(a["loc_off"]), c32(a["k_ptr"]), c32(a["k_row"]), c32(a["k_col"]),
            c32(a["k_cell"]), c64(a["k_coef"]), c32(a["c_ptr"]), c32(a["c_row"]), c32(a["c_slot"]), c64(a["c_coef"]),
            c32(a["d_ptr"]), c32(a["d_slot"]), c32(a["d_col"]), c64(a["d_coef"]), c32(cd.bif_node))

    with pytest.raises(RuntimeError, match="K entry"):
        call(kl=1)  # entries outside the declared band
    with pytest.raises(RuntimeError, match="bad type_n"):
        call(n_max=cd.n_max - 1)
    bad_kind = t["loc_kind"].copy()
    bad_kind[0] = 9
    with pytest.raises(RuntimeError, match="unknown kind"):
        call(loc_kind=bad_kind)
    bad_slot = t["c_slot"].copy()
    bad_slot[0] = 4
    with pytest.raises(RuntimeError, match="C entry"):
        call(c_slot=bad_slot)
    # the failed calls left the context without tables: a direct solve says so
    with pytest.raises(RuntimeError, match="nxfx_set_condensation"):
        solver.solve()
    call()  # the right tables again
    x = np.concatenate([fn.x.array for fn in solver.solve()])
    assert np.isfinite(x).all() and solver.ksp.getIterationNumber() == 1


def test_higher_order_long_edges():
    """40 cells per edge (the arterial demo's refinement): the per-edge band no longer fits the shared-memory
    variant of the LU, the global-memory variant takes over; same direct solve."""
    G = ng.make_arterial_tree(N=5, direction=np.array([0.1, 1.0, 0.0]))
    N = 40
    nc = N * G.number_of_edges()
    rng = np.random.default_rng(40)
    R, f = rng.uniform(0.5, 2.0, nc), rng.normal(size=nc)
    for fd, pd in ((2, 1), (3, 2), (2, 0)):
        nm, asm, solver, net, A, b = run_ho_case(G, N, nx.coloring.strategy_largest_first, fd, pd, P_Y, R=R, f=f)
        sol = solver.solve()
        x = np.concatenate([fn.x.array for fn in sol])
        assert helpers.rel_l2(x, net.solve(A, b)) < 1e-9, (fd, pd)
        assert solver.ksp.getIterationNumber() <= 2


def test_higher_order_accumulated_and_rhs_only():
    """ADD_VALUES twice and an rhs-only reassembly after changing R on the table-driven path: R*h accumulates with
    the values, the condensation follows the matrix as it stands (same contract as the P1/DG0 path)."""
    G = helpers.random_tree(50, 11)
    N, fd, pd = 3, 2, 1
    nc = N * G.number_of_edges()
    rng = np.random.default_rng(3)
    R1, R2 = rng.uniform(0.5, 2.0, nc), rng.uniform(0.5, 2.0, nc)
    nm, asm, solver, net, A1, b1 = run_ho_case(G, N, "smallest_last", fd, pd, P_Y, R=R1)
    asm.compute_forms(p_bc_ex=P_Y, R=R2)
    asm.assemble(solver.A, solver.b)  # no zeroEntries: A = A(R1) + A(R2), b = b1 + b2
    A2, b2 = net.assemble(net.eval_pbc(P_Y), R=R2)
    sol = solver.solve()
    x = np.concatenate([fn.x.array for fn in sol])
    assert helpers.rel_l2(x, net.solve((A1 + A2).tocsr(), b1 + b2)) < 1e-10
    assert solver.ksp.getIterationNumber() == 1
    # rhs only, other R: the matrix (and its factorisation) stay
    asm.compute_forms(p_bc_ex=lambda x: 2.0 * x[1], R=R1)
    solver.assemble(lhs=False, rhs=True)
    _, b3 = net.assemble(net.eval_pbc(lambda x: 2.0 * x[1]), R=R1)
    sol = solver.solve()
    x = np.concatenate([fn.x.array for fn in sol])
    assert helpers.rel_l2(x, net.solve((A1 + A2).tocsr(), b3)) < 1e-10


def test_higher_order_matches_low_order_flux():
    """f = 0: the exact flux is constant per edge, so P2/P1 and P1/DG0 give the same flux and
    multipliers (a physical cross-check independent of the oracle's higher-order tables)."""
    G = ng.make_tree(6, 3, 4)
    out = {}
    for fd, pd in ((1, 0), (2, 1)):
        nm = nxfx.NetworkMesh(G, N=3, color_strategy="smallest_last")
        asm = nxfx.HydraulicNetworkAssembler(nm, flux_degree=fd, pressure_degree=pd)
        asm.compute_forms(p_bc_ex=P_Y, R=1.5)
        solver = nxfx.Solver(asm)
        solver.assemble()
        sol = solver.solve()
        per_edge = fd * 3 + 1
        q_first = np.concatenate([fn.x.array[::per_edge] for fn in sol[:-2]])
        out[(fd, pd)] = (q_first, sol[-1].x.array.copy())
    np.testing.assert_allclose(out[(2, 1)][0], out[(1, 0)][0], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(out[(2, 1)][1], out[(1, 0)][1], rtol=1e-8, atol=1e-11)


def test_adaptive_refinement():
    """The refinement step is applied only when the first direct solve is not yet at refine_rtol
    (decided on the device); forcing it gives the same answer."""
    G = ng.make_tree(9, 4, 5)
    rng = np.random.default_rng(5)
    R = 10.0 ** rng.uniform(-4, 4, G.number_of_edges() * 2)  # badly scaled resistances
    its = {}
    xs = {}
    for name, extra in (("default", {}), ("forced", {"nxfx_refine_rtol": 0.0}), ("tight", {"nxfx_refine_rtol": 1e-30}),
                        ("none", {"nxfx_refine_steps": 0})):
        opts = {"ksp_type": "preonly", "pc_type": "lu", "nxfx_final_residual": True, **extra}
        nm, asm, solver, sol, net, A, b = run_case(G, 2, "smallest_last", P_Y, R=R, petsc_options=opts)
        its[name] = solver.ksp.getIterationNumber()
        xs[name] = np.concatenate([f.x.array for f in sol])
        assert helpers.rel_l2(xs[name], net.solve(A, b)) < 1e-8
        assert solver.info.residual_norm <= 1e-10 * solver.info.rhs_norm
    assert its["forced"] == 2 and its["tight"] == 2 and its["none"] == 1 and its["default"] in (1, 2)
    assert helpers.rel_l2(xs["forced"], xs["tight"]) == 0.0


@pytest.mark.parametrize("seed", range(8))
def test_random_graphs_with_random_orientation(seed):
    """Random recursive trees with random edge directions (several inlets / outlets per component)
    plus, for odd seeds, a few cycle-closing edges: pattern, values and solution against the oracle."""
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(5, 60))
    G = nx.DiGraph()
    for i in range(n):
        G.add_node(i, pos=rng.normal(size=3))
    und = set()
    for k in range(1, n):
        p = 0 if k == 1 else int(rng.integers(0, k))
        und.add((p, k))
        G.add_edge(*((p, k) if rng.random() < 0.5 else (k, p)))
    if seed % 2:
        for _ in range(3):
            a, b = (int(v) for v in rng.integers(0, n, 2))
            if a != b and (min(a, b), max(a, b)) not in und:
                und.add((min(a, b), max(a, b)))
                G.add_edge(a, b)
    N = int(rng.integers(1, 5))
    nc = N * G.number_of_edges()
    nm, asm, solver, sol, net, A, b = run_case(G, N, "largest_first" if seed % 3 else None,
                                               lambda x: x[0] - x[1] + 0.3 * x[2],
                                               R=rng.uniform(0.5, 2.0, nc), f=rng.normal(size=nc))
    check_system(solver, net, A, b)
    check_solution(sol, net, A, b, tol=1e-9)
    assert solver._schedule.is_forest == (A.shape[0] > 0 and nx.is_forest(G.to_undirected()))


def test_reversed_tree_many_inlets():
    A_ = ng.make_tree(9, 2, 3, as_arrays=True)
    rev = ng.ArrayGraph(A_.pos, A_.edges[:, ::-1].copy())
    nm, asm, solver, sol, net, A, b = run_case(rev, 2, "smallest_last", P_Y)
    assert solver._schedule.is_forest
    check_system(solver, net, A, b)
    check_solution(sol, net, A, b)


def test_top_chunk_with_many_child_links():
    """Regression: the top chunk lists its own links AND the roots of all bottom chunks; with tiny
    bottom chunks that is about twice its node count (here 2047 nodes, 4094 links), which used to
    overflow the shared-memory child table (first seen on a 23-generation tree)."""
    from networks_fenicsx_b200.schedule import build_tree_schedule

    G = ng.make_tree(14, 3, 5, as_arrays=True)
    nm = nxfx.NetworkMesh(G, N=1, color_strategy="smallest_last")
    asm = nxfx.HydraulicNetworkAssembler(nm)
    asm.compute_forms(p_bc_ex=P_Y)
    sched = build_tree_schedule(nm.graph_edges, nm.node_multiplier_index, nm.bifurcation_values.size,
                                root_hint_nodes=nm._boundary_out_nodes, chunk_nodes=4)
    top = sched.lvl_ptr[sched.chunk_lptr[-1]] - sched.lvl_ptr[sched.chunk_lptr[-2]]
    links = sched.t_cptr[sched.lvl_ptr[sched.chunk_lptr[-1]]] - sched.t_cptr[sched.lvl_ptr[sched.chunk_lptr[-2]]]
    assert top == 2047 and links == 4094
    solver = nxfx.Solver(asm, schedule=sched, petsc_options={"ksp_type": "preonly", "pc_type": "lu",
                                                             "nxfx_final_residual": True, "nxfx_refine_steps": 0})
    solver.assemble()
    sol = solver.solve()
    net = helpers.oracle_for(nm, 1)
    A, b = net.assemble(net.eval_pbc(P_Y))
    check_solution(sol, net, A, b)
    assert solver.info.residual_norm <= 1e-12 * solver.info.rhs_norm


# ---- matrices are independent objects; the factorisation follows the matrix -----------
def _x_of(sol):
    return np.concatenate([f.x.array for f in sol])


@pytest.mark.parametrize("N", [1, 3])
def test_solve_of_accumulated_matrix(N):
    """ADD_VALUES without zeroEntries (assembly.py:355,362): after a second assemble(A, b) the matrix
    is 2 A0 -- the solver must factorise THAT matrix, not the element data of one assembly."""
    G = ng.make_tree(7, 3, 4)
    rng = np.random.default_rng(11)
    nc = N * G.number_of_edges()
    R1, R2 = rng.uniform(0.5, 2.0, nc), rng.uniform(0.1, 3.0, nc)
    nm = nxfx.NetworkMesh(G, N=N, color_strategy="smallest_last")
    asm = nxfx.HydraulicNetworkAssembler(nm)
    asm.compute_forms(p_bc_ex=P_Y, R=R1)
    opts = {"ksp_type": "preonly", "pc_type": "lu", "ksp_error_if_not_converged": True, "nxfx_final_residual": True}
    solver = nxfx.Solver(asm, petsc_options=opts)
    solver.assemble()
    x0 = _x_of(solver.solve()).copy()
    net = helpers.oracle_for(nm, N, G, "smallest_last")
    A1, b1 = net.assemble(net.eval_pbc(P_Y), R=R1)
    assert helpers.rel_l2(x0, net.solve(A1, b1)) < 1e-10
    # (i) the same system twice: 2 A0 x = 2 b0
    asm.assemble(solver.A, solver.b)
    assert solver.A.accumulated == 2
    assert np.array_equal(solver.A.getValuesCSR()[2], 2 * A1.data)
    x = _x_of(solver.solve())
    assert solver.info.residual_norm <= 1e-12 * solver.info.rhs_norm
    assert helpers.rel_l2(x, x0) < 1e-10
    # (ii) a third assembly with ANOTHER resistance: A = 2 A(R1) + A(R2), b = 3 b0
    asm.compute_forms(p_bc_ex=P_Y, R=R2)
    asm.assemble(solver.A, solver.b)
    A2, _ = net.assemble(net.eval_pbc(P_Y), R=R2)
    Asum = (2 * A1 + A2).tocsr()
    x = _x_of(solver.solve())
    assert solver.info.residual_norm <= 1e-12 * solver.info.rhs_norm
    assert helpers.rel_l2(x, net.solve(Asum, 3 * b1)) < 1e-9
    # (iii) GMRES on the accumulated matrix with the rescaled Schur preconditioner: exact => 1-2 iterations
    s2 = nxfx.Solver(asm, petsc_options={"ksp_type": "gmres", "pc_type": "lu", "ksp_rtol": 1e-12})
    s2.assemble()
    asm.assemble(s2.A, s2.b)
    x = _x_of(s2.solve())
    assert s2.ksp.getIterationNumber() <= 3
    assert helpers.rel_l2(x, net.solve(A2, b1)) < 1e-9


def test_rhs_only_reassembly_keeps_factorisation_of_the_matrix():
    """compute_forms(R=new) + assemble(assemble_lhs=False): A still holds the OLD resistance and the
    solve must use it (the factorisation is built from data written with the matrix)."""
    G = ng.make_tree(8, 3, 4)
    for N in (1, 2):
        nm = nxfx.NetworkMesh(G, N=N, color_strategy="smallest_last")
        asm = nxfx.HydraulicNetworkAssembler(nm)
        asm.compute_forms(p_bc_ex=P_Y, R=2.0)
        solver = nxfx.Solver(asm, petsc_options={"ksp_type": "preonly", "pc_type": "lu", "nxfx_final_residual": True,
                                                 "ksp_error_if_not_converged": True})
        solver.assemble()
        x_old = _x_of(solver.solve()).copy()
        asm.compute_forms(p_bc_ex=lambda x: 2.0 * x[1], R=7.0)
        solver.assemble(lhs=False, rhs=True)  # new rhs (p_bc doubled), matrix untouched
        x = _x_of(solver.solve())
        assert solver.info.residual_norm <= 1e-12 * solver.info.rhs_norm
        assert solver.ksp.getIterationNumber() == 1, "the stale-factorisation repair path was taken"
        assert helpers.rel_l2(x, 2.0 * x_old) < 1e-10
        solver.assemble()  # now the matrix follows: q ~ p_bc / R
        x_new = _x_of(solver.solve())
        nq = sum(asm.block_sizes[:-2])
        assert helpers.rel_l2(x_new[:nq], x_old[:nq] * 2.0 * 2.0 / 7.0) < 1e-10


def test_two_assemblers_and_solvers_on_one_network_do_not_alias():
    """Parameter study on ONE NetworkMesh: independent matrices, boundary data and solutions
    (the reference returns independent PETSc objects)."""
    G = ng.make_tree(6, 2, 3)
    nm = nxfx.NetworkMesh(G, N=2, color_strategy="smallest_last")
    net = helpers.oracle_for(nm, 2, G, "smallest_last")
    asm1 = nxfx.HydraulicNetworkAssembler(nm)
    asm2 = nxfx.HydraulicNetworkAssembler(nm)
    asm1.compute_forms(p_bc_ex=P_Y, R=1.0)
    asm2.compute_forms(p_bc_ex=P_X, R=5.0)
    s1, s2 = nxfx.Solver(asm1), nxfx.Solver(asm2)
    s1.assemble()
    s2.assemble()
    A1, b1 = net.assemble(net.eval_pbc(P_Y), R=1.0)
    A2, b2 = net.assemble(net.eval_pbc(P_X), R=5.0)
    assert np.array_equal(s1.A.getValuesCSR()[2], A1.data) and np.array_equal(s1.b.array_r, b1)
    assert np.array_equal(s2.A.getValuesCSR()[2], A2.data) and np.array_equal(s2.b.array_r, b2)
    x1 = _x_of(s1.solve())  # s2 was assembled last: s1 must still solve ITS system
    x2 = _x_of(s2.solve())
    x1b = _x_of(s1.solve())  # and again after the factors were rebuilt for the other matrix
    assert helpers.rel_l2(x1, net.solve(A1, b1)) < 1e-10
    assert helpers.rel_l2(x2, net.solve(A2, b2)) < 1e-10
    assert helpers.rel_l2(x1b, x1) < 1e-14
    # a matrix returned by assemble() is a third independent object; a fresh one multiplies as zero
    A3, b3 = asm1.assemble()
    assert np.array_equal(A3.getValuesCSR()[2], A1.data) and np.array_equal(s2.A.getValuesCSR()[2], A2.data)
    xv, yv = s1.x.duplicate(), s1.x.duplicate()
    xv.array[:] = 1.0
    fresh = asm1.create_matrix()
    fresh.mult(xv, yv)
    assert not yv.array_r.any()
    A3.mult(xv, yv)
    assert np.array_equal(yv.array_r, A1 @ np.ones(net.n_dofs))
    # an assembler with other polynomial degrees needs its own NetworkMesh (one pattern per context)
    asm_ho = nxfx.HydraulicNetworkAssembler(nm, flux_degree=2, pressure_degree=1)
    asm_ho.compute_forms(p_bc_ex=P_Y)
    with pytest.raises(RuntimeError, match="second NetworkMesh"):
        nxfx.Solver(asm_ho)


def test_solution_mirror_matches_the_staged_download():
    """Solver.create_functions(): views into one pinned buffer that the library fills on a side stream while
    the residual check runs (nxfx_set_solution_mirror) -- also when a refinement correction changes x after
    the first copy, with FGMRES, and without leaking the mirror into later plain solves."""
    G = ng.make_tree(9, 4, 5)
    rng = np.random.default_rng(2)
    R = 10.0 ** rng.uniform(-3, 3, G.number_of_edges() * 2)
    for opts in ({"ksp_type": "preonly", "pc_type": "lu"},
                 {"ksp_type": "preonly", "pc_type": "lu", "nxfx_refine_rtol": 0.0, "nxfx_refine_steps": 2},
                 {"ksp_type": "gmres", "pc_type": "lu", "ksp_rtol": 1e-13}):
        nm, asm, solver, sol, net, A, b = run_case(G, 2, "smallest_last", P_Y, R=R, petsc_options=opts)
        x_staged = _x_of(sol).copy()
        fns = solver.create_functions()
        assert all(fn.x.array.base is not None for fn in fns)
        solver.assemble()
        sol2 = solver.solve(fns)
        assert sol2 is fns
        assert np.array_equal(_x_of(fns), x_staged)
        assert np.array_equal(solver.x.array_r, x_staged)
        fns[0].x.array[:] = -7.0  # a later solve without these functions must not touch the mirror
        solver.assemble()
        solver.solve()
        assert np.all(fns[0].x.array == -7.0)
