"""bench.py's host-side parity checks (matrix-free residual, Kirchhoff sums, Thevenin closed form),
validated on the CPU against the oracle's direct solve: a correct solution passes, a perturbed one fails."""

import types

import numpy as np
import pytest

import bench
import networks_fenicsx_b200 as nxfx
from networks_fenicsx_b200 import network_generation as ng
from oracle import reference_port as rp


def fake_problem(G, N, R_edge, f_cell, x, nm, net):
    asm = types.SimpleNamespace(_pbc_host=net.eval_pbc(bench.p_bc))
    solver = types.SimpleNamespace(x=types.SimpleNamespace(array_r=x))
    nm._x_host = net.x3  # stands in for the device-generated vertices
    return types.SimpleNamespace(nm=nm, asm=asm, solver=solver, N=N, ds=None, G=G, R_edge=R_edge, f_cell=f_cell)


@pytest.mark.parametrize("workload,n,N", [("tree", 7, 1), ("tree", 5, 3), ("arterial", 6, 1), ("arterial", 5, 4)])
def test_host_parity_accepts_the_direct_solution_and_rejects_a_wrong_one(workload, n, N):
    G, R, f = bench.make_workload(workload, n, N)
    nm = nxfx.NetworkMesh(G, N=N, color_strategy="smallest_last")
    net = rp.OracleNetwork(G.pos, G.edges, nm.edge_colors, N)
    A, b = net.assemble(net.eval_pbc(bench.p_bc), R=1.0 if R is None else np.repeat(R, N), f=0.0 if f is None else f)
    x = net.solve(A, b)
    par = bench.host_parity(fake_problem(G, N, R, f, x, nm, net), None, None)
    true_res = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
    assert par["true_residual_recomputed"] < 1e-12 and abs(par["true_residual_recomputed"] - true_res) < 1e-13
    assert par["kirchhoff_max"] < 1e-12
    if f is None:
        assert par["q_const_max"] < 1e-10
        assert par["rel_l2_vs_closed_form"]["flux"] < 1e-10 and par["rel_l2_vs_closed_form"]["multipliers"] < 1e-10
    else:
        assert par["rel_l2_vs_closed_form"] is None
    bad = x.copy()
    bad[net.loff + 1] += 1e-3
    par_bad = bench.host_parity(fake_problem(G, N, R, f, bad, nm, net), None, None)
    assert par_bad["true_residual_recomputed"] > 1e-6
    if f is None:
        assert par_bad["rel_l2_vs_closed_form"]["multipliers"] > 1e-6


def test_closed_form_with_radius_dependent_resistance():
    """f = 0 with R_e = 8 mu / (pi r^4): the Thevenin reduction equals the oracle's resistor-network solve."""
    G, R, _ = bench.make_workload("arterial", 7, 1)
    nm = nxfx.NetworkMesh(G, N=2, color_strategy="smallest_last")
    net = rp.OracleNetwork(G.pos, G.edges, nm.edge_colors, 2)
    A, b = net.assemble(net.eval_pbc(bench.p_bc), R=np.repeat(R, 2))
    x = net.solve(A, b)
    par = bench.host_parity(fake_problem(G, 2, R, None, x, nm, net), None, None)
    assert par["rel_l2_vs_closed_form"]["flux"] < 1e-9 and par["rel_l2_vs_closed_form"]["multipliers"] < 1e-9
    q_edge, lam = net.resistor_network_solution(net.eval_pbc(bench.p_bc), R=np.repeat(R, 2))
    assert np.allclose(x[net.loff:], lam, rtol=1e-9, atol=1e-12)
