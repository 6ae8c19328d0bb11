"""Reference-element tables of the equispaced Lagrange elements the assembler uses
(assembly.py:127-145): flux P_fd, pressure P_pd on the unit interval, dof order
``[X=0, X=1, interior ascending]``.

``M_ref[a,b] = int phi_a phi_b``, ``B_ref[r,a] = int psi_r phi_a'``, ``w_ref[r] = int psi_r``;
element tensors on a cell of length h: ``M_e = R h M_ref``, ``B_e = B_ref`` (independent of h and
of the embedding), ``L_e = f h w_ref``.  Built from the monomial coefficients of the
nodal basis (Vandermonde solve) and exact integration of the product polynomials.
"""

from __future__ import annotations

import functools

import numpy as np


def nodes(degree: int) -> np.ndarray:
    if degree == 0:
        return np.array([0.5])
    return np.concatenate([[0.0, 1.0], np.arange(1, degree) / degree])


def _basis_coefficients(degree: int) -> np.ndarray:
    """Row i = monomial coefficients (ascending powers) of the i-th nodal basis function."""
    x = nodes(degree)
    V = np.vander(x, degree + 1, increasing=True)  # V[i, k] = x_i^k
    return np.linalg.solve(V, np.eye(degree + 1)).T


def _integrate_product(p: np.ndarray, q: np.ndarray) -> float:
    """int_0^1 p(X) q(X) dX for ascending-power coefficient vectors."""
    prod = np.convolve(p, q)
    return float(np.sum(prod / np.arange(1, prod.size + 1)))


@functools.lru_cache(maxsize=None)
def tables(flux_degree: int, pressure_degree: int):
    fd, pd = int(flux_degree), int(pressure_degree)
    if fd < 1 or pd < 0:
        raise ValueError("flux_degree >= 1 and pressure_degree >= 0 required")
    phi, psi = _basis_coefficients(fd), _basis_coefficients(pd)
    dphi = phi[:, 1:] * np.arange(1, fd + 1)[None, :]
    M = np.array([[_integrate_product(phi[a], phi[b]) for b in range(fd + 1)] for a in range(fd + 1)])
    B = np.array([[_integrate_product(psi[r], dphi[a]) for a in range(fd + 1)] for r in range(pd + 1)])
    w = np.array([_integrate_product(psi[r], np.array([1.0])) for r in range(pd + 1)])
    trace0 = np.zeros(fd + 1)
    trace1 = np.zeros(fd + 1)
    trace0[0] = 1.0  # the vertex trace of a nodal basis is exactly 0/1
    trace1[1] = 1.0
    return M, B, w, trace0, trace1
