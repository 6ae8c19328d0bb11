"""Reference-element tables of the equispaced Lagrange elements the assembler uses
(assembly.py:127-145): flux P_fd, pressure P_pd on the unit interval, dof order
``[X=0, X=1, interior ascending]``.

``M_ref[a,b] = int phi_a phi_b``, ``B_ref[r,a] = int psi_r phi_a'``, ``w_ref[r] = int psi_r``;
element tensors on a cell of length h: ``M_e = R h M_ref``, ``B_e = B_ref`` (independent of h and
of the embedding), ``L_e = f h w_ref``.  Built in exact rational arithmetic (the nodal basis in product form, exact integration of the product
polynomials) and rounded to binary64 once, so every table entry is the correctly rounded value -- the
first version solved a Vandermonde system in floating point and was only good to 2e-12 at degree 4
(found by the quadrature literal in oracle/ffcx_literal.py).
"""

from __future__ import annotations

import functools
from fractions import Fraction

import numpy as np


def nodes(degree: int) -> np.ndarray:
    if degree == 0:
        return np.array([0.5])
    return np.concatenate([[0.0, 1.0], np.arange(1, degree) / degree])


def _basis_coefficients(degree: int) -> list[list[Fraction]]:
    """Row i = monomial coefficients (ascending powers, exact rationals) of the i-th nodal basis function
    ``prod_{j != i} (X - x_j) / (x_i - x_j)``."""
    if degree == 0:
        return [[Fraction(1)]]
    x = [Fraction(0), Fraction(1)] + [Fraction(i, degree) for i in range(1, degree)]
    rows = []
    for i in range(degree + 1):
        c = [Fraction(1)]
        for j in range(degree + 1):
            if j == i:
                continue
            d = x[i] - x[j]
            nxt = [Fraction(0)] * (len(c) + 1)
            for k, ck in enumerate(c):  # multiply by (X - x_j) / d
                nxt[k] += ck * (-x[j]) / d
                nxt[k + 1] += ck / d
            c = nxt
        rows.append(c)
    return rows


def _integrate_product(p, q) -> float:
    """int_0^1 p(X) q(X) dX for ascending-power rational coefficient vectors: exact, rounded once."""
    prod = [Fraction(0)] * (len(p) + len(q) - 1)
    for i, pi in enumerate(p):
        for j, qj in enumerate(q):
            prod[i + j] += pi * qj
    return float(sum(c / (k + 1) for k, c in enumerate(prod)))


@functools.lru_cache(maxsize=None)
def tables(flux_degree: int, pressure_degree: int):
    fd, pd = int(flux_degree), int(pressure_degree)
    if fd < 1 or pd < 0:
        raise ValueError("flux_degree >= 1 and pressure_degree >= 0 required")
    phi, psi = _basis_coefficients(fd), _basis_coefficients(pd)
    dphi = [[k * c for k, c in enumerate(row)][1:] for row in phi]
    M = np.array([[_integrate_product(phi[a], phi[b]) for b in range(fd + 1)] for a in range(fd + 1)])
    B = np.array([[_integrate_product(psi[r], dphi[a]) for a in range(fd + 1)] for r in range(pd + 1)])
    w = np.array([_integrate_product(psi[r], [Fraction(1)]) for r in range(pd + 1)])
    trace0 = np.zeros(fd + 1)
    trace1 = np.zeros(fd + 1)
    trace0[0] = 1.0  # the vertex trace of a nodal basis is exactly 0/1
    trace1[1] = 1.0
    return M, B, w, trace0, trace1
