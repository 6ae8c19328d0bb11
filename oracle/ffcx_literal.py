"""Second, independent derivation of the element tensors of the hot path: the way an FFCx-generated
``tabulate_tensor`` evaluates them -- by Gauss quadrature on the reference interval from the cell's
coordinate dofs, with ``detJ = ||J||`` and the pseudo-inverse ``K = J^T / ||J||^2`` of the 3x1 Jacobian
of an interval embedded in space -- instead of the closed forms ``R h M_ref`` / ``B_ref`` used by
``reference_port.py`` and by the CUDA kernels.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Forms restated (upstream
``src/networks_fenicsx/assembly.py``):

* ``:253``  ``a[i][i] = R * inner(q, v) * dx``                      -> ``M[a, b] = sum_g w_g R phi_a phi_b detJ``
* ``:254``  ``a[P][i] = + phi * inner(grad(q), t) * dx``            -> ``B[r, a] = sum_g w_g psi_r (grad phi_a . t) detJ``
* ``:255``  ``a[i][P] = - p * inner(grad(v), t) * dx``              -> ``-B^T``
* ``:262``  ``L[P] = f * phi * dx``                                 -> ``L[r] = sum_g w_g f psi_r detJ``
* ``:238-242`` ``t = orientation * J[:, 0] / ||J[:, 0]||``.

``grad phi = K^T dphi/dX`` is the physical (3-vector) gradient of a reference basis function.  The
Lagrange bases are evaluated in barycentric form at the quadrature points (no monomial coefficients, no
Vandermonde solve: a third way of getting the same numbers as ``elements.py`` and ``lagrange_tables``).
"""

from __future__ import annotations

import numpy as np


def _nodes(degree: int) -> np.ndarray:
    """Equispaced Lagrange nodes, dof order [X=0, X=1, interior ascending] (assembly.py:127-132)."""
    if degree == 0:
        return np.array([0.5])
    return np.concatenate([[0.0, 1.0], np.arange(1, degree) / degree])


def _lagrange(degree: int, X: np.ndarray):
    """(values[n_dofs, n_pts], derivatives[n_dofs, n_pts]) of the nodal basis at the points X, from the
    product form  phi_i = prod_{j != i} (X - x_j) / (x_i - x_j)  and its product-rule derivative."""
    x = _nodes(degree)
    n = x.size
    val = np.ones((n, X.size))
    der = np.zeros((n, X.size))
    if degree == 0:
        return val, der
    for i in range(n):
        others = [j for j in range(n) if j != i]
        denom = np.prod([x[i] - x[j] for j in others])
        val[i] = np.prod([X - x[j] for j in others], axis=0) / denom
        acc = np.zeros_like(X)
        for k in others:
            acc += np.prod([X - x[j] for j in others if j != k], axis=0) if len(others) > 1 else np.ones_like(X)
        der[i] = acc / denom
    return val, der


def cell_tensors(x0, x1, flux_degree: int, pressure_degree: int, R: float = 1.0, f: float = 0.0,
                 orientation: float = 1.0):
    """Element tensors of ONE interval cell with end points ``x0, x1`` (3-vectors):
    ``(M[fd+1, fd+1], B[pd+1, fd+1], L[pd+1])`` -- mass, divergence coupling, source vector."""
    fd, pd = int(flux_degree), int(pressure_degree)
    x0, x1 = np.asarray(x0, dtype=np.float64), np.asarray(x1, dtype=np.float64)
    J = (x1 - x0).reshape(3, 1)  # affine cell: x(X) = x0 (1 - X) + x1 X
    detJ = float(np.sqrt((J * J).sum()))
    K = J.T / detJ**2  # pseudo-inverse, 1 x 3
    t = orientation * J[:, 0] / np.linalg.norm(J[:, 0])
    gx, gw = np.polynomial.legendre.leggauss(fd + pd + 2)  # exact for the polynomial integrands
    X, W = 0.5 * (gx + 1.0), 0.5 * gw
    phi, dphi = _lagrange(fd, X)
    psi, _ = _lagrange(pd, X)
    M = np.zeros((fd + 1, fd + 1))
    B = np.zeros((pd + 1, fd + 1))
    L = np.zeros(pd + 1)
    for g in range(X.size):
        grad = K.T * dphi[:, g][None, :]  # [3, fd+1]: physical gradient of every flux basis function
        dq_dt = grad.T @ t                # inner(grad(phi_a), t)
        M += W[g] * R * np.outer(phi[:, g], phi[:, g]) * detJ
        B += W[g] * np.outer(psi[:, g], dq_dt) * detJ
        L += W[g] * f * psi[:, g] * detJ
    return M, B, L


def assemble_cellwise(net, pbc_vertex, R=1.0, f=0.0):
    """The block system of a (small) network assembled the way DOLFINx does it -- a loop over cells that
    tabulates the element tensors by quadrature (``cell_tensors``) and adds them into a DENSE matrix through
    the cell dof maps, then a loop over the bifurcation / boundary end-vertices for the point terms
    (assembly.py:258-260, 268-277).  ``net`` is an ``OracleNetworkHO`` (or ``OracleNetwork`` with
    ``fd = 1, pd = 0``): only its dof maps are used.  An independent check of the vectorised COO assembly of
    ``reference_port.py`` and of the product's table-driven path (both multiply reference tables by R h)."""
    fd = getattr(net, "fd", 1)
    pd = getattr(net, "pd", 0)
    n = net.n_dofs
    A = np.zeros((n, n))
    b = np.zeros(n)
    qd = net.cell_flux_dofs() if hasattr(net, "cell_flux_dofs") else np.stack(
        [net.fb[np.repeat(np.arange(net.E), net.N)] + np.tile(np.arange(net.N), net.E),
         net.fb[np.repeat(np.arange(net.E), net.N)] + np.tile(np.arange(net.N), net.E) + 1], axis=1)
    pdofs = net.cell_pressure_dofs() if hasattr(net, "cell_pressure_dofs") else (net.poff + np.arange(net.E * net.N))[:, None]
    nc = net.cells.shape[0]
    Rc = np.broadcast_to(np.asarray(R, dtype=np.float64), (nc,))
    fc = np.broadcast_to(np.asarray(f, dtype=np.float64), (nc,))
    for c in range(nc):
        x0, x1 = net.x3[net.cells[c, 0]], net.x3[net.cells[c, 1]]
        M, B, L = cell_tensors(x0, x1, fd, pd, R=float(Rc[c]), f=float(fc[c]), orientation=float(net.orientation[c]))
        for a in range(fd + 1):
            for bb in range(fd + 1):
                A[qd[c, a], qd[c, bb]] += M[a, bb]
            for r in range(pd + 1):
                A[pdofs[c, r], qd[c, a]] += B[r, a]   # a[P][i] = + phi grad(q).t
                A[qd[c, a], pdofs[c, r]] -= B[r, a]   # a[i][P] = - p grad(v).t
        for r in range(pd + 1):
            b[pdofs[c, r]] += L[r]
    N = net.N
    for e, (u, v) in enumerate(net.edges):
        first, last = e * N, e * N + N - 1
        if net.lm_index[v] >= 0:   # edge enters bifurcation v: + mu q and + lam v at the END vertex (local dof 1)
            lm = net.loff + net.lm_index[v]
            A[lm, qd[last, 1]] += 1.0
            A[qd[last, 1], lm] += 1.0
        else:                      # outlet: + p_bc v
            b[qd[last, 1]] += pbc_vertex[v]
        if net.lm_index[u] >= 0:   # edge leaves bifurcation u: - mu q and - lam v at the START vertex (local dof 0)
            lm = net.loff + net.lm_index[u]
            A[lm, qd[first, 0]] -= 1.0
            A[qd[first, 0], lm] -= 1.0
        else:                      # inlet: - p_bc v
            b[qd[first, 0]] -= pbc_vertex[u]
    return A, b
