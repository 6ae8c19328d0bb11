"""Pins the oracle against numbers produced by the UNMODIFIED reference (DOLFINx/PETSc/MUMPS), when
``tests/golden/reference_solution.npz`` exists (written by tests/golden/make_reference_solution.py inside the
reference's CI container -- it cannot be produced in this repository's build container).  Until then the
test skips and the assembled values / solutions remain "parity unpinned by the reference"."""

import pathlib

import numpy as np
import pytest

from oracle import reference_port as rp
from tests import helpers

FIXTURE = pathlib.Path(__file__).parent / "golden" / "reference_solution.npz"


@pytest.mark.skipif(not FIXTURE.exists(), reason="reference_solution.npz has not been produced yet (needs DOLFINx)")
@pytest.mark.parametrize("case", ["y", "double_y", "tree", "arterial"])
def test_oracle_equals_the_reference(case):
    d = np.load(FIXTURE)
    p_bc = (lambda x: x[0]) if case == "double_y" else (lambda x: x[1])
    net = rp.OracleNetwork(d[f"{case}/pos"], d[f"{case}/edges"], d[f"{case}/colors"].astype(np.int32), int(d[f"{case}/N"]))
    A, b = net.assemble(net.eval_pbc(p_bc))
    assert np.array_equal(A.indptr, d[f"{case}/indptr"]) and np.array_equal(A.indices, d[f"{case}/indices"]), \
        "sparsity pattern (explicit zeros included) differs from the reference's"
    np.testing.assert_allclose(A.data, d[f"{case}/values"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(b, d[f"{case}/b"], rtol=1e-12, atol=1e-14)
    assert helpers.rel_l2(net.solve(A, b), d[f"{case}/x"]) < 1e-8
