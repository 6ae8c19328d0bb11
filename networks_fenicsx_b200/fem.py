"""``dolfinx.fem`` stand-ins: function spaces with closed-form dof tables and ``Function`` objects
whose ``x.array`` is (pinned) host memory filled straight from the device solution
(solver.py:120-134, post_processing.py:26-51)."""

from __future__ import annotations

import numpy as np
import numpy.typing as npt


class _BasixElement:
    def __init__(self, degree: int, discontinuous: bool, family="Lagrange"):
        self.degree = degree
        self.discontinuous = discontinuous
        self.family = family


class _Element:
    def __init__(self, degree: int, discontinuous: bool):
        self.basix_element = _BasixElement(degree, discontinuous)

    @property
    def interpolation_points(self):
        k = self.basix_element.degree
        if k == 0:
            return np.array([[0.5]])
        return np.concatenate([[0.0, 1.0], np.linspace(0, 1, k + 1)[1:-1]])[:, None]


class _DofMap:
    """Cell -> dof table, built on demand from ``cell_dofs_fn``."""

    def __init__(self, num_dofs: int, cell_dofs_fn):
        from .mesh import IndexMap  # noqa: PLC0415

        self.index_map = IndexMap(num_dofs)
        self.index_map_bs = 1
        self.bs = 1
        self._fn = cell_dofs_fn
        self._list = None

    @property
    def list(self) -> npt.NDArray[np.int32]:
        if self._list is None:
            self._list = np.asarray(self._fn(), dtype=np.int32)
        return self._list

    def cell_dofs(self, c: int) -> npt.NDArray[np.int32]:
        return self.list[c]


class FunctionSpace:
    """Function space on a (sub)mesh.  ``offset`` is the first global row of the space in the
    blocked system ``[q_0 .. q_{C-1}, p, lambda]`` (assembly.py:318-321)."""

    def __init__(self, mesh, degree: int, discontinuous: bool, num_dofs: int, offset: int,
                 cell_dofs_fn, name: str):
        self.mesh = mesh
        self.element = _Element(degree, discontinuous)
        self.dofmap = _DofMap(num_dofs, cell_dofs_fn)
        self.num_dofs = int(num_dofs)
        self.offset = int(offset)
        self.name = name

    def ufl_domain(self):
        return self.mesh

    def tabulate_dof_coordinates(self) -> npt.NDArray[np.float64]:
        raise NotImplementedError


class _Vector:
    """``Function.x``: host array (pinned when allocated by the solver)."""

    def __init__(self, array: np.ndarray):
        self.array = array

    def scatter_forward(self) -> None:
        return None

    @property
    def petsc_vec(self):
        from petsc4py import PETSc  # noqa: PLC0415

        return PETSc.Vec().createWithArray(self.array)


class Function:
    """``dolfinx.fem.Function`` stand-in."""

    def __init__(self, V: FunctionSpace, name: str = "f", array: np.ndarray | None = None):
        self.function_space = V
        self.name = name
        self.x = _Vector(np.zeros(V.num_dofs) if array is None else array)

    @property
    def mesh(self):
        return self.function_space.mesh

    def interpolate(self, u, cells0=None, cells1=None) -> None:
        """Nodal interpolation of a callable ``u(x)`` (x of shape (3, npoints)) at the vertices of a
        P1 space on the parent mesh (assembly.py:225-234)."""
        if callable(u):
            x = self.function_space.mesh.geometry.x
            self.x.array[:] = np.asarray(u(x.T), dtype=np.float64) * np.ones(x.shape[0])
        else:
            raise TypeError("only callables can be interpolated")


class SpatialCoordinate:
    """Minimal ``ufl.SpatialCoordinate`` stand-in for ``p_bc_ex=x[i]``-style expressions
    (demo_Y_bifurcation.py:21-23); components support + - * / with scalars and each other."""

    def __init__(self, mesh):
        self.mesh = mesh

    def __getitem__(self, i: int) -> "CoordinateExpr":
        return CoordinateExpr(lambda x, i=i: x[i])


class CoordinateExpr:
    def __init__(self, fn):
        self._fn = fn

    def __call__(self, x):
        return self._fn(x)

    @staticmethod
    def _lift(o):
        return o if isinstance(o, CoordinateExpr) else CoordinateExpr(lambda x, o=o: o + 0 * x[0])

    def _bin(self, o, op):
        o = self._lift(o)
        return CoordinateExpr(lambda x: op(self(x), o(x)))

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._lift(o)._bin(self, np.add)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._lift(o)._bin(self, np.subtract)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._lift(o)._bin(self, np.multiply)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __rtruediv__(self, o): return self._lift(o)._bin(self, np.divide)
    def __neg__(self): return CoordinateExpr(lambda x: -self(x))
    def __pow__(self, p): return CoordinateExpr(lambda x: self(x) ** p)
