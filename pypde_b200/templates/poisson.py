"""
Poisson solver plans (pypde/templates/poisson.py:4-110).

2-D: eigen-decomposition along axis 1.  The dense inverse / eigen-decomposition /
projections (CyI, wy, Qy, Hy) are computed on the host with the SAME NumPy /
LAPACK calls as the reference (poisson.py:93-98) - cond(Cy) is ~1e12, so any
other route to these matrices would break parity - and uploaded once.
"""
import numpy as np

from ..field import Field
from ..solver.plans import PlanLHS, PlanRHS
from ..solver.solverplan import SolverPlan
from ..solver.utils import eigdecomp
from .hholtz import _axis_matrices


def solverplan_poisson1d(bases, singular=False):
    field = Field(bases)
    assert field.ndim == 1
    Sx, Bx, Ix = _axis_matrices(field.xs[0])
    Ax = (Ix @ Sx).tolil()
    if singular:
        assert Ax[0, 0] == 0, "Matrix does not look singular"
        Ax[0, 0] += 1e-20
    solver = SolverPlan()
    solver.add_rhs(PlanRHS(Bx, ndim=1, axis=0))
    solver.add_lhs(PlanLHS(Ax.tocsr(), ndim=1, axis=0, method="twodma"))
    return solver


# host-side eigen-decomposition cache: depends only on the y basis, its size and scale (ensembles
# build many solvers of the same grid)
_EIG_CACHE = {}


def solverplan_poisson2d(bases, singular=False, scale=(1, 1)):
    field = Field(bases)
    assert field.ndim == 2
    Sx, Bx, Ix = _axis_matrices(field.xs[0])
    Ax = Ix @ Sx * (1.0 / scale[0] ** 2.0)
    Cx = Bx @ Sx

    key = (field.xs[1].id, field.xs[1].N, float(scale[1]), bool(singular))
    if key not in _EIG_CACHE:
        Sy, By, Iy = _axis_matrices(field.xs[1])
        By = By.toarray()
        Ay = (Iy @ Sy * (1.0 / scale[1] ** 2.0)).toarray()
        Cy = (By @ Sy)
        Cy = np.asarray(Cy)

        CyI = np.linalg.inv(Cy)
        wy, Qy, QyI = eigdecomp(CyI @ Ay)
        if singular:
            wy[0] += 1e-20
        Hy = QyI @ CyI @ By
        _EIG_CACHE[key] = (wy, Qy, Hy)
    wy, Qy, Hy = _EIG_CACHE[key]

    solver = SolverPlan()
    solver.add_rhs(PlanRHS(Bx, ndim=2, axis=0))
    solver.add_rhs(PlanRHS(Hy, ndim=2, axis=1))
    solver.add_lhs(PlanLHS(Ax, alpha=wy, C=Cx, ndim=2, axis=0, method="poisson", singular=True))
    solver.add_lhs(PlanLHS(Qy, ndim=2, axis=1, method="multiply"))
    return solver
