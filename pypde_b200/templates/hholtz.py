"""
Helmholtz solver plans (pypde/templates/hholtz.py:4-93).

    (1 - lam D2) u = rhs + uold, pre-multiplied by the pseudo-inverse B of D2:
    (B - lam I) S v = B rhs + B S uold

Matrices are assembled from their diagonals with scipy.sparse (every entry is a
sum of at most two products, so the values are bit-identical to the reference's
dense assembly), LU-factored on the host and uploaded.
"""
from ..field import Field
from ..solver.plans import PlanLHS, PlanRHS
from ..solver.solverplan import SolverPlan


def _axis_matrices(base):
    fam = base.family
    return base.S_sp, fam.B_sp(2, 2), fam.I_sp(2)


def solverplan_hholtz1d(bases, lam):
    field = Field(bases)
    assert field.ndim == 1
    Sx, Bx, Ix = _axis_matrices(field.xs[0])
    Ax = Bx @ Sx - lam * Ix @ Sx
    solver = SolverPlan()
    solver.add_rhs(PlanRHS(Bx, ndim=1, axis=0))
    solver.add_old(PlanRHS(Bx @ Sx, ndim=1, axis=0))
    solver.add_lhs(PlanLHS(Ax, ndim=1, axis=0, method="fdma"))
    return solver


def solverplan_hholtz2d_adi(bases, lam, scale=(1, 1)):
    """Alternating-direction factorisation (B - lam I) S per axis; rhs / old / lhs plans are
    appended axis 0 first, then axis 1 (hholtz.py:84-91)."""
    field = Field(bases)
    assert field.ndim == 2
    solver = SolverPlan()
    for axis in (0, 1):
        S, B, I = _axis_matrices(field.xs[axis])
        A = B @ S - lam * (1.0 / scale[axis] ** 2.0) * I @ S
        solver.add_rhs(PlanRHS(B, ndim=2, axis=axis))
        solver.add_old(PlanRHS(B @ S, ndim=2, axis=axis))
        solver.add_lhs(PlanLHS(A, ndim=2, axis=axis, method="fdma"))
    return solver
