"""
pypde_b200 - B200-native (sm_100a) implementation of pypde's Chebyshev
spectral-Galerkin time-step path, behind pypde's own API.

`from pypde_b200 import *` exports the names reference scripts obtain from
`from pypde import *` (pypde/__init__.py:1-7): np, Base, Field, FieldBC, MultiField,
Integrator, SolverPlan, PlanRHS, PlanLHS, grad, galerkin_to_cheby, cheby_to_galerkin,
conv_term, convective_term, avg_x, avg_vol, interpolate, memoized, initplot, plot.
"""
import numpy as np
import scipy.sparse as sp

from .bases import *
from .bases import (Base, MetaBase, Chebyshev, GalerkinChebyshev, ChebDirichlet, ChebNeumann, DirichletC,
                    NeumannC, SpectralSpace, SpectralSpaceBC, memoized, zero_pad, zero_unpad)
from .solver import *
from .solver import (MetaPlan, PlanRHS, PlanLHS, Plan_fdma, Plan_twodma, Plan_Poisson, Plan_numpy, SolverPlan,
                     Integrator, eigdecomp)
from .field import Field, FieldBC, MultiField, FieldBase
from .field_operations import (grad, cheby_to_galerkin, galerkin_to_cheby, conv_term, convective_term, avg_x,
                               avg_vol, interpolate)

__version__ = "0.1.0"


def initplot(*args, **kwargs):
    """No-op: plotting (pypde/plot) is outside the time-step path."""


def plot(*args, **kwargs):
    raise NotImplementedError("pypde_b200 has no plotting layer; copy fields with .cpu().numpy()")
