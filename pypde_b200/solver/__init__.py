from .plans import *
from .plans import MetaPlan, PlanRHS, PlanLHS, Plan_fdma, Plan_twodma, Plan_Poisson, Plan_numpy
from .solverplan import SolverPlan
from .integrator import Integrator
from .utils import eigdecomp
