from . import rbc2d
from .rbc2d import NavierStokes
from .rbc2d_adj import NavierStokesAdjoint
