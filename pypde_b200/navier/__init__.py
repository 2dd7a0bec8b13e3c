from . import rbc2d
from .rbc2d import NavierStokes
