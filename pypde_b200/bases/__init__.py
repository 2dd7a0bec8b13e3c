from .spectralbase import *
from .spectralbase import Base, MetaBase
from .chebyshev import *
from .chebyshev import Chebyshev, GalerkinChebyshev, ChebDirichlet, ChebNeumann, DirichletC, NeumannC
from .spectralspace import SpectralSpace, SpectralSpaceBC
from .memoize import memoized
from .utils import zero_pad, zero_unpad
