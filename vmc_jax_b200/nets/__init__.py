from .rbm import CpxRBM, RBM  # noqa: F401
from . import rbm  # noqa: F401
