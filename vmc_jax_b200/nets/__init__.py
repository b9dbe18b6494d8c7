from .rbm import CpxRBM, RBM  # noqa: F401
from . import rbm  # noqa: F401
from . import sym_wrapper  # noqa: F401
from .sym_wrapper import SymNet  # noqa: F401
