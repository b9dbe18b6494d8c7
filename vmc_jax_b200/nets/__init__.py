from .rbm import CpxRBM, RBM  # noqa: F401
from . import rbm  # noqa: F401
from . import sym_wrapper  # noqa: F401
from .sym_wrapper import SymNet  # noqa: F401
from . import cnn, activation_functions  # noqa: F401
from .cnn import CNN  # noqa: F401
from . import two_nets_wrapper  # noqa: F401
from .two_nets_wrapper import TwoNets  # noqa: F401
