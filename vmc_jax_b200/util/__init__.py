from .tdvp import TDVP, realFun, imagFun, transform_helper  # noqa: F401
from .minsr import MinSR  # noqa: F401
from .stepper import Euler, Heun, AdaptiveHeun  # noqa: F401
from .util import measure, ground_state_search, get_iterable  # noqa: F401
from .output_manager import OutputManager  # noqa: F401
from . import tdvp, minsr, stepper, util, output_manager, symmetries  # noqa: F401
