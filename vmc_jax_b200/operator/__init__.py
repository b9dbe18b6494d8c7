from .base import *  # noqa: F401,F403
from .branch_free import *  # noqa: F401,F403
from . import base, branch_free  # noqa: F401
