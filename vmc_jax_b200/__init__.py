"""vmc_jax_b200 -- B200-native implementation of jVMC's per-step VMC hot path behind jVMC's Python API.

    import vmc_jax_b200 as jVMC        # or simply `import jVMC` (alias package at the repo root)

Sub-modules mirror the reference: sampler (MCSampler, ExactSampler), vqs (NQS), operator
(BranchFreeOperator, get_s_primes / get_O_loc), util (TDVP, MinSR, steppers, measure), stats (SampledObs),
mpi_wrapper (global_sum / global_mean ... over NCCL), nets (CpxRBM, RBM), global_defs."""
from . import global_defs  # noqa: F401
from . import mpi_wrapper  # noqa: F401
from . import nets  # noqa: F401
from . import vqs  # noqa: F401
from . import sampler  # noqa: F401
from . import operator  # noqa: F401
from . import stats  # noqa: F401
from . import util  # noqa: F401
from .global_defs import set_pmap_devices  # noqa: F401

__version__ = "0.1.0"
