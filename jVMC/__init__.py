"""Drop-in alias: `import jVMC` resolves to the B200-native implementation (vmc_jax_b200), so scripts
written against markusschmitt/vmc_jax keep their imports (jVMC.sampler.MCSampler, jVMC.vqs.NQS,
jVMC.operator.BranchFreeOperator, jVMC.util.TDVP / MinSR, jVMC.mpi_wrapper ...)."""
import sys

import vmc_jax_b200 as _impl

_self = sys.modules[__name__]
for _k, _v in list(sys.modules.items()):
    if _k == "vmc_jax_b200" or _k.startswith("vmc_jax_b200."):
        sys.modules["jVMC" + _k[len("vmc_jax_b200"):]] = _v
sys.modules["jVMC"] = _impl
