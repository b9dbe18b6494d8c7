"""Network descriptors mirroring jVMC/nets/rbm.py (CpxRBM :17-38, RBM :67-88).

The reference's nets are Flax modules; here a net is a small descriptor (numHidden, bias) whose
evaluation, gradients and Metropolis updates are CUDA kernels dispatched on the net's type
(no user Python runs inside a kernel).  Parameters live in a dict with Flax's leaf
names: {"Dense_0": {"bias": [M], "kernel": [N, M]}} (sorted-key leaf order: bias, kernel)."""
import numpy as np
import torch


class _RBMBase:
    cpx = True

    def __init__(self, numHidden=2, bias=False):
        self.numHidden = int(numHidden)
        self.bias = bool(bias)

    def __repr__(self):
        return "%s(numHidden=%d, bias=%s)" % (type(self).__name__, self.numHidden, self.bias)

    def init(self, seed, sampleShape, device):
        """Random initial parameters.  CpxRBM: Re, Im ~ U[0, 0.01) (jVMC/nets/initializers.py:17-20),
        RBM: lecun_normal; bias 0.  jax's PRNG stream is not reproduced: numpy default_rng(seed)."""
        N = int(np.prod(sampleShape))
        M = self.numHidden
        rng = np.random.default_rng(int(seed))
        if self.cpx:
            W = rng.uniform(0, 0.01, (N, M)) + 1j * rng.uniform(0, 0.01, (N, M))
            b = np.zeros(M, np.complex128)
        else:
            # lecun_normal: truncated normal (+-2 sigma), variance 1/fan_in
            std = np.sqrt(1.0 / N) / 0.87962566103423978
            W = np.clip(rng.normal(size=(N, M)), -2, 2) * std
            b = np.zeros(M, np.float64)
        p = {"kernel": torch.as_tensor(W).to(device)}
        if self.bias:
            p["bias"] = torch.as_tensor(b).to(device)
        return {"Dense_0": p}


class CpxRBM(_RBMBase):
    """Restricted Boltzmann machine with complex parameters, logpsi = sum_j logcosh((2s-1) W + b)."""
    cpx = True


class RBM(_RBMBase):
    """Restricted Boltzmann machine with real parameters (log(cosh(.)) activation, rbm.py:88)."""
    cpx = False
