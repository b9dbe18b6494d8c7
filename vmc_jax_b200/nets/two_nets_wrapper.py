"""Mirror of jVMC/nets/two_nets_wrapper.py:8-31: psi(s) = exp(r(s) + i phi(s)) with an amplitude network r and a
phase network phi, both with real parameters.  Device kernels exist for two real RBMs (the combination the
reference's tests use, tests/vqs_test.py:142-144, tests/sampler_test.py:194-198): both are evaluated by the RBM
kernels, the sampler runs on the amplitude network alone (eval_real, reference :24-26)."""
from .rbm import RBM


class TwoNets:
    def __init__(self, nets):
        nets = tuple(nets)
        if len(nets) != 2 or not all(isinstance(n, RBM) for n in nets):
            raise NotImplementedError("TwoNets has device kernels for a pair of real jVMC.nets.RBM only; got %r" % (nets,))
        self.nets = nets
        self.cpx = False

    def __repr__(self):
        return "TwoNets(%r, %r)" % self.nets

    def init(self, seed, sampleShape, device):
        """Flax names the sub-modules of the tuple field nets_0, nets_1 (leaf order: nets_0/Dense_0/bias, .../kernel,
        nets_1/...)."""
        out = {}
        for i, n in enumerate(self.nets):
            out["nets_%d/Dense_0" % i] = n.init(int(seed) + i, sampleShape, device)["Dense_0"]
        return out
