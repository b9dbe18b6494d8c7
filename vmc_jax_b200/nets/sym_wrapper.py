"""Mirror of jVMC/nets/sym_wrapper.py: SymNet, the symmetrisation wrapper
Psi(s) = sum_{tau in orbit} factor_tau psi(tau(s))  (reference :22-66, avgFun_Coefficients_Exp :8-10).

On the device the wrapper is supported around CpxRBM / RBM with avgFun_Coefficients_Exp (the default): the orbit
acts on the input, so every orbit element is an RBM with permuted / sign-flipped weight rows and the sampler keeps
all of them incrementally (csrc/symrbm.cu)."""
from .rbm import _RBMBase
from ..util.symmetries import LatticeSymmetry


def avgFun_Coefficients_Exp(coeffs, sym_factor):
    raise NotImplementedError("marker function: the device kernels implement logsumexp(coeffs, b=sym_factor)")


def avgFun_Coefficients_Log(coeffs, sym_factor):
    raise NotImplementedError("avgFun_Coefficients_Log has no device kernel")


def avgFun_Coefficients_Sep(coeffs, sym_factor):
    raise NotImplementedError("avgFun_Coefficients_Sep has no device kernel")


class SymNet:
    """SymNet(orbit, net, avgFun=avgFun_Coefficients_Exp) -- same field names as the reference's Flax module."""

    def __init__(self, orbit=None, net=None, avgFun=avgFun_Coefficients_Exp):
        if not isinstance(orbit, LatticeSymmetry):
            raise TypeError("orbit must be a jVMC.util.symmetries.LatticeSymmetry")
        if not isinstance(net, _RBMBase):
            raise NotImplementedError("SymNet has device kernels only around CpxRBM / RBM; got %r" % (net,))
        if avgFun is not avgFun_Coefficients_Exp:
            raise NotImplementedError("only avgFun_Coefficients_Exp has a device kernel")
        self.orbit = orbit
        self.net = net
        self.avgFun = avgFun

    # the wrapped net's descriptor fields
    @property
    def cpx(self):
        return self.net.cpx

    @property
    def numHidden(self):
        return self.net.numHidden

    @property
    def bias(self):
        return self.net.bias

    def init(self, seed, sampleShape, device):
        return self.net.init(seed, sampleShape, device)

    def __repr__(self):
        return "SymNet(orbit=<%d elements>, net=%r)" % (self.orbit.shape[0], self.net)
