"""Oracle: exact basis enumeration and the reference-faithful Metropolis sampler.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The MCMC here is the *reference's algorithm* (full forward pass per proposal,
jVMC/sampler.py:327-356), vectorised over chains with NumPy; its random stream is NumPy's
(the reference's threefry stream is parity-unpinned).
"""
import numpy as np


def basis_states(N, lDim=2, first=0, count=None):
    """ExactSampler.get_basis (jVMC/sampler.py:421-472).  lDim=2: bit i of the integer
    representation is site i (:447); general lDim: most significant digit first (:461-468)."""
    count = lDim ** N - first if count is None else count
    ints = np.arange(first, first + count, dtype=np.int64)
    out = np.zeros((count, N), np.int32)
    if lDim == 2:
        for i in range(N):
            out[:, i] = (ints >> i) & 1
    else:
        c = ints.copy()
        for i in range(N):
            out[:, N - 1 - i] = c % lDim
            c //= lDim
    return out


def exact_probabilities(logpsi, logProbFactor=0.5, lastNorm=0.0):
    """ExactSampler.sample (jVMC/sampler.py:474-526): p = exp(Re(logpsi-lastNorm)/kappa),
    normalised; returns (p, new lastNorm)."""
    p = np.exp(np.real(logpsi - lastNorm) / logProbFactor)
    nrm = np.sum(p)
    return p / nrm, lastNorm + logProbFactor * np.log(nrm)


def distribute_sampling(numSamples, commSize=1, rank=0, localDevices=None, numChainsPerDevice=1):
    """mpi.distribute_sampling (jVMC/mpi_wrapper.py:53-95). Returns (perChain|perProcess, glob)."""
    spp = numSamples // commSize + (1 if rank < numSamples % commSize else 0)
    if localDevices is None:
        return spp, numSamples
    ncp = localDevices * numChainsPerDevice

    def spc(x):
        return int((x + ncp - 1) // ncp)
    a = numSamples % commSize
    glob = (a * spc(1 + numSamples // commSize) + (commSize - a) * spc(numSamples // commSize)) * ncp
    return spc(spp), glob


def propose(rng, states, kind):
    """Proposers jVMC/sampler.py:15-19 (spin_flip), :30-39 (spin_flip_Z2: then a global
    flip with probability 1/5), :42-64 (zeroMag exchange + 1/5 global flip)."""
    C, N = states.shape
    new = states.copy()
    ar = np.arange(C)
    if kind in ("spin_flip", "spin_flip_Z2"):
        idx = rng.integers(0, N, C)
        new[ar, idx] = 1 - new[ar, idx]
    elif kind == "spin_flip_zeroMag":
        bu = rng.integers(1, N // 2 + 1, C)
        bd = rng.integers(1, N // 2 + 1, C)
        cu = np.cumsum(states, axis=1)
        cd = np.cumsum(1 - states, axis=1)
        iu = np.array([np.searchsorted(cu[c], bu[c]) for c in range(C)])
        idn = np.array([np.searchsorted(cd[c], bd[c]) for c in range(C)])
        new[ar, iu] = 0
        new[ar, idn] = 1
    else:
        raise ValueError(kind)
    if kind != "spin_flip":
        flip = rng.integers(0, 5, C) == 0
        new[flip] = 1 - new[flip]
    return new


class MCSampler:
    """Reference-faithful parallel-chain Metropolis (jVMC/sampler.py:264-367), one process."""

    def __init__(self, relogpsi_fn, N, numChains=1, proposer="spin_flip", thermalizationSweeps=10,
                 sweepSteps=10, mu=2, logProbFactor=0.5, seed=0, initState=None):
        if mu < 0 or mu > 2:
            raise ValueError("mu must be in the range [0, 2]")
        self.f = relogpsi_fn
        self.N, self.C, self.kind = N, numChains, proposer
        self.therm, self.K, self.mu, self.kappa = thermalizationSweeps, sweepSteps, mu, logProbFactor
        self.rng = np.random.default_rng(seed)
        self.states = np.zeros((numChains, N), np.int32) if initState is None else \
            np.repeat(np.asarray(initState, np.int32).reshape(1, N), numChains, 0)
        self.numProposed = 0
        self.numAccepted = 0

    def _sweep(self, steps):
        for _ in range(steps):
            new = propose(self.rng, self.states, self.kind)
            newLog = self.mu * self.f(new)
            with np.errstate(over='ignore'):
                P = np.exp(newLog - self.logAcc)
            acc = self.rng.random(self.C) < P      # bernoulli(P), P may exceed 1 (:336-340)
            self.numProposed += self.C
            self.numAccepted += int(acc.sum())
            self.states = np.where(acc[:, None], new, self.states)
            self.logAcc = np.where(acc, newLog, self.logAcc)

    def sample(self, numSamples):
        """Returns configs [spc*C, N] (time-major, chain-minor, :323) and the number generated."""
        self.logAcc = self.mu * self.f(self.states)     # _mc_init :360-367
        self.numProposed = self.numAccepted = 0
        spc, glob = distribute_sampling(numSamples, localDevices=1, numChainsPerDevice=self.C)
        self._sweep(self.therm * self.K)                # thermalise on every call (:309-310)
        out = np.empty((spc, self.C, self.N), np.int32)
        for t in range(spc):
            self._sweep(self.K)
            out[t] = self.states
        return out.reshape(spc * self.C, self.N), glob

    def weights(self, logpsi, globNorm=None):
        """MCSampler.sample tail (jVMC/sampler.py:234-235)."""
        p = np.exp((1.0 / self.kappa - self.mu) * np.real(logpsi))
        return p / (np.sum(p) if globNorm is None else globNorm)
