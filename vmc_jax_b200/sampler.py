"""Mirror of jVMC/sampler.py: MCSampler (parallel-chain Metropolis) and ExactSampler.

MCSampler runs the whole thermalise + sweep + emit loop of the reference
(_get_samples / _sweep, reference :301-356) as ONE kernel launch (jvmc_rbm_mcmc): one warp per chain,
incremental tanh(theta) updates instead of a full forward pass per proposal, Philox counter RNG keyed by
the global chain id (results do not depend on the number of GPUs).  User-supplied Python proposers
cannot run inside a kernel: the three built-in proposers are dispatched by identity."""
import math

import numpy as np
import torch

from . import global_defs
from . import kernels as K
from . import mpi_wrapper as mpi


class _Proposer:
    """Built-in update proposer, executed on the device inside the sampler kernel."""

    def __init__(self, name, doc):
        self.name = name
        self.kernel_id = K.PROPOSER_IDS[name]
        self.__doc__ = doc
        self.__name__ = "propose_" + name

    def __call__(self, key, s, info):
        raise NotImplementedError("%s runs inside the CUDA sampler kernel and cannot be called from Python"
                                  % self.__name__)

    def __repr__(self):
        return "<device proposer %s>" % self.__name__


propose_spin_flip = _Proposer("spin_flip", "Flip one uniformly chosen spin (reference jVMC/sampler.py:15-19).")
propose_spin_flip_Z2 = _Proposer("spin_flip_Z2", "Single spin flip, then with probability 1/5 a global flip "
                                 "(reference jVMC/sampler.py:30-39).")
propose_spin_flip_zeroMag = _Proposer("spin_flip_zeroMag", "Exchange a random up spin with a random down spin, then "
                                      "with probability 1/5 a global flip (reference jVMC/sampler.py:42-64).")


def _seed_from_key(key):
    if isinstance(key, (int, np.integer)):
        return int(key)
    if isinstance(key, torch.Generator):
        return int(key.initial_seed())
    try:   # jax PRNGKey-like array: use its raw words
        arr = np.asarray(key).astype(np.uint64).ravel()
        out = 0
        for w in arr[-2:]:
            out = (out << 32) | int(w & 0xFFFFFFFF)
        return int(out)
    except Exception:
        raise TypeError("key must be an int seed, a torch.Generator or a PRNG key array")


class MCSampler:
    """Samples from p_mu(s) ~ |psi(s)|^mu with parallel Markov chains (reference :67-380).

    Same constructor as the reference (:103-104)."""

    def __init__(self, net, sampleShape, key, updateProposer=None, numChains=1, updateProposerArg=None,
                 numSamples=100, thermalizationSweeps=10, sweepSteps=10, initState=None, mu=2, logProbFactor=0.5):
        self.sampleShape = tuple(sampleShape)
        self.net = net
        if (not net.is_generator) and (updateProposer is None):
            raise RuntimeError("Instantiation of MCSampler: `updateProposer` is `None` and cannot be used for MCMC "
                               "sampling.")
        if not isinstance(updateProposer, _Proposer):
            raise NotImplementedError("only the built-in proposers propose_spin_flip, propose_spin_flip_Z2 and "
                                      "propose_spin_flip_zeroMag run on the device")
        self.orbit = None
        N = int(np.prod(self.sampleShape))
        dev = global_defs.myDevice
        if initState is None:
            init = torch.zeros(self.sampleShape, dtype=torch.int32, device=dev)
        else:
            init = torch.as_tensor(np.asarray(initState)).to(torch.int32).to(dev).reshape(self.sampleShape)
        self.states = init.reshape(1, 1, *self.sampleShape).repeat(1, numChains, *([1] * len(self.sampleShape))) \
            .contiguous()
        # make sure the net is initialised (reference :124)
        self.net(self.states)
        self.logProbFactor = logProbFactor
        self.mu = mu
        if mu < 0 or mu > 2:
            raise ValueError("mu must be in the range [0, 2]")
        self.updateProposer = updateProposer
        self.updateProposerArg = updateProposerArg
        self.key = _seed_from_key(key)
        self.thermalizationSweeps = thermalizationSweeps
        self.sweepSteps = sweepSteps
        self.numSamples = numSamples
        self.numChains = numChains
        self._stepCounter = 0          # Philox step offset: successive sample() calls continue the stream
        self._counters = torch.zeros(2, dtype=torch.int64, device=dev)
        self.numProposed = self._counters[0:1].reshape(1, 1)
        self.numAccepted = self._counters[1:2].reshape(1, 1)
        self.globNumSamples = 0
        self.refreshEvery = 1

    def set_number_of_samples(self, N):
        self.numSamples = N

    def set_random_key(self, key):
        self.key = _seed_from_key(key)
        self._stepCounter = 0

    def get_last_number_of_samples(self):
        """reference :183-193."""
        return self.globNumSamples

    def sample(self, parameters=None, numSamples=None, multipleOf=1):
        """Returns (configs int32[1,B,*shape], logPsi complex128[1,B], p float64[1,B]) with
        p = exp((1/logProbFactor - mu) Re logPsi) normalised over all ranks (reference :195-235)."""
        if numSamples is None:
            numSamples = self.numSamples
        configs, logPsi = self._get_samples_mcmc(parameters, numSamples, multipleOf)
        expo = 1.0 / self.logProbFactor - self.mu
        if expo == 0.0 and multipleOf == 1:
            # Born sampling: weights are exactly 1/N_glob; tagged so that the Gram kernel can use
            # a scalar weight without inspecting the array
            p = torch.full(logPsi.shape, 1.0 / self.globNumSamples, dtype=torch.float64, device=logPsi.device)
            p._jvmc_uniform = 1.0 / self.globNumSamples
            return configs, logPsi, p
        p = torch.exp(expo * logPsi.real)
        return configs, logPsi, p / mpi.global_sum(p)

    def _get_samples_mcmc(self, params, numSamples, multipleOf=1):
        psi = self.net
        tmpP = None
        if params is not None:
            tmpP = psi.params
            psi.set_parameters(params)
        try:
            isCnn = getattr(psi, "kind", "rbm") == "cnn"
            if not isCnn:
                W, b = psi._cW()
                tables = psi.flip_tables()
            C = self.numChains
            spc, self.globNumSamples = mpi.distribute_sampling(
                numSamples, localDevices=global_defs.device_count(), numChainsPerDevice=math.lcm(C, multipleOf))
            self._counters.zero_()                                    # _mc_init, reference :360-367
            states = self.states.reshape(C, -1)
            K_ = int(self.sweepSteps)
            therm = int(self.thermalizationSweeps) * K_               # thermalise on every call (:309-310)
            if isCnn:
                cfg = K.cnn_mcmc(states, psi.get_parameters(), psi._cnnDesc, self.key, self._stepCounter, mpi.rank * C,
                                 self.updateProposer.kernel_id, float(self.mu), K_, therm, spc, self._counters)
            elif getattr(psi, "sym", None) is not None:
                if self.updateProposer.kernel_id != 0:
                    raise NotImplementedError("SymNet sampling runs on the device with propose_spin_flip only")
                cfg = K.symrbm_mcmc(states, W, b, psi.sym_tables(), tables, self.key, self._stepCounter, mpi.rank * C,
                                    float(self.mu), K_, therm, spc, self._counters, self.refreshEvery)
            else:
                cfg = K.rbm_mcmc(states, W, b, tables, self.key, self._stepCounter, mpi.rank * C,
                                 self.updateProposer.kernel_id, float(self.mu), K_, therm, spc, self._counters,
                                 self.refreshEvery)
            self._stepCounter += therm + spc * K_
            configs = cfg.reshape((1, spc * C) + self.sampleShape)
            coeffs = psi(configs)                                     # re-evaluation, reference :293-296
        finally:
            if tmpP is not None:
                psi.params = tmpP
        return configs, coeffs

    def acceptance_ratio(self):
        """reference :369-380."""
        numProp = mpi.global_sum(self.numProposed.to(torch.float64))
        if float(numProp) > 0:
            return mpi.global_sum(self.numAccepted.to(torch.float64)) / numProp
        return torch.zeros(1, dtype=torch.float64)


class ExactSampler:
    """Full enumeration of the computational basis (reference :385-532)."""

    def __init__(self, net, sampleShape, lDim=2, logProbFactor=0.5):
        self.psi = net
        if isinstance(sampleShape, (int, np.integer)):
            sampleShape = (int(sampleShape),)
        self.sampleShape = tuple(sampleShape)
        self.N = int(np.prod(self.sampleShape))
        self.lDim = lDim
        self.logProbFactor = logProbFactor
        self.get_basis()
        self.psi(self.basis)
        self.lastNorm = 0.

    def get_basis(self):
        """reference :421-472 (bit i of the integer label is site i for lDim = 2)."""
        myNumStates, _ = mpi.distribute_sampling(self.lDim ** self.N)
        first = mpi.first_sample_id()
        dev = global_defs.myDevice
        ints = torch.arange(first, first + myNumStates, dtype=torch.int64, device=dev)
        self.numStatesPerDevice = torch.tensor([myNumStates], device=dev)
        if self.lDim == 2:
            sh = torch.arange(self.N, dtype=torch.int64, device=dev)
            basis = ((ints[:, None] >> sh[None, :]) & 1).to(torch.int32)
        else:
            cols, c = [], ints.clone()
            for _ in range(self.N):
                cols.append((c % self.lDim).to(torch.int32))
                c = c // self.lDim
            basis = torch.stack(cols[::-1], dim=1)
        self.basis = basis.reshape((1, myNumStates) + self.sampleShape).contiguous()

    def sample(self, parameters=None, numSamples=None, multipleOf=None):
        """Returns (basis, logPsi, p) with p = |psi|^(1/logProbFactor) normalised (reference :492-526)."""
        if parameters is not None:
            tmpP = self.psi.get_parameters()
            self.psi.set_parameters(parameters)
        logPsi = self.psi(self.basis)
        if parameters is not None:
            self.psi.set_parameters(tmpP)
        p = torch.exp((logPsi.real - self.lastNorm) / self.logProbFactor)
        nrm = mpi.global_sum(p)
        p = p / nrm
        self.lastNorm = self.lastNorm + self.logProbFactor * float(torch.log(nrm))
        return self.basis, logPsi, p

    def set_number_of_samples(self, N):
        pass

    def get_last_number_of_samples(self):
        return float("inf")
