"""Mirror of jVMC/operator/base.py: the Operator interface (get_s_primes / get_O_loc).

Two device paths:
  * API-parity path: jvmc_bfo_matels + jvmc_bfo_emit reproduce get_s_primes bit-exactly (ordering,
    1e-6 threshold, padding with the last operator's s', reference :91-160); O_loc then needs
    psi(s') on B*Kmax configurations (reference :189) and jvmc_oloc_reduce (:162-164).
  * fused path (get_O_loc with a (Cpx)RBM): jvmc_rbm_eloc_bfo evaluates psi(s')/psi(s) from the cached
    tanh(theta) with per-parameter tables; s' is never materialised."""
import abc
import os

import torch

from .. import global_defs
from .. import kernels as K

opDtype = global_defs.tCpx
# the fused E_loc kernel raises a device flag for strings it cannot evaluate (more than two flipped sites); reading it
# costs one stream sync per get_O_loc call (JVMC_CHECK_DEVICE_FLAGS=0 skips it)
_CHECK_ELOC_FLAG = os.environ.get("JVMC_CHECK_DEVICE_FLAGS", "1") != "0"

__all__ = ["Operator", "opDtype"]


class Operator(metaclass=abc.ABCMeta):
    """Interface for operators defined by their non-zero matrix elements (reference :20-317)."""

    def __init__(self, ElocBatchSize=-1):
        self.compiled = False
        self.compiled_argnum = -1
        self.ElocBatchSize = ElocBatchSize
        self._tables = None

    @abc.abstractmethod
    def compile(self):
        """Returns the host tables of the operator (see BranchFreeOperator.compile)."""

    def _ensure_compiled(self, nargs):
        if (not self.compiled) or self.compiled_argnum != nargs:
            self._tables = self.compile()
            self.compiled = True
            self.compiled_argnum = nargs
        return self._tables

    def get_s_primes(self, s, *args):
        """Connected configurations s' and matrix elements <s|O|s'> (reference :116-160).

        Returns ``(sp int32[dev, B*Kmax, *shape], matEl complex128[dev, B, Kmax])`` and caches
        ``self.sp``, ``self.matEl``, ``self.numNonzero`` like the reference (used by
        get_O_loc_unbatched, :212)."""
        tab = self._ensure_compiled(len(args))
        s = torch.as_tensor(s).to(global_defs.myDevice).to(torch.int32)
        lead, shape = tuple(s.shape[:2]), tuple(s.shape[2:])
        flat = s.reshape(lead[0] * lead[1], -1).contiguous()
        pref = tab.eval_prefactors(*args)
        sp, matEl, cnt = K.bfo_s_primes(flat, tab.device_tables(), pref)
        Kmax = matEl.shape[1]
        self.sp = sp.reshape((lead[0], lead[1] * Kmax) + shape)
        self.matEl = matEl.reshape(lead + (Kmax,))
        self.numNonzero = cnt.reshape(lead)
        return self.sp, self.matEl

    def _get_O_loc(self, matEl, logPsiS, logPsiSP):
        """reference :162-164."""
        lead = tuple(matEl.shape[:2])
        Kmax = matEl.shape[2]
        out = K.oloc_reduce(matEl.reshape(-1, Kmax), logPsiS.reshape(-1), logPsiSP.reshape(-1))
        return out.reshape(lead)

    @staticmethod
    def _eval_deduplicated(psi, sp):
        """psi on the connected configurations, evaluating runs of identical consecutive rows once.  get_s_primes keeps
        the reference's layout (XX and YY strings of a bond give the same s' twice, padding repeats the last s'), so
        for exchange-type operators about half of the forward passes are duplicates."""
        if getattr(psi, "khatri_rao", False) or sp.shape[1] < 2:
            return psi(sp)
        flat = sp.reshape(sp.shape[1], -1)
        uniq, inverse = torch.unique_consecutive(flat, dim=0, return_inverse=True)
        if uniq.shape[0] > 0.75 * flat.shape[0]:
            return psi(sp)
        vals = psi(uniq.reshape((1, uniq.shape[0]) + tuple(sp.shape[2:])))
        return vals.reshape(-1)[inverse].reshape(1, -1)

    def get_O_loc(self, samples, psi, logPsiS=None, *args):
        """O_loc(s) = sum_s' O_{s,s'} psi(s')/psi(s) (reference :166-192)."""
        samples = torch.as_tensor(samples).to(global_defs.myDevice).to(torch.int32)
        if logPsiS is None:
            logPsiS = psi(samples)
        tab = self._ensure_compiled(len(args))
        if tab.fused_ok(psi):
            lead = tuple(samples.shape[:2])
            flat = samples.reshape(lead[0] * lead[1], -1).contiguous()
            out, err = K.rbm_eloc(flat, psi._tau(flat), psi.flip_tables(), tab.device_tables(),
                                  tab.eval_prefactors(*args))
            self._last_err = err
            if _CHECK_ELOC_FLAG and int(err.item()) != 0:
                raise RuntimeError("fused E_loc kernel reported an unsupported operator string (flag %d)" % int(err.item()))
            return out.reshape(lead)
        if tab.cnn_fused_ok(psi):
            lead = tuple(samples.shape[:2])
            flat = samples.reshape(lead[0] * lead[1], -1).contiguous()
            res = K.cnn_eloc(flat, psi.get_parameters(), psi._cnnDesc, tab.device_tables(), tab.eval_prefactors(*args))
            if res is not None:
                out, err = res
                self._last_err = err
                if _CHECK_ELOC_FLAG and int(err.item()) != 0:
                    raise RuntimeError("fused CNN E_loc kernel reported an unsupported operator string (flag %d)" % int(err.item()))
                return out.reshape(lead)
        if self.ElocBatchSize > 0:
            return self.get_O_loc_batched(samples, psi, logPsiS, self.ElocBatchSize, *args)
        sampleOffdConfigs, _ = self.get_s_primes(samples, *args)
        logPsiSP = self._eval_deduplicated(psi, sampleOffdConfigs)
        if not psi.logarithmic:
            logPsiSP = torch.log(logPsiSP)
        return self.get_O_loc_unbatched(logPsiS, logPsiSP)

    def get_O_loc_unbatched(self, logPsiS, logPsiSP):
        """reference :194-212: uses the matrix elements cached by the last get_s_primes call."""
        return self._get_O_loc(self.matEl, logPsiS, logPsiSP)

    def get_O_loc_batched(self, samples, psi, logPsiS, batchSize, *args):
        """reference :214-278 (host loop over slices of the batch axis; same slicing rule)."""
        numSamples = samples.shape[1]
        numBatches = numSamples // batchSize
        remainder = numSamples % batchSize
        if remainder > 0:
            batchSize = numSamples // (numBatches + 1)
            numBatches = numSamples // batchSize
            remainder = numSamples % batchSize
        Oloc = torch.zeros(samples.shape[:2], dtype=torch.complex128, device=samples.device)
        bounds = [(b * batchSize, (b + 1) * batchSize) for b in range(numBatches)]
        if remainder > 0:
            bounds.append((numBatches * batchSize, numSamples))
        for lo, hi in bounds:
            batch = samples[:, lo:hi].contiguous()
            sp, _ = self.get_s_primes(batch, *args)
            logPsiSP = self._eval_deduplicated(psi, sp)
            if not psi.logarithmic:
                logPsiSP = torch.log(logPsiSP)
            Oloc[:, lo:hi] = self.get_O_loc_unbatched(logPsiS[:, lo:hi], logPsiSP)
        return Oloc

    def get_estimator_function(self, psi, *args):
        raise NotImplementedError("get_estimator_function needs autodiff through the net and is outside the "
                                  "B200 hot path (SURVEY 8)")
