"""Mirror of jVMC/util/tdvp.py: the TDVP / stochastic-reconfiguration equation and its regularised solve.

Differences in *how* (not in what) it is computed: for (Cpx)RBM states the per-sample gradients stay
factorised (stats.RBMGradientObs); S is assembled by the DMMA Gram kernel at the complex-parameter level
(P_c x P_c instead of the reference's doubled P x P zgemm) and expanded into the reference's layout only
for the eigensolver; the cutoff loop of the pseudo-inverse runs in one kernel without host round trips."""
import warnings

import numpy as np
import torch

from .. import global_defs
from .. import kernels as K
from .. import mpi_wrapper as mpi
from ..stats import SampledObs, RBMGradientObs


def realFun(x):
    """reference :14-15."""
    return x.real if isinstance(x, torch.Tensor) else np.real(x)


def imagFun(x):
    """reference :18-19: (x - conj x)/2 = i Im x."""
    if isinstance(x, torch.Tensor):
        return 0.5 * (x - x.conj())
    return 0.5 * (x - np.conj(x))


def transform_helper(x, rhsPrefactor, makeReal):
    return makeReal((-rhsPrefactor) * x)


def _is_exact_sampler(sampler):
    from ..sampler import ExactSampler
    return isinstance(sampler, ExactSampler)


class TDVP:
    """Solves q[S] theta_dot = -q[x F] by a regularised pseudo-inverse (reference :26-326).

    Constructor arguments as in the reference (:74)."""

    def __init__(self, sampler, snrTol=2, pinvTol=1e-14, pinvCutoff=1e-8, makeReal='imag', rhsPrefactor=1.j,
                 diagonalShift=0., crossValidation=False, diagonalizeOnDevice=True, svdTol=None):
        self.sampler = sampler
        self.snrTol = snrTol
        self.pinvTol = pinvTol if svdTol is None else svdTol   # stale kwarg of the reference's examples
        self.pinvCutoff = pinvCutoff
        self.diagonalShift = diagonalShift
        self.rhsPrefactor = rhsPrefactor
        self.crossValidation = crossValidation
        self.diagonalizeOnDevice = diagonalizeOnDevice
        self.metaData = None
        self._mode = 1 if makeReal == 'imag' else 0
        self.makeReal = imagFun if makeReal == 'imag' else realFun
        self.trafo_helper = lambda x: transform_helper(x, rhsPrefactor=self.rhsPrefactor, makeReal=self.makeReal)

    def set_diagonal_shift(self, delta):
        self.diagonalShift = delta

    def set_cross_validation(self, crossValidation=True):
        self.crossValidation = crossValidation

    def _get_tdvp_error(self, update):
        """reference :102-104."""
        u = update.to(torch.float64)
        if self._gradObs is not None:
            quad = self._gradObs.quad_form(u)
        else:
            quad = (u.to(self.S0.dtype) @ (self.S0 @ u.to(self.S0.dtype))).real
        lin = (u.to(self.F0.dtype) @ self.F0).real
        return torch.abs(1. + (quad - 2. * lin) / self.ElocVar0)

    def get_residuals(self):
        return self.metaData["tdvp_error"], self.metaData["tdvp_residual"]

    def get_snr(self):
        return self.metaData["SNR"]

    def get_spectrum(self):
        return self.metaData["spectrum"]

    def get_metadata(self):
        return self.metaData

    def get_energy_variance(self):
        return self.ElocVar0

    def get_energy_mean(self):
        return self.ElocMean0.real

    def get_S(self):
        return self.S

    @property
    def S0(self):
        if self._S0 is None and self._gradObs is not None:
            self._S0 = self._gradObs.covar()
        return self._S0

    def get_tdvp_equation(self, Eloc, gradients):
        """reference :134-148.  Returns (S, F); for RBMGradientObs S is produced column-major by the
        expand kernel (self._St) and self.S is its row-major view."""
        self.ElocMean = Eloc.mean()[0]
        self.ElocVar = Eloc.var()[0]
        self.F0 = (-self.rhsPrefactor) * gradients.covar(Eloc).reshape(-1)
        F = self.makeReal(self.F0)
        self._S0 = None
        self._gradObs = None
        if isinstance(gradients, RBMGradientObs) and gradients.holomorphic:
            self._gradObs = gradients
            A = gradients.gram_A()
            self._St = K.expand_S(A, gradients.M, gradients.N, gradients.hasBias, self._mode, float(self.diagonalShift))
            S = self._St if self._mode == 0 else self._St.T      # real-symmetric: S^T = S
            return S, F
        self._S0 = gradients.covar()
        S = self.makeReal(self._S0)
        if self.diagonalShift > 1e-10:
            S = S + torch.diag(self.diagonalShift * torch.diag(S))
        self._St = None
        return S, F

    def _transform_to_eigenbasis(self, S, F):
        """reference :153-171.  cuSOLVER syevd/heevd through the C ABI; self.V has eigenvectors as columns."""
        if self._St is not None:
            St = self._St.clone()
        else:
            St = S.T.contiguous().clone()
        if not self.diagonalizeOnDevice:
            ev, V = np.linalg.eigh(S.cpu().numpy())
            self.ev = torch.as_tensor(ev).to(S.device)
            self._Vt = torch.as_tensor(np.ascontiguousarray(V.T)).to(S.device)
        else:
            self.ev, self._Vt, info = K.eigh_inplace(St)
        Fc = F.to(torch.complex128)
        self.VtF = torch.mv(self._Vt.conj().to(torch.complex128), Fc)

    @property
    def V(self):
        return self._Vt.T

    def _get_snr(self, Eloc, gradients):
        """reference :173-181."""
        if isinstance(gradients, RBMGradientObs):
            self.rhoVar = gradients.snr_rho_var(Eloc, self._Vt, self.rhsPrefactor, self._mode)
        else:
            EO = gradients.covar_data(Eloc).transform(linearFun=self._Vt.conj(), nonLinearFun=self.trafo_helper)
            self.rhoVar = EO.var().reshape(-1)
        self.snr = torch.sqrt(torch.abs(mpi.globNumSamples * (self.VtF.conj() * self.VtF).real / self.rhoVar)).reshape(-1)

    def solve(self, Eloc, gradients):
        """reference :183-213."""
        self.S, F = self.get_tdvp_equation(Eloc, gradients)
        self._transform_to_eigenbasis(self.S, F)
        exact = _is_exact_sampler(self.sampler)
        self._get_snr(Eloc, gradients)      # computed for every sampler, used unless ExactSampler (:203)
        Fc = F.to(torch.complex128)
        pinvEv, scal = K.tdvp_regularize(self.ev, self.VtF, None if exact else self.snr, Fc, float(self.pinvTol),
                                         float(self.pinvCutoff), float(self.snrTol))
        self.invEv = torch.where(torch.abs(self.ev / self.ev[-1]) > 1e-14, 1. / self.ev, torch.zeros_like(self.ev))
        update = torch.mv(self._Vt.to(torch.complex128).T, (pinvEv * self.VtF)).real
        residual, cutoff = scal[0], scal[1]
        return update, residual, cutoff

    def S_dot(self, v):
        return self.S0 @ torch.as_tensor(v).to(self.S0.device).to(self.S0.dtype)

    def __call__(self, netParameters, t, *, psi, hamiltonian, **rhsArgs):
        """Right-hand side theta_dot = S^-1 F for the ODE steppers (reference :219-326)."""
        tmpParameters = psi.get_parameters()
        psi.set_parameters(netParameters)
        outp = rhsArgs.get("outp", None)
        self.outp = outp
        numSamples = rhsArgs.get("numSamples", None)

        def start_timing(name):
            if outp is not None:
                outp.start_timing(name)

        def stop_timing(name):
            if outp is not None:
                torch.cuda.synchronize() if torch.cuda.is_available() else None
                outp.stop_timing(name)

        start_timing("sampling")
        sampleConfigs, sampleLogPsi, p = self.sampler.sample(numSamples=numSamples)
        stop_timing("sampling")

        start_timing("compute Eloc")
        Eloc = hamiltonian.get_O_loc(sampleConfigs, psi, sampleLogPsi, t)
        stop_timing("compute Eloc")
        Eloc = SampledObs(Eloc, p)

        start_timing("compute gradients")
        if getattr(psi, "khatri_rao", False):
            sampleGradients = RBMGradientObs(psi, sampleConfigs, p)      # factorised: O is never formed
        else:
            sampleGradients = SampledObs(psi.gradients(sampleConfigs), p)
        stop_timing("compute gradients")

        start_timing("solve TDVP eqn.")
        update, solverResidual, pinvCutoff = self.solve(Eloc, sampleGradients)
        stop_timing("solve TDVP eqn.")

        if outp is not None:
            outp.add_timing("MPI communication", mpi.get_communication_time())

        psi.set_parameters(tmpParameters)

        if "intStep" in rhsArgs and rhsArgs["intStep"] == 0:
            self.ElocMean0 = self.ElocMean
            self.ElocVar0 = self.ElocVar
            self.metaData = {
                "tdvp_error": self._get_tdvp_error(update),
                "tdvp_residual": solverResidual,
                "pinv_cutoff": pinvCutoff,
                "SNR": self.snr,
                "spectrum": self.ev,
            }
            if self.crossValidation:
                Eloc1 = Eloc.subset(start=0, step=2)
                sampleGradients1 = sampleGradients.subset(start=0, step=2)
                Eloc2 = Eloc.subset(start=1, step=2)
                sampleGradients2 = sampleGradients.subset(start=1, step=2)
                update_1, _, _ = self.solve(Eloc1, sampleGradients1)
                S2, F2 = self.get_tdvp_equation(Eloc2, sampleGradients2)
                u1 = update_1.to(S2.dtype)
                validation_tdvpErr = self._get_tdvp_error(update_1)
                num = torch.linalg.norm(S2 @ u1 - F2.to(S2.dtype)) / torch.linalg.norm(F2)
                update, solverResidual, _ = self.solve(Eloc, sampleGradients)
                validation_residual = num / solverResidual
                self.crossValidationFactor_residual = validation_residual
                self.crossValidationFactor_tdvpErr = validation_tdvpErr / self.metaData["tdvp_error"]
                self.metaData["tdvp_residual_cross_validation_ratio"] = self.crossValidationFactor_residual
                self.metaData["tdvp_error_cross_validation_ratio"] = self.crossValidationFactor_tdvpErr
                self.S, _ = self.get_tdvp_equation(Eloc, sampleGradients)
        return update
