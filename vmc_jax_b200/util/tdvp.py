"""Mirror of jVMC/util/tdvp.py: the TDVP / stochastic-reconfiguration equation and its regularised solve.

Differences in *how* (not in what) it is computed: for (Cpx)RBM states the per-sample gradients stay
factorised (stats.RBMGradientObs); S is assembled by the DMMA Gram kernel at the complex-parameter level
(P_c x P_c instead of the reference's doubled P x P zgemm) and expanded into the reference's layout only
for the eigensolver; the cutoff loop of the pseudo-inverse runs in one kernel without host round trips."""

import os

import numpy as np
import torch

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
        # holomorphic (Cpx)RBM: diagonalise the P_c x P_c complex Gram A instead of the reference's doubled P x P matrix
        # (S0 = A (x) [[1, i], [-i, 1]], so q(S0) has the spectrum of A twice ('real') or +-A ('imag')); set to False to
        # force the reference-layout path
        self.pcLevel = os.environ.get("JVMC_TDVP_PC_LEVEL", "1") != "0"
        self._S = None
        self._S_lazy = None
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
    def S(self):
        """q(S0) (+ diagonal shift) in the reference layout; built on demand after a P_c-level solve."""
        if self._S is None and self._S_lazy is not None:
            self._S = self._S_lazy()
            self._S_lazy = None
        return self._S

    @S.setter
    def S(self, val):
        self._S = val
        self._S_lazy = None

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
            St = S.T.clone(memory_format=torch.contiguous_format)
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
        """eigenvectors of q(S0) as columns (reference layout); after a P_c-level solve the doubled basis is built on
        demand from the eigenvectors of A"""
        if self._Vt is None and getattr(self, "_pc", None) is not None:
            evc, Vtc = self._pc
            G = self._gradObs
            Pc = evc.shape[0]
            if self._mode == 0:
                rows = []
                for part in (0, 1):
                    a, b = (Vtc.real, Vtc.imag) if part == 0 else (-Vtc.imag, Vtc.real)
                    rows.append(torch.stack([self._flat_from_pairs(G, a[k], b[k]) for k in range(Pc)]))
                self._Vt = torch.stack(rows, dim=1).reshape(2 * Pc, 2 * Pc)
            else:
                r2 = 2.0 ** -0.5
                plus = torch.stack([self._flat_from_pairs(G, Vtc[k] * r2, -1j * Vtc[k] * r2) for k in range(Pc)])
                minus = torch.stack([self._flat_from_pairs(G, Vtc[k].conj() * r2, 1j * Vtc[k].conj() * r2) for k in range(Pc)])
                self._Vt = torch.cat([minus.flip(0), plus])
        return self._Vt.T

    def _get_snr(self, Eloc, gradients):
        """reference :173-181."""
        if isinstance(gradients, RBMGradientObs):
            self.rhoVar = gradients.snr_rho_var(Eloc, self._Vt, self.rhsPrefactor, self._mode)
        else:
            EO = gradients.covar_data(Eloc).transform(linearFun=self._Vt.conj(), nonLinearFun=self.trafo_helper)
            self.rhoVar = EO.var().reshape(-1)
        self.snr = torch.sqrt(torch.abs(mpi.globNumSamples * (self.VtF.conj() * self.VtF).real / self.rhoVar)).reshape(-1)

    # ------------------------------------------------------------------ P_c-level solve (holomorphic RBM)
    def _flat_from_pairs(self, G, gpart, igpart):
        """real flat vector in the reference layout (per leaf [g-part, ig-part]) from its two P_c-level halves"""
        Mb = G.M if G.hasBias else 0
        parts = ([gpart[:Mb], igpart[:Mb]] if G.hasBias else []) + [gpart[Mb:], igpart[Mb:]]
        return torch.cat(parts)

    def _solve_pc(self, Eloc, G):
        """TDVP.solve (reference :183-213) with the eigen-decomposition of the complex P_c x P_c Gram matrix A.

        Flat layout index f = (c, part), part 0 = d/dRe, part 1 = d/dIm: S0 = A (x) K, K = [[1, i], [-i, 1]].
        'real': q(S0) is the real representation of A: eigenpairs (lambda_k, [Re v_k; Im v_k]), (lambda_k, [-Im v_k; Re v_k]);
        all projections are Re / Im of zeta = v_k^dagger z with z = -x <conj(dO) dE> at the P_c level.
        'imag': q(S0) = i Im S0 has eigenpairs (+lambda_k, v_k (x) (1,-i)/sqrt2), (-lambda_k, conj(v_k) (x) (1,i)/sqrt2) with
        projections zeta/sqrt2 and -conj(zeta)/sqrt2.  The degenerate pairs of 'real' are regularised with the
        eigenspace-invariant SNR (see below), so the update does not depend on the basis chosen inside a pair."""
        x = self.rhsPrefactor
        self.ElocMean = Eloc.mean()[0]
        self.ElocVar = Eloc.var()[0]
        fk = G.kr_covar_with(Eloc).reshape(-1)
        self.F0 = (-x) * G._kr_to_flat_conj(fk)
        F = self.makeReal(self.F0)
        z = ((-x) * fk).to(torch.complex128)
        self._gradObs = G
        self._S0 = None
        self._St = None
        A = G.gram_A()
        shift = float(self.diagonalShift)
        mode = self._mode
        self.S = None
        self._S_lazy = lambda: (lambda St: St if mode == 0 else St.T)(K.expand_S(A, G.M, G.N, G.hasBias, mode, shift))
        At = A.T.clone(memory_format=torch.contiguous_format)   # private column-major copy: eigh_inplace overwrites it
        if mode == 0 and shift > 1e-10:
            At.diagonal().mul_(1.0 + shift)           # S[f][f] *= 1 + shift acts on diag(Re A) in both halves
        evc, Vtc, info = K.eigh_inplace(At)           # rows of Vtc = eigenvectors v_k
        self._pc = (evc, Vtc)
        self._Vt = None
        zeta = torch.mv(Vtc.conj(), z)
        # per-sample projections zeta_kn = v_k^dagger (-x conj(dO_n) dE_n): weighted first and second moments
        B = G._s.shape[0]
        Pc = evc.shape[0]
        dE = (Eloc._data.reshape(-1) / torch.sqrt(G._p)).to(torch.complex128)
        mu = G.kr_mean().reshape(-1)
        Bglob = int(mpi.globNumSamples) if mpi.globNumSamples else B
        # Both modes only need the first moment sum_n w_n zeta_kn = zeta_k and the second moment sum_n w_n |zeta_kn|^2:
        #   'imag': the two members +-lambda_k carry zeta/sqrt2 and -conj(zeta)/sqrt2: Var rho = (E|zeta|^2 - |E zeta|^2)/2;
        #   'real': the eigenvalue lambda_k of q(S0) is doubly degenerate, the reference's per-eigenvector SNR depends on the
        #           basis LAPACK happens to return inside the 2-d eigenspace.  The eigenspace-invariant
        #           definition is used for both members: SNR^2 = N (|Re zeta|^2 + |Im zeta|^2) / (Var Re zeta + Var Im zeta)
        #           -- it equals the reference's value whenever that is basis independent (equal member SNRs).
        if Pc >= 256 and (Bglob >= 2 * Pc or (Pc >= 8192 and Bglob >= Pc)):
            # sum_n w_n |zeta_kn|^2 = |x|^2 sum_n w'_n |v_k^T (O_n - mu)|^2, w'_n = w_n |dE_n|^2: a second Gram matrix
            # A' = sum w' conj(O) O^T (same tensor-core kernel) and P_c^3 instead of N_s P_c^2 projection flops (large P_c:
            # already from N_s >= P_c on, the second Gram costs a small fraction of either product there)
            w2 = (G._p * (dE.conj() * dE).real).contiguous()
            Ap = G.weighted_second_moment(w2)
            m1 = mpi._all_reduce_sum(K.rbm_moments(G._s, G._tau, w2.to(torch.complex128), G.hasBias, 0)).reshape(-1)
            W2 = mpi._all_reduce_sum(w2.sum().reshape(1))[0]
            q = ((Ap @ Vtc.T) * Vtc.conj().T).sum(0).real             # v_k^dagger A' v_k
            b = torch.mv(Vtc, mu)                                      # v_k^T mu
            a1 = torch.mv(Vtc, m1)                                     # v_k^T sum w' O
            second = abs(x) ** 2 * (q - 2.0 * (b.conj() * a1).real + W2 * (b.conj() * b).real)
            s1 = zeta                                                  # sum_n w_n zeta_kn = v_k^dagger z
            del Ap
        else:
            s1 = torch.zeros(Pc, dtype=torch.complex128, device=A.device)
            second = torch.zeros(Pc, dtype=torch.float64, device=A.device)
            chunk = max(1, min(B, (2 ** 28) // max(Pc, 1)))
            VcH = Vtc.conj().T.contiguous()
            for lo in range(0, B, chunk):
                hi = min(B, lo + chunk)
                g = K.rbm_grad(G._s[lo:hi].contiguous(), G._tau[lo:hi].contiguous(), G.hasBias, 1)     # Khatri-Rao order
                rho = ((-x) * (g - mu[None, :]).conj() * dE[lo:hi, None]) @ VcH
                w = G._p[lo:hi]
                s1 += (w[:, None] * rho).sum(0)
                second += (w[:, None] * (rho.conj() * rho).real).sum(0)
            s1, second = mpi._all_reduce_sum(s1), mpi._all_reduce_sum(second)
        rv = 0.5 * (second - (s1.conj() * s1).real)                    # variance of rho per member of the pair
        if mode == 0:
            self.ev = torch.repeat_interleave(evc, 2)
            self.VtF = torch.stack([zeta.real, zeta.imag], dim=1).reshape(-1).to(torch.complex128)
            self.rhoVar = torch.repeat_interleave(rv, 2)
            snr_pair = torch.sqrt(torch.abs(mpi.globNumSamples * 0.5 * (zeta.conj() * zeta).real / rv))
            self.snr = torch.repeat_interleave(snr_pair, 2)
        else:
            r2 = 2.0 ** -0.5
            self.ev = torch.cat([-evc.flip(0), evc])
            self.VtF = torch.cat([(-zeta.conj() * r2).flip(0), zeta * r2])
            self.rhoVar = torch.cat([rv.flip(0), rv])
            self.snr = torch.sqrt(torch.abs(mpi.globNumSamples * (self.VtF.conj() * self.VtF).real / self.rhoVar)).reshape(-1)
        exact = _is_exact_sampler(self.sampler)
        pinvEv, scal = K.tdvp_regularize(self.ev, self.VtF, None if (exact or self.snrTol == 0) else self.snr, F.to(torch.complex128),
                                         float(self.pinvTol), float(self.pinvCutoff), float(self.snrTol))
        self.invEv = torch.where(torch.abs(self.ev / self.ev[-1]) > 1e-14, 1. / self.ev, torch.zeros_like(self.ev))
        coef = pinvEv * self.VtF
        V = Vtc.T                                       # columns = eigenvectors
        if mode == 0:
            c = coef.reshape(-1, 2)
            u = torch.mv(V, (c[:, 0].real + 1j * c[:, 1].real).to(torch.complex128))
            update = self._flat_from_pairs(G, u.real, u.imag)
        else:
            cm, cp = coef[:Pc].flip(0), coef[Pc:]
            P1 = torch.mv(V, cp)
            P2 = torch.mv(V, cm.conj()).conj()
            r2 = 2.0 ** -0.5
            update = self._flat_from_pairs(G, (P1 + P2).real * r2, (P1.imag - P2.imag) * r2)
        return update, scal[0], scal[1]

    def _solve_capi(self, Eloc, gradients):
        """TDVP.solve (reference :183-213) for dense gradient data in ONE C-ABI call (jvmc_tdvp_solve: eigh, V^dagger F,
        per-sample SNR projections, cutoff loop, update), the moments of rho summed over ranks in-stream."""
        self.S, F = self.get_tdvp_equation(Eloc, gradients)
        S = self.S
        St = S.contiguous().clone() if self._mode == 0 else S.T.clone(memory_format=torch.contiguous_format)
        P = St.shape[0]
        D = gradients._data.reshape(-1, P)
        res = K.tdvp_solve(St, F.to(torch.complex128), D, Eloc._data.reshape(-1), gradients._weights.reshape(-1),
                           self.rhsPrefactor, mpi.globNumSamples, self.snrTol, self.pinvTol, self.pinvCutoff,
                           useSnr=not (_is_exact_sampler(self.sampler) or self.snrTol == 0), comm=mpi.capi_comm())
        self.ev, self._Vt, self.VtF = res["ev"], res["Vt"], res["VtF"]
        self.rhoVar, self.snr = res["rhoVar"], res["snr"]
        self.invEv = torch.where(torch.abs(self.ev / self.ev[-1]) > 1e-14, 1. / self.ev, torch.zeros_like(self.ev))
        return res["update"], res["scal"][0], res["scal"][1]

    def solve(self, Eloc, gradients):
        """reference :183-213."""
        if self.pcLevel and self.diagonalizeOnDevice and isinstance(gradients, RBMGradientObs) and gradients.holomorphic:
            return self._solve_pc(Eloc, gradients)
        if self.diagonalizeOnDevice and not isinstance(gradients, RBMGradientObs) and \
                (mpi.commSize == 1 or mpi.capi_comm() is not None):
            return self._solve_capi(Eloc, gradients)
        self.S, F = self.get_tdvp_equation(Eloc, gradients)
        self._transform_to_eigenbasis(self.S, F)
        exact = _is_exact_sampler(self.sampler)
        self._get_snr(Eloc, gradients)      # computed for every sampler, used unless ExactSampler (:203)
        Fc = F.to(torch.complex128)
        pinvEv, scal = K.tdvp_regularize(self.ev, self.VtF, None if (exact or self.snrTol == 0) else self.snr, Fc, float(self.pinvTol),
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
