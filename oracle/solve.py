"""Oracle: TDVP / SR equation, regularised pseudo-inverse solve, MinSR solve.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  NumPy / LAPACK fp64.
"""
import numpy as np

from .stats import SampledObs


def real_fun(x):
    """jVMC/util/tdvp.py:14-15."""
    return np.real(x)


def imag_fun(x):
    """jVMC/util/tdvp.py:18-19: (x - conj x)/2 = i Im x (purely imaginary complex)."""
    return 0.5 * (x - np.conj(x))


class TDVP:
    """jVMC/util/tdvp.py:74-213 restated on NumPy arrays (single process)."""

    def __init__(self, snrTol=2, pinvTol=1e-14, pinvCutoff=1e-8, makeReal='imag', rhsPrefactor=1.j,
                 diagonalShift=0., exact_sampler=False):
        self.snrTol, self.pinvTol, self.pinvCutoff = snrTol, pinvTol, pinvCutoff
        self.makeReal = imag_fun if makeReal == 'imag' else real_fun
        self.rhsPrefactor = rhsPrefactor
        self.diagonalShift = diagonalShift
        self.exact_sampler = exact_sampler  # tdvp.py:203 isinstance(sampler, ExactSampler)

    def get_tdvp_equation(self, Eloc, grads):
        """tdvp.py:134-148."""
        self.ElocMean = Eloc.mean()[0]
        self.ElocVar = Eloc.var()[0]
        self.F0 = (-self.rhsPrefactor) * grads.covar(Eloc).ravel()
        F = self.makeReal(self.F0)
        self.S0 = grads.covar()
        S = self.makeReal(self.S0)
        if self.diagonalShift > 1e-10:
            S = S + np.diag(self.diagonalShift * np.diag(S))
        return S, F

    def solve(self, Eloc, grads, numSamplesGlobal):
        """tdvp.py:153-213 (eigh -> SNR -> cutoff loop -> update)."""
        self.S, F = self.get_tdvp_equation(Eloc, grads)
        self.ev, self.V = np.linalg.eigh(self.S)
        self.VtF = np.conj(self.V).T @ F
        # _get_snr, tdvp.py:173-181
        pref = self.rhsPrefactor
        EO = grads.covar_data(Eloc).transform(linearFun=np.conj(self.V).T,
                                              nonLinearFun=lambda x: self.makeReal((-pref) * x))
        self.rhoVar = EO.var().ravel()
        with np.errstate(divide='ignore', invalid='ignore'):
            self.snr = np.sqrt(np.abs(numSamplesGlobal * (np.conj(self.VtF) * self.VtF) / self.rhoVar)).ravel()
            relev = np.abs(self.ev / self.ev[-1])
            self.invEv = np.where(relev > 1e-14, 1. / self.ev, 0.)
            residual, cutoff = 1.0, 1e-2
            Fn = np.linalg.norm(F)
            pinvEv = None
            while residual > self.pinvTol and cutoff > self.pinvCutoff:
                cutoff *= 0.8
                reg = 1. / (1. + (max(cutoff, self.pinvCutoff) / relev) ** 6)
                if not self.exact_sampler:
                    reg = reg * (1. / (1. + (self.snrTol / self.snr) ** 6))
                pinvEv = self.invEv * reg
                residual = np.linalg.norm((pinvEv * self.ev - np.ones_like(pinvEv)) * self.VtF) / Fn
        update = np.real(self.V @ (pinvEv * self.VtF))
        return update, residual, max(cutoff, self.pinvCutoff)

    def tdvp_error(self, update):
        """tdvp.py:102-104."""
        return np.abs(1. + np.real(update @ (self.S0 @ update) - 2. * np.real(update @ self.F0)) / self.ElocVar)


def pinv_hermitian(T, rtol):
    """jnp.linalg.pinv(T, rtol=, hermitian=True) (third-party: jax 0.4.12-0.7.2, not vendored):
    eigh-based, eigenvalues with |ev| <= rtol*max|ev| are dropped."""
    ev, V = np.linalg.eigh(T)
    cut = rtol * np.max(np.abs(ev))
    inv = np.where(np.abs(ev) > cut, 1.0 / np.where(ev == 0, 1.0, ev), 0.0)
    return (V * inv[None, :]) @ np.conj(V).T


def minsr_solve(Eloc, grads, holomorphic, pinvTol=1e-14, diagonalShift=0.):
    """MinSR.solve (jVMC/util/minsr.py:53-80)."""
    if holomorphic:
        T = grads.tangent_kernel()
        Tinv = pinv_hermitian(T, pinvTol)
        e = Eloc._data.reshape(-1)
        return -np.conj(grads._data).T @ (Tinv @ e)
    G = np.concatenate([np.real(grads._data), np.imag(grads._data)], axis=0)
    T = G @ G.T + diagonalShift * np.eye(G.shape[0])
    Tinv = pinv_hermitian(T, pinvTol)
    e = Eloc._data.reshape(-1)
    e = np.concatenate([np.real(e), np.imag(e)])
    return -G.T @ (Tinv @ e)


def tdvp_rhs(W, b, s, p, eloc, tdvp, numSamplesGlobal, grad_fn):
    """One right-hand-side evaluation given samples and local energies: the part of
    TDVP.__call__ (tdvp.py:275-284) after sampling / E_loc."""
    E = SampledObs(eloc, p)
    G = SampledObs(grad_fn(s, W, b), p)
    return tdvp.solve(E, G, numSamplesGlobal)
