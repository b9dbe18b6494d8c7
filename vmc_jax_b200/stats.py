"""Mirror of jVMC/stats.py: SampledObs statistics on device arrays.

``SampledObs`` is the generic container (any observable, arrays [dev, B, ...]); its reductions are
plain device tensor operations followed by the in-stream all-reduce of mpi_wrapper.
``RBMGradientObs`` is the B200-native specialisation for the per-sample gradients of a (Cpx)RBM: it
stores only (configs, tanh(theta), weights) -- the Khatri-Rao factors of O -- and evaluates mean, force,
S (Gram), SNR projections and the MinSR contraction with the kernels of csrc/stats.cu and csrc/gram.cu
without ever materialising O [N_s x P]."""
import os

import numpy as np
import torch

from . import global_defs
from . import kernels as K
from . import mpi_wrapper as mpi


GRAM_BACKEND = os.environ.get("JVMC_GRAM_BACKEND", "i8")


def _gram_backend():
    return {"i8": K.rbm_gram_S_auto, "i8only": K.rbm_gram_S_i8, "dmma": K.rbm_gram_S}[GRAM_BACKEND]


def _w_like(w, data):
    return w.reshape(w.shape + (1,) * (data.dim() - 2))


class SampledObs:
    """Statistics of a sampled observable (reference jVMC/stats.py:136-336).

    Args: ``observations`` [dev, B, ...], ``weights`` [dev, B] (normalised over all ranks).
    ``estimator`` / ``params`` (autodiff estimators) are not supported on the B200 path."""

    def __init__(self, observations=None, weights=None, estimator=None, params=None):
        if estimator is not None:
            raise NotImplementedError("estimator functions need autodiff and are outside the B200 hot path")
        self._weights = None if weights is None else torch.as_tensor(weights).to(global_defs.myDevice)
        self._data = None
        self._mean = None
        self._configs = None

        def estimator_not_implemented(p, s):
            raise Exception("No estimator function given.")
        self._estimator = estimator_not_implemented
        self._estimator_grad = estimator_not_implemented
        if observations is not None:
            observations = torch.as_tensor(observations).to(global_defs.myDevice)
        self._compute_data_and_mean(observations)

    def _compute_data_and_mean(self, observations):
        """reference :197-205: mean = sum_n w_n O_n, data = sqrt(w_n) (O_n - mean)."""
        if (observations is not None) and (self._weights is not None):
            if observations.dim() == 2:
                observations = observations[..., None]
            if not observations.is_floating_point() and not observations.is_complex():
                observations = observations.to(torch.float64)
            w = _w_like(self._weights, observations)
            self._mean = mpi.global_sum((w.to(observations.dtype) * observations))
            self._data = torch.sqrt(w) * (observations - self._mean)

    def mean(self, params=None):
        if params is not None:
            raise NotImplementedError("estimator functions are not supported")
        return self._mean

    def covar(self, other=None):
        """reference :235-245: sum_n conj(data1_n) (x) data2_n."""
        if other is None:
            other = self
        if isinstance(other, RBMGradientObs) and not isinstance(self, RBMGradientObs):
            return other.covar(self).conj().T
        d1 = self._data.reshape(-1, int(np.prod(self._data.shape[2:])))
        d2 = other._data.reshape(-1, int(np.prod(other._data.shape[2:])))
        if d1.dtype != d2.dtype:
            ct = torch.promote_types(d1.dtype, d2.dtype)
            d1, d2 = d1.to(ct), d2.to(ct)
        return mpi._all_reduce_sum(d1.conj().T @ d2)

    def var(self):
        """reference :248-252."""
        return mpi.global_sum((self._data.conj() * self._data).real)

    def covar_data(self, other=None):
        """reference :255-265: per-sample outer(conj(d1), d2)/w as a new SampledObs."""
        if other is None:
            other = self
        d1 = self._data.reshape(self._data.shape[:2] + (-1,))
        d2 = other._data.reshape(other._data.shape[:2] + (-1,))
        obs = d1.conj()[..., :, None] * d2[..., None, :] / self._weights[..., None, None]
        return SampledObs(obs, self._weights)

    def covar_var(self, other=None):
        """reference :268-279."""
        if other is None:
            other = self
        d1 = self._data.reshape(self._data.shape[:2] + (-1,))
        d2 = other._data.reshape(other._data.shape[:2] + (-1,))
        outer = d1.conj()[..., :, None] * d2[..., None, :]
        return mpi.global_sum(outer.abs() ** 2 / self._weights[..., None, None]) - self.covar(other).abs() ** 2

    def transform(self, nonLinearFun=lambda x: x, linearFun=None):
        """reference :282-292: f(data/sqrt(w) + mean), optionally followed by a matrix product."""
        x = nonLinearFun(self._data / torch.sqrt(_w_like(self._weights, self._data)) + self._mean)
        if linearFun is not None:
            lf = torch.as_tensor(linearFun).to(x.device)
            ct = torch.promote_types(lf.dtype, x.dtype)
            x = torch.matmul(lf.to(ct), x.to(ct))
        return SampledObs(x, self._weights)

    def select(self, ixs):
        """reference :295-307."""
        newObs = SampledObs()
        newObs._data = self._data[:, :, ixs]
        newObs._mean = self._mean[ixs]
        newObs._weights = self._weights
        return newObs

    def subset(self, start=None, end=None, step=None):
        """reference :310-329."""
        sl = slice(start, end, step)
        newObs = SampledObs()
        w = self._weights[:, sl]
        normalization = mpi.global_sum(w)
        d = self._data[:, sl] / torch.sqrt(normalization)
        w = w / normalization
        newObs._weights = w
        newObs._mean = mpi.global_sum(torch.sqrt(_w_like(w, d)) * d) + self._mean
        newObs._data = d + torch.sqrt(_w_like(w, d)) * (self._mean - newObs._mean)
        return newObs

    def tangent_kernel(self):
        """reference :332-336."""
        all_data = mpi.gather(self._data)
        all_data = all_data.reshape(all_data.shape[0], -1)
        return all_data @ all_data.conj().T


class RBMGradientObs(SampledObs):
    """Per-sample (Cpx)RBM gradients as Khatri-Rao factors: O_n[(r, j)] = sigma_{n,r} tau_{n,j}.

    Holds configs int32[B, N], tau complex128[B, M], weights float64[B]; everything the TDVP / MinSR
    equations need is computed from these by the fused kernels."""

    def __init__(self, psi, configs, weights):
        SampledObs.__init__(self)
        s = torch.as_tensor(configs).to(global_defs.myDevice).to(torch.int32)
        self.psi = psi
        self._lead = tuple(s.shape[:2])
        self._s = s.reshape(self._lead[0] * self._lead[1], -1).contiguous()
        self._tau = psi._tau(self._s)
        self._uniform = getattr(weights, "_jvmc_uniform", None)
        self._weights = torch.as_tensor(weights).to(global_defs.myDevice)
        self._p = self._weights.reshape(-1).contiguous()
        self.N, self.M = self._s.shape[1], self._tau.shape[1]
        self.hasBias = psi.b is not None
        self.holomorphic = psi.holomorphic
        self.R = self.N + (1 if self.hasBias else 0)
        self._mu = None          # [R, M] Khatri-Rao mean (global)
        self._A = None
        self._sigT = None
        self._dense = None

    # ---- Khatri-Rao <-> reference flat layout
    def _kr_to_flat(self, v):
        """complex vector in c = r*M + j order -> reference flat layout (holomorphic: per leaf [g, i g])."""
        v = v.reshape(-1)
        if not self.holomorphic:
            return v
        Mb = self.M if self.hasBias else 0
        parts = []
        if self.hasBias:
            parts += [v[:Mb], 1j * v[:Mb]]
        parts += [v[Mb:], 1j * v[Mb:]]
        return torch.cat(parts)

    def _kr_to_flat_conj(self, v):
        """covar(grads, E) layout: sum conj(dO) dE in c = r*M + j order -> flat (holomorphic: per leaf [f, -i f])."""
        v = v.reshape(-1)
        if not self.holomorphic:
            return v
        Mb = self.M if self.hasBias else 0
        parts = ([v[:Mb], -1j * v[:Mb]] if self.hasBias else []) + [v[Mb:], -1j * v[Mb:]]
        return torch.cat(parts)

    def kr_mean(self):
        if self._mu is None:
            mu = K.rbm_moments(self._s, self._tau, self._p.to(torch.complex128), self.hasBias, 0)
            self._mu = mpi._all_reduce_sum(mu)
        return self._mu

    def mean(self, params=None):
        if self._mean is None:
            self._mean = self._kr_to_flat(self.kr_mean())
        return self._mean

    @property
    def _data(self):
        """Dense centred data sqrt(w)(O - mean) [dev, B, P] -- only for small problems / generic callers."""
        if self._dense is None and self._s is not None:
            B = self._s.shape[0]
            P = (2 if self.holomorphic else 1) * self.R * self.M
            if B * P * 16 > 8 * 2 ** 30:
                raise MemoryError("dense gradient matrix would need %.1f GB; use the implicit RBMGradientObs methods"
                                  % (B * P * 16 / 2 ** 30))
            g = K.rbm_grad(self._s, self._tau, self.hasBias, 0 if self.holomorphic else 1)
            d = torch.sqrt(self._p)[:, None] * (g - self.mean()[None, :])
            self._dense = d.reshape(self._lead + (P,))
        return self._dense

    @_data.setter
    def _data(self, val):
        self._dense = val

    def kr_covar_with(self, other):
        """Khatri-Rao order vector sum_n conj(sqrt(w) (O_n - mu)) d_n for a scalar observable ``other``."""
        d = other._data.reshape(-1).to(torch.complex128)
        wgt = torch.sqrt(self._p) * d
        Fk = K.rbm_moments(self._s, self._tau, wgt, self.hasBias, 1)
        Fk = Fk - self.kr_mean().conj() * wgt.sum()
        return mpi._all_reduce_sum(Fk)

    def covar(self, other=None):
        if other is None:
            return self._expand_S0()
        if isinstance(other, RBMGradientObs):
            return SampledObs.covar(self, other)
        nd = int(np.prod(other._data.shape[2:]))
        if nd != 1:
            return SampledObs.covar(self, other)
        Fk = self.kr_covar_with(other)
        if self.holomorphic:
            Mb = self.M if self.hasBias else 0
            parts = []
            if self.hasBias:
                parts += [Fk.reshape(-1)[:Mb], -1j * Fk.reshape(-1)[:Mb]]
            parts += [Fk.reshape(-1)[Mb:], -1j * Fk.reshape(-1)[Mb:]]
            return torch.cat(parts)[:, None]
        return Fk.reshape(-1)[:, None]

    def _gram_over_ranks(self, Y, mu, alpha, kappa):
        """The Hermitian Gram matrix of this rank's samples, summed over ranks.  With several ranks and the int8 kernel
        the site pairs are launched in groups of block rows (r0 ascending): as soon as the launches of a group are
        enqueued its rows of the upper block triangle are final on this rank, and a side stream packs, all-reduces
        (NCCL) and unpacks them while the next group is computed; the lower triangle is rebuilt by conjugate
        transposition at the end.  Replaces the host-staged MPI.Allreduce of S inside covar()
        (reference jVMC/mpi_wrapper.py:114-147 via stats.py:245)."""
        gram = _gram_backend()
        kw = {"s": self._s, "hasBias": self.hasBias} if gram is K.rbm_gram_S_auto else {}
        overlap = mpi.commSize > 1 and gram is not K.rbm_gram_S and os.environ.get("JVMC_GRAM_OVERLAP", "1") != "0"
        if not overlap:
            A = gram(Y, self._sigT, mu, alpha, kappa, **kw)
            # Hermitian: from 8 ranks on communicate the upper block triangle only (half the bytes) and rebuild the
            # rest; below that the pack / unpack / mirror passes cost more than the saved NVLink time (measured)
            half = os.environ.get("JVMC_HERMITIAN_ALLREDUCE")
            half = (mpi.commSize >= 8) if half is None else (half == "1")
            return mpi.all_reduce_hermitian_blocks(A, self.M) if half else mpi._all_reduce_sum(A)
        import torch.distributed as dist
        R, M = self.R, self.M
        Pc = R * M
        # block-row groups with shrinking numbers of pairs: the reduction of the last group is the only exposed one
        fr, bounds, acc = (0.4, 0.7, 0.9, 1.0), [0], 0
        total = R * (R + 1) // 2
        for f in fr:
            r = bounds[-1]
            while r < R and K.pairs_before_row(R, r) < f * total:
                r += 1
            if r > bounds[-1]:
                bounds.append(r)
        if bounds[-1] != R:
            bounds.append(R)
        groups = [(K.pairs_before_row(R, a), K.pairs_before_row(R, b) - K.pairs_before_row(R, a))
                  for a, b in zip(bounds[:-1], bounds[1:])]
        packed = torch.empty(K.hermitian_packed_offset(Pc, M, R), dtype=torch.complex128, device=Y.device)
        main = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=Y.device, priority=-1)
        holder = {}

        def after_group(k):
            A_ = holder["A"]
            ev = torch.cuda.Event()
            ev.record(main)
            ra, rb = bounds[k], bounds[k + 1]
            with torch.cuda.stream(side):
                side.wait_event(ev)
                K.hermitian_pack_rows(A_, M, ra * M, (rb - ra) * M, packed)
                sl = packed[K.hermitian_packed_offset(Pc, M, ra):K.hermitian_packed_offset(Pc, M, rb)]
                dist.all_reduce(torch.view_as_real(sl), op=dist.ReduceOp.SUM)
                K.hermitian_pack_rows(A_, M, ra * M, (rb - ra) * M, packed, unpack=True)
        A = torch.empty((Pc, Pc), dtype=torch.complex128, device=Y.device)
        holder["A"] = A
        A = gram(Y, self._sigT, mu, alpha, kappa, A, pair_groups=groups, after_group=after_group, **kw)
        if not K.LAST_GRAM.get("grouped", False) and gram is K.rbm_gram_S_auto:
            # this rank sent the whole matrix through the fp64 kernel (no groups ran).  The backend choice is rank-local
            # (it depends on the rank's samples), the sequence of collectives must not be: reduce the same row groups
            for k in range(len(groups)):
                after_group(k)
        main.wait_stream(side)
        K.hermitian_mirror_blocks(A, M)
        return A

    def gram_A(self):
        """A[(r,j),(r',l)] = <conj(O) O>_c, complex Hermitian [R*M, R*M] (global)."""
        if self._A is None:
            mu = self.kr_mean()
            if self._sigT is None:
                self._sigT = K.pack_sigma(self._s, self.hasBias)
            p = self._p
            kappa = 1.0 / mpi.commSize
            # backend: "i8" = tcgen05 INT8 tensor cores with error-free splitting (fp64-equivalent, default; outlier
            #          samples of heavy-tailed tau columns go through the fp64 kernel, kernels.rbm_gram_S_auto),
            #          "i8only", "dmma" = fp64 DMMA
            if self._uniform is not None:
                self._A = self._gram_over_ranks(self._tau, mu, float(self._uniform), kappa)
            else:
                self._A = self._gram_over_ranks(self._tau * torch.sqrt(p)[:, None], mu, 1.0, kappa)
        return self._A

    def weighted_second_moment(self, w2):
        """sum_n w2_n conj(O_n) O_n^T in Khatri-Rao order (uncentred, Hermitian [R*M, R*M], global) for real weights
        w2 >= 0 -- the same Gram kernel as gram_A with another weight vector (SNR second moments, util/tdvp.py)."""
        if self._sigT is None:
            self._sigT = K.pack_sigma(self._s, self.hasBias)
        Y = self._tau * torch.sqrt(w2.to(torch.float64))[:, None]
        return self._gram_over_ranks(Y, torch.zeros((self.R, self.M), dtype=torch.complex128, device=Y.device), 1.0, 0.0)

    def _expand_S0(self):
        A = self.gram_A()
        if not self.holomorphic:
            return A
        Mb = self.M if self.hasBias else 0
        Pc = A.shape[0]
        idx = []
        for lo, hi in ([(0, Mb)] if self.hasBias else []) + [(Mb, Pc)]:
            r = torch.arange(lo, hi, device=A.device)
            idx += [(r, 1.0), (r, 1j)]
        rows = torch.cat([i for i, _ in idx])
        ph = torch.cat([torch.full((len(i),), f, dtype=torch.complex128, device=A.device) for i, f in idx])
        return ph.conj()[:, None] * A[rows][:, rows] * ph[None, :]

    def var(self):
        """sum_n w_n |O_n - mean|^2 per flat component: |sigma tau|^2 = |tau|^2 for every site."""
        second = mpi._all_reduce_sum((self._p[:, None] * (self._tau.conj() * self._tau).real).sum(0))
        v = second[None, :].expand(self.R, self.M).reshape(-1) - (self.kr_mean().conj() * self.kr_mean()).real.reshape(-1)
        if not self.holomorphic:
            return v
        Mb = self.M if self.hasBias else 0
        parts = ([v[:Mb], v[:Mb]] if self.hasBias else []) + [v[Mb:], v[Mb:]]
        return torch.cat(parts)

    def quad_form(self, u):
        """u^T S0 u for a real flat vector u (TDVP error, reference jVMC/util/tdvp.py:102-104) from A."""
        u = torch.as_tensor(u).to(self._tau.device)
        if not self.holomorphic:
            z = u.to(torch.complex128)
        else:
            Mb = self.M if self.hasBias else 0
            NM = self.N * self.M
            zs = []
            off = 0
            for n in ([Mb] if self.hasBias else []) + [NM]:
                zs.append(u[off:off + n] + 1j * u[off + n:off + 2 * n])
                off += 2 * n
            z = torch.cat(zs).to(torch.complex128)
        return (z.conj() @ (self.gram_A() @ z)).real

    def subset(self, start=None, end=None, step=None):
        """reference :310-329 -- equals a fresh observable on the slice with renormalised weights."""
        sl = slice(start, end, step)
        w = self._weights[:, sl]
        w = w / mpi.global_sum(w)
        cfg = self._s.reshape(self._lead + (self.N,))[:, sl].contiguous()
        return RBMGradientObs(self.psi, cfg, w.contiguous())

    def snr_rho_var(self, Eloc, Vt, prefactor, mode):
        """Variance over samples of rho_n = V^dagger q(-x conj(dO_n) dE_n)  (reference
        jVMC/util/tdvp.py:173-181 via stats.py covar_data/transform/var), chunked over samples so that
        only [chunk, P] rows of O exist at a time.  Vt: rows = eigenvectors."""
        B = self._s.shape[0]
        P = Vt.shape[0]
        dE = (Eloc._data.reshape(-1) / torch.sqrt(self._p)).to(torch.complex128)
        mean = self.mean()
        s1 = torch.zeros(P, dtype=torch.complex128, device=Vt.device)
        s2 = torch.zeros(P, dtype=torch.float64, device=Vt.device)
        chunk = max(1, min(B, (2 ** 28) // max(P, 1)))
        VtH = Vt.conj().to(torch.complex128)
        for lo in range(0, B, chunk):
            hi = min(B, lo + chunk)
            g = K.rbm_grad(self._s[lo:hi].contiguous(), self._tau[lo:hi].contiguous(), self.hasBias,
                           0 if self.holomorphic else 1)
            x = (-prefactor) * (g - mean[None, :]).conj() * dE[lo:hi, None]
            if mode == 0 and not Vt.is_complex():
                rho = (x.real @ Vt.T).to(torch.complex128)
            else:
                x = x.real.to(torch.complex128) if mode == 0 else 1j * x.imag
                rho = x @ VtH.T
            w = self._p[lo:hi]
            s1 += (w[:, None] * rho).sum(0)
            s2 += (w[:, None] * (rho.conj() * rho).real).sum(0)
        s1 = mpi._all_reduce_sum(s1)
        s2 = mpi._all_reduce_sum(s2)
        return s2 - (s1.conj() * s1).real

    def tangent_kernel(self):
        """Obar Obar^dagger over the samples of all ranks (reference stats.py:332-336) from the gathered
        Khatri-Rao factors (configs, tau, weights) -- O itself is never formed or communicated."""
        mu = self.kr_mean()
        scale = 2.0 if self.holomorphic else 1.0
        if mpi.commSize == 1:
            return K.rbm_gram_T(self._s, self._tau, self._p, mu, self.hasBias, scale)
        # every rank needs all Khatri-Rao factors (N_T (4N + 16M + 8) bytes in total), but T itself is formed ONCE over
        # the ranks: rank k computes the k-th range of the Hermitian tile pairs, one SUM all-reduce assembles the kernel
        s_all = mpi.gather(self._s[None])
        tau_all = mpi.gather(self._tau[None])
        p_all = mpi.gather(self._p[None])
        v_all = mpi.gather(K.rbm_krmatvec(self._s, self._tau, mu.conj().contiguous(), self.hasBias)[None])
        T = K.rbm_gram_T(s_all, tau_all, p_all, mu, self.hasBias, scale, part=(mpi.rank, mpi.commSize), v=v_all)
        return mpi._all_reduce_sum(T)

    def minsr_contract(self, x):
        """-Obar^dagger x in the reference flat layout (jVMC/util/minsr.py:65) for the gathered vector x,
        evaluated as a Khatri-Rao moment reduction (no dense O)."""
        B = self._s.shape[0]
        off = mpi.gather_offset(B)
        xl = x[off:off + B].to(torch.complex128)
        wgt = torch.sqrt(self._p) * xl
        Fk = K.rbm_moments(self._s, self._tau, wgt, self.hasBias, 1)
        Fk = mpi._all_reduce_sum(Fk - self.kr_mean().conj() * wgt.sum()).reshape(-1)
        if not self.holomorphic:
            return -Fk
        Mb = self.M if self.hasBias else 0
        parts = ([Fk[:Mb], -1j * Fk[:Mb]] if self.hasBias else []) + [Fk[Mb:], -1j * Fk[Mb:]]
        return -torch.cat(parts)
