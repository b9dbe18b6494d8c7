"""Mirror of jVMC/util/minsr.py: energy minimisation via MinSR (arXiv:2302.01941)."""
import torch

from .. import kernels as K
from .. import mpi_wrapper as mpi
from ..stats import SampledObs, RBMGradientObs


def _device_image(T):
    """private column-major image of T for the solver (which overwrites it); no host path"""
    if not T.is_cuda:
        raise RuntimeError("MinSR pseudo-inverse needs the CUDA library (no CPU fallback)")
    return T.T.clone(memory_format=torch.contiguous_format)


def pinv_hermitian(T, rtol):
    """jnp.linalg.pinv(T, rtol=..., hermitian=True) (reference jVMC/util/minsr.py:61,73): eigenvalues with
    |ev| <= rtol * max|ev| are dropped."""
    ev, Vt, _ = K.eigh_inplace(_device_image(T))
    V = Vt.T
    cut = rtol * ev.abs().max()
    inv = torch.where(ev.abs() > cut, 1.0 / torch.where(ev == 0, torch.ones_like(ev), ev), torch.zeros_like(ev))
    return (V * inv.to(V.dtype)[None, :]) @ V.conj().T


def pinv_hermitian_apply(T, rtol, b):
    """pinv(T, rtol, hermitian=True) @ b without forming the pseudo-inverse -- one C-ABI call (jvmc_minsr_solve:
    eigen-decomposition, eigenvalue cut, V (w * V^dagger b))."""
    x, _ = K.minsr_solve(_device_image(T), b.to(torch.complex128), rtol)
    return x if T.is_complex() or b.is_complex() else x.real


class MinSR:
    """reference jVMC/util/minsr.py:16-165; constructor as in the reference (:28)."""

    def __init__(self, sampler, pinvTol=1e-14, diagonalShift=0., diagonalizeOnDevice=True):
        self.sampler = sampler
        self.pinvTol = pinvTol
        self.diagonalShift = diagonalShift
        self.diagonalizeOnDevice = diagonalizeOnDevice
        self.metaData = None

    def set_pinv_tol(self, tol):
        self.pinvTol = tol

    def get_metadata(self):
        return self.metaData

    def get_energy_variance(self):
        return self.ElocVar0

    def get_energy_mean(self):
        return self.ElocMean0.real

    def solve(self, eloc, gradients, holomorphic):
        """reference :53-80."""
        if holomorphic:
            T = gradients.tangent_kernel()
            eloc_all = mpi.gather(eloc._data).reshape(-1).to(T.dtype)
            x = pinv_hermitian_apply(T, self.pinvTol, eloc_all)        # = pinv(T) @ eloc (reference :61-62)
            if isinstance(gradients, RBMGradientObs):
                return gradients.minsr_contract(x)
            gradients_all = mpi.gather(gradients._data)
            gradients_all = gradients_all.reshape(gradients_all.shape[0], -1)
            return -gradients_all.conj().T @ x
        gradients_all = mpi.gather(gradients._data)
        gradients_all = gradients_all.reshape(gradients_all.shape[0], -1)
        G = torch.cat([gradients_all.real, gradients_all.imag], dim=0)
        T = G @ G.T
        T = T + self.diagonalShift * torch.eye(T.shape[-1], dtype=T.dtype, device=T.device)
        eloc_all = mpi.gather(eloc._data).reshape(-1)
        eloc_all = torch.cat([eloc_all.real, eloc_all.imag], dim=0)
        return -G.T @ pinv_hermitian_apply(T, self.pinvTol, eloc_all)

    def __call__(self, netParameters, t, *, psi, hamiltonian, **rhsArgs):
        """reference :82-165."""
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
            sampleGradients = RBMGradientObs(psi, sampleConfigs, p)
        else:
            sampleGradients = SampledObs(psi.gradients(sampleConfigs), p)
        stop_timing("compute gradients")
        start_timing("solve MinSR eqn.")
        update = self.solve(Eloc, sampleGradients, holomorphic=psi.holomorphic)
        stop_timing("solve MinSR eqn.")
        if outp is not None:
            outp.add_timing("MPI communication", mpi.get_communication_time())
        psi.set_parameters(tmpParameters)
        if "intStep" in rhsArgs and rhsArgs["intStep"] == 0:
            self.ElocMean0 = Eloc.mean()[0]
            self.ElocVar0 = Eloc.var()[0]
            self.metaData = {}
        return update
