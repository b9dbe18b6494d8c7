"""Times the non-Gram kernels of the VMC step at a BASELINE shape and reports achieved fp64 rates (development aid).
    python tools/kernel_bench.py [--shape 10 10] [--alpha 4] [--samples 65536] [--chains 2368]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vmc_jax_b200 as jVMC  # noqa: E402
import vmc_jax_b200.operator as op  # noqa: E402
from vmc_jax_b200 import kernels as K  # noqa: E402
from vmc_jax_b200.stats import SampledObs, RBMGradientObs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", type=int, nargs="+", default=[10, 10])
ap.add_argument("--alpha", type=int, default=4)
ap.add_argument("--samples", type=int, default=65536)
ap.add_argument("--chains", type=int, default=2368)
ap.add_argument("--therm", type=int, default=25)
ap.add_argument("--g", type=float, default=3.04)
ap.add_argument("--refresh", type=int, default=8)
ap.add_argument("--generic", action="store_true", help="force the generic sampler kernel")
ap.add_argument("--minsr", type=int, default=0, help="also time the MinSR tangent kernel T on the first N_T samples")
a = ap.parse_args()
if a.generic:
    from vmc_jax_b200 import _lib
    _lib.load().jvmc_mcmc_set_generic(1)
shape = tuple(a.shape)
N = int(np.prod(shape))
M = a.alpha * N
dev = jVMC.global_defs.myDevice
psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=M, bias=False), seed=1)
psi(torch.zeros((1, 1) + shape, dtype=torch.int32, device=dev))
rng = np.random.default_rng(4321)
aa = 1.0 / np.sqrt(N)
W = rng.uniform(-aa, aa, (N, M)) + 1j * rng.uniform(-aa, aa, (N, M))
psi.set_parameters(torch.as_tensor(np.concatenate([W.ravel().real, W.ravel().imag])))
H = op.BranchFreeOperator()
if len(shape) == 1:
    for l in range(N):
        H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz((l + 1) % N))))
        H.add(op.scal_opstr(a.g, (op.Sx(l),)))
else:
    Lx, Ly = shape
    for x in range(Lx):
        for y in range(Ly):
            l = x * Ly + y
            H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz(x * Ly + (y + 1) % Ly))))
            H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz(((x + 1) % Lx) * Ly + y))))
            H.add(op.scal_opstr(a.g, (op.Sx(l),)))
smp = jVMC.sampler.MCSampler(psi, shape, 4321, updateProposer=jVMC.sampler.propose_spin_flip, numChains=a.chains,
                             sweepSteps=N, thermalizationSweeps=a.therm, numSamples=a.samples)
smp.refreshEvery = a.refresh


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    out = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return out, best


(s, logPsi, p), t_s = timed(lambda: smp.sample())
B = s.shape[1]
acc = float(smp.acceptance_ratio())
spc = B // a.chains
props = a.chains * N * (a.therm + spc)
fl_s = props * M * (14 + 30 * acc)
print("sample(): %.2f ms for %d samples (%d proposals, acceptance %.3f): %.2f TFLOP/s fp64 (SURVEY 8d count), %.1f M proposals/s"
      % (t_s, B, props, acc, fl_s / t_s / 1e9, props / t_s / 1e3))
flat = s.reshape(B, -1).contiguous()
Wd, bd = psi._cW()
_, t_l = timed(lambda: K.rbm_logpsi(flat, Wd, bd))
print("rbm_logpsi: %.3f ms: %.2f TFLOP/s (4NM adds) + %d transcendentals" % (t_l, 4.0 * N * M * B / t_l / 1e9, B * M))
E, t_e = timed(lambda: H.get_O_loc(s, psi, logPsi))
print("fused E_loc (API call): %.3f ms: %.2f TFLOP/s (14 N M flop/sample)" % (t_e, 14.0 * N * M * B / t_e / 1e9))
tab = H._ensure_compiled(0)
tau_d, tables_d, dt_d, pref_d = psi._tau(flat), psi.flip_tables(), tab.device_tables(), tab.eval_prefactors()
_, t_k = timed(lambda: K.rbm_eloc(flat, tau_d, tables_d, dt_d, pref_d))
print("fused E_loc (kernel): %.3f ms: %.2f TFLOP/s (14 N M flop/sample), %.2f TB/s of weight rows if unshared"
      % (t_k, 14.0 * N * M * B / t_k / 1e9, 16.0 * M * B * tab.numOps / t_k / 1e9))
G = RBMGradientObs(psi, s, p)
_, t_m = timed(lambda: K.rbm_moments(G._s, G._tau, G._p.to(torch.complex128), False, 0))
print("rbm_moments: %.3f ms: %.2f TFLOP/s (8 N M flop/sample)" % (t_m, 8.0 * N * M * B / t_m / 1e9))

if a.minsr > 0:
    NT = min(a.minsr, B)
    Gs = RBMGradientObs(psi, s[:, :NT], (p[:, :NT] / p[:, :NT].sum()))
    mu = Gs.kr_mean()
    _, t_t = timed(lambda: K.rbm_gram_T(Gs._s, Gs._tau, Gs._p, mu, False, 2.0), reps=2)
    print("MinSR tangent kernel T (N_T = %d, %0.1f GB): %.2f ms: %.2f TFLOP/s (4 N_T^2 M, Hermitian half), dense O O^dagger would be "
          "%.1f PFLOP" % (NT, NT * NT * 16 / 1e9, t_t, 4.0 * NT * NT * M / t_t / 1e9, 4.0 * NT * NT * N * M / 1e15))
