"""ONE full TDVP step (sample + E_loc + S/F + eigen-decomposition + SNR + regularised solve) at BASELINE configs[1]:
2D TFIM 10x10, CpxRBM alpha=4 (P_c = 40 000), 2^16 samples, imaginary time (rhsPrefactor=1, makeReal='real').
The 40 000 x 40 000 Hermitian eigen-decomposition is beyond cuSOLVER's dense solvers (n <= 32768): it runs through
kernels.eigh_large (hetrd + one level of Cuppen tearing + unmtr), about two minutes: a one-off measurement, not part of
bench.py.   python tools/tdvp_config2.py [--alpha 4] [--L 10]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vmc_jax_b200 as jVMC  # noqa: E402
import vmc_jax_b200.operator as op  # noqa: E402

import argparse  # noqa: E402
_ap = argparse.ArgumentParser()
_ap.add_argument("--alpha", type=int, default=4)
_ap.add_argument("--L", type=int, default=10)
_args = _ap.parse_args()
os.environ.setdefault("JVMC_EIGH_VERBOSE", "1")
Lx = Ly = _args.L
N, M = Lx * Ly, _args.alpha * Lx * Ly
dev = jVMC.global_defs.myDevice
psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=M, bias=False), seed=1234)
psi(torch.zeros((1, 1, Lx, Ly), dtype=torch.int32, device=dev))
W, _ = bench.o1_weights(N, M, False)
psi.set_parameters(torch.as_tensor(bench.flat_params(W, None)))
H = op.BranchFreeOperator()
for x in range(Lx):
    for y in range(Ly):
        l = x * Ly + y
        H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz(x * Ly + (y + 1) % Ly))))
        H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz(((x + 1) % Lx) * Ly + y))))
        H.add(op.scal_opstr(3.04, (op.Sx(l),)))
smp = jVMC.sampler.MCSampler(psi, (Lx, Ly), 4321, updateProposer=jVMC.sampler.propose_spin_flip, numChains=2368,
                             sweepSteps=N, numSamples=2 ** 16, thermalizationSweeps=25)
tdvp = jVMC.util.TDVP(smp, snrTol=2, pinvTol=1e-8, rhsPrefactor=1., diagonalShift=1e-3, makeReal='real')
outp = jVMC.util.OutputManager(None)
torch.cuda.synchronize()
t0 = time.perf_counter()
upd = tdvp(psi.get_parameters(), 0.0, hamiltonian=H, psi=psi, numSamples=None, outp=outp)
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print("config 2 full TDVP step: %.1f s, %d samples, P = %d, |update| = %.3e, peak memory %.1f GB"
      % (tot, smp.get_last_number_of_samples(), upd.numel(), float(upd.norm()), torch.cuda.max_memory_allocated() / 1e9))
for k, v in outp.timings.items():
    print("   %-22s %10.2f s" % (k, v["total"]))
print("   spectrum: max %.3e, min %.3e; cutoff-regularised components with SNR > 2: %d of %d"
      % (float(tdvp.ev.max()), float(tdvp.ev.min()), int((tdvp.snr > 2).sum()), tdvp.snr.numel()))
