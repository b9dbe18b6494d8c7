"""Phase timings of TDVP.__call__ (reference phase names, jVMC/util/tdvp.py:267-285) on configs 1 and 3."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vmc_jax_b200 as jVMC  # noqa: E402
import vmc_jax_b200.operator as op  # noqa: E402


def run(L, alpha, bias, nsamp, chains, x, makeReal, shift, proposer, scale):
    dev = jVMC.global_defs.myDevice
    psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=alpha * L, bias=bias), seed=1234)
    psi(torch.zeros((1, 1, L), dtype=torch.int32, device=dev))
    if scale:
        W, b = bench.o1_weights(L, alpha * L, bias)
        psi.set_parameters(torch.as_tensor(bench.flat_params(scale * W, None if b is None else scale * b)))
    H = op.BranchFreeOperator()
    for l in range(L):
        H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz((l + 1) % L))))
        H.add(op.scal_opstr(-0.7, (op.Sx(l),)))
    smp = jVMC.sampler.MCSampler(psi, (L,), 4321, updateProposer=proposer, numChains=chains, sweepSteps=L,
                                 numSamples=nsamp, thermalizationSweeps=25)
    tdvp = jVMC.util.TDVP(smp, rhsPrefactor=x, pinvTol=1e-8, diagonalShift=shift, makeReal=makeReal)
    outp = jVMC.util.OutputManager(None)
    for k in range(4):
        if k == 1:
            outp.timings = {}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tdvp(psi.get_parameters(), 0.0, hamiltonian=H, psi=psi, numSamples=None, outp=outp)
        torch.cuda.synchronize()
        tot = (time.perf_counter() - t0) * 1e3
    print("L=%d P=%d samples=%d: last call %.1f ms" % (L, psi.numParameters, smp.get_last_number_of_samples(), tot))
    for k, v in outp.timings.items():
        print("   %-22s %8.2f ms / call" % (k, 1e3 * v["total"] / max(v["count"], 1)))


run(20, 2, False, 4096, 500, 1., 'real', 10, jVMC.sampler.propose_spin_flip_Z2, None)
run(40, 2, True, 2 ** 16, 2368, 1.j, 'imag', 0., jVMC.sampler.propose_spin_flip, 0.3)
