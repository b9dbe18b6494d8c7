#!/usr/bin/env python
"""Real-time TDVP after a field quench of the 1-D Ising chain with the adaptive Heun stepper, observables through
jVMC.util.measure.  Workflow of the reference's time-evolution example, own code.

    python examples/tfim_quench.py [--L 6] [--tmax 0.5] [--mc]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jVMC  # noqa: E402
from jVMC.operator import BranchFreeOperator, Sx, Sz, scal_opstr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=6)
    ap.add_argument("--tmax", type=float, default=0.5)
    ap.add_argument("--mc", action="store_true", help="Metropolis sampling (20 000 samples) instead of full enumeration")
    a = ap.parse_args()
    L, g, h = a.L, -0.7, 0.1

    psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=10, bias=True), seed=1234)
    H, X = BranchFreeOperator(), BranchFreeOperator()
    for i in range(L):
        H.add(scal_opstr(-1.0, (Sz(i), Sz((i + 1) % L))))
        H.add(scal_opstr(g, (Sx(i),)))
        H.add(scal_opstr(h, (Sz(i),)))
        X.add(scal_opstr(1.0 / L, (Sx(i),)))
    observables = {"energy": H, "X": X}

    if a.mc:
        smp = jVMC.sampler.MCSampler(psi, (L,), 1234, updateProposer=jVMC.sampler.propose_spin_flip, numChains=500,
                                     sweepSteps=L, numSamples=20000, thermalizationSweeps=20)
    else:
        smp = jVMC.sampler.ExactSampler(psi, L)
    tdvp = jVMC.util.TDVP(smp, pinvTol=1e-8, rhsPrefactor=1.0j, makeReal="imag")
    heun = jVMC.util.stepper.AdaptiveHeun(timeStep=1e-3, tol=1e-4)

    t, n = 0.0, 0
    first = jVMC.util.measure(observables, psi, smp)
    e_start = float(first["energy"]["mean"][0])
    tic = time.perf_counter()
    while t < a.tmax:
        theta, dt = heun.step(0, tdvp, psi.get_parameters(), hamiltonian=H, psi=psi)
        psi.set_parameters(theta)
        t += float(dt)
        n += 1
        if n % 25 == 0:
            obs = jVMC.util.measure(observables, psi, smp)
            err, res = tdvp.get_residuals()
            print("t = %.4f  dt = %.2e  <X> = %.5f  E = %.6f  tdvp_err = %.1e  residual = %.1e" % (
                t, float(dt), float(obs["X"]["mean"][0]), float(obs["energy"]["mean"][0]), float(err), float(res)))
    last = jVMC.util.measure(observables, psi, smp)
    print("%d adaptive steps in %.1f s, energy drift %.2e, <X>(%.2f) = %.4f" % (
        n, time.perf_counter() - tic, abs(float(last["energy"]["mean"][0]) - e_start), t, float(last["X"]["mean"][0])))


if __name__ == "__main__":
    main()
