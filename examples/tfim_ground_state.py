#!/usr/bin/env python
"""Stochastic-reconfiguration ground-state search for the 1-D transverse-field Ising chain through the jVMC API
(NQS, MCSampler, BranchFreeOperator, TDVP, Euler).  Workflow of the reference's first example, own code.

    python examples/tfim_ground_state.py [--L 10] [--g -0.7] [--hidden 8] [--samples 5000] [--steps 300] [--cnn]
"""
import argparse
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jVMC  # noqa: E402  (alias of vmc_jax_b200)
from jVMC.operator import BranchFreeOperator, Sx, Sz, scal_opstr  # noqa: E402


def exact_energy_per_site(g, L):
    """free-fermion solution of the periodic chain (even parity sector)"""
    ks = [2.0 * math.pi * (n + 0.5) / L for n in range(-L // 2, L // 2)]
    return -sum(math.sqrt(1.0 + g * g - 2.0 * g * math.cos(k)) for k in ks) / L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=10)
    ap.add_argument("--g", type=float, default=-0.7)
    ap.add_argument("--hidden", type=int, default=8)
    ap.add_argument("--samples", type=int, default=5000)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--cnn", action="store_true", help="real CNN ansatz instead of the complex RBM")
    a = ap.parse_args()

    net = jVMC.nets.CNN(F=(a.L,), channels=(16,), strides=(1,)) if a.cnn else jVMC.nets.CpxRBM(numHidden=a.hidden, bias=False)
    psi = jVMC.vqs.NQS(net, seed=1234)

    H = BranchFreeOperator()
    for site in range(a.L):
        H.add(scal_opstr(-1.0, (Sz(site), Sz((site + 1) % a.L))))
        H.add(scal_opstr(a.g, (Sx(site),)))

    mc = jVMC.sampler.MCSampler(psi, (a.L,), 4321, updateProposer=jVMC.sampler.propose_spin_flip_Z2, numChains=100,
                                sweepSteps=a.L, numSamples=a.samples, thermalizationSweeps=25)
    sr = jVMC.util.TDVP(mc, rhsPrefactor=1.0, pinvTol=1e-8, diagonalShift=10, makeReal="real")
    euler = jVMC.util.stepper.Euler(timeStep=1e-2)

    e0 = exact_energy_per_site(a.g, a.L)
    print("exact E0/L = %.8f" % e0)
    tic = time.perf_counter()
    for it in range(a.steps):
        theta, _ = euler.step(0, sr, psi.get_parameters(), hamiltonian=H, psi=psi, numSamples=None)
        psi.set_parameters(theta)
        if it % 25 == 0 or it == a.steps - 1:
            print("%4d  E/L = %.6f   (E - E0)/L = %.2e   Var(E)/L = %.3e" % (
                it, float(sr.ElocMean0.real) / a.L, float(sr.ElocMean0.real) / a.L - e0, float(sr.ElocVar0) / a.L))
    print("%d SR steps in %.1f s" % (a.steps, time.perf_counter() - tic))


if __name__ == "__main__":
    main()
