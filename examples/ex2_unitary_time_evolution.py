#!/usr/bin/env python
"""Unitary time evolution after a quench: the reference's examples/ex2_unitary_time_evolution.py with the jax and
matplotlib lines removed; every jVMC call is written as there.

    python examples/ex2_unitary_time_evolution.py [--tmax 0.5] [--mc]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jVMC  # noqa: E402
from jVMC.util import measure  # noqa: E402
import jVMC.operator as op  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tmax", type=float, default=2.0)
ap.add_argument("--mc", action="store_true", help="Monte Carlo sampling instead of the exact sampler")
args = ap.parse_args()

L = 6
g = -0.7
h = 0.1

dt = 1e-3  # Initial time step
integratorTol = 1e-4  # Adaptive integrator tolerance
tmax = args.tmax  # Final time

# Set up variational wave function
net = jVMC.nets.CpxRBM(numHidden=10, bias=True)
psi = jVMC.vqs.NQS(net, seed=1234)  # Variational wave function

# Set up hamiltonian for time evolution
hamiltonian = jVMC.operator.BranchFreeOperator()
for l in range(L):
    hamiltonian.add(op.scal_opstr(-1., (op.Sz(l), op.Sz((l + 1) % L))))
    hamiltonian.add(op.scal_opstr(g, (op.Sx(l), )))
    hamiltonian.add(op.scal_opstr(h, (op.Sz(l),)))

# Set up observables
observables = {
    "energy": hamiltonian,
    "X": jVMC.operator.BranchFreeOperator(),
}
for l in range(L):
    observables["X"].add(op.scal_opstr(1. / L, (op.Sx(l), )))

if args.mc:
    sampler = jVMC.sampler.MCSampler(psi, (L,), 1234, updateProposer=jVMC.sampler.propose_spin_flip, numChains=500,
                                     sweepSteps=L, numSamples=20000, thermalizationSweeps=20)
else:
    sampler = jVMC.sampler.ExactSampler(psi, L)

# Set up TDVP
tdvpEquation = jVMC.util.tdvp.TDVP(sampler, pinvTol=1e-8,
                                   rhsPrefactor=1.j,
                                   makeReal='imag')

t = 0.0  # Initial time

# Set up stepper
stepper = jVMC.util.stepper.AdaptiveHeun(timeStep=dt, tol=integratorTol)

# Measure initial observables
obs = measure(observables, psi, sampler)
data = []
data.append([t, float(obs["energy"]["mean"][0]), float(obs["X"]["mean"][0])])

n = 0
t0 = time.perf_counter()
while t < tmax:
    tic = time.perf_counter()
    dp, dt = stepper.step(0, tdvpEquation, psi.get_parameters(), hamiltonian=hamiltonian, psi=psi)
    psi.set_parameters(dp)
    t += float(dt)
    obs = measure(observables, psi, sampler)
    data.append([t, float(obs["energy"]["mean"][0]), float(obs["X"]["mean"][0])])
    tdvpErr, tdvpRes = tdvpEquation.get_residuals()
    toc = time.perf_counter()
    if n % 20 == 0:
        print(">  t = %f   dt = %f   tdvp_err = %.2e, solver_res = %.2e   Energy = %f +/- %f   <X> = %f   (%.3f s/step)"
              % (t, float(dt), float(tdvpErr), float(tdvpRes), float(obs["energy"]["mean"][0]),
                 float(obs["energy"]["MC_error"][0]), float(obs["X"]["mean"][0]), toc - tic))
    n += 1

npdata = np.array(data)
print("%d steps in %.1f s; energy drift %.2e; <X>(t_max) = %.4f" % (n, time.perf_counter() - t0,
                                                                 abs(npdata[-1, 1] - npdata[0, 1]), npdata[-1, 2]))
