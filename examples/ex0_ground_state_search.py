#!/usr/bin/env python
"""Ground-state search for the transverse-field Ising chain: the script of the reference's
examples/ex0_ground_state_search.py with the jax imports removed -- every jVMC call is written as there
(`import jVMC` resolves to the B200-native implementation).  Differences: the PRNG key is an integer instead of
jax.random.PRNGKey(4321), printed values are torch scalars, no matplotlib output.

    python examples/ex0_ground_state_search.py [--cnn] [--steps N]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jVMC  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cnn", action="store_true", help="the reference's GPU branch: CNN(F=(L,), channels=(16,)) with 40000 samples")
ap.add_argument("--steps", type=int, default=None)
args = ap.parse_args()

L = 10
g = -0.7

if args.cnn:
    # the reference uses actFun = 1 + elu(x) here, a user-defined Python function; device kernels exist for the named
    # activations of jVMC.nets.activation_functions, so plain elu is used
    net = jVMC.nets.CNN(F=(L,), channels=(16,), strides=(1,), periodicBoundary=True,
                        actFun=(jVMC.nets.activation_functions.elu,))
    n_steps = 1000
    n_Samples = 40000
else:
    net = jVMC.nets.CpxRBM(numHidden=8, bias=False)
    n_steps = 300
    n_Samples = 5000
if args.steps is not None:
    n_steps = args.steps

psi = jVMC.vqs.NQS(net, seed=1234)  # Variational wave function


def energy_single_p_mode(h_t, P):
    return np.sqrt(1 + h_t**2 - 2 * h_t * np.cos(P))


def ground_state_energy_per_site(h_t, N):
    Ps = 0.5 * np.arange(- (N - 1), N - 1 + 2, 2)
    Ps = Ps * 2 * np.pi / N
    energies_p_modes = np.array([energy_single_p_mode(h_t, P) for P in Ps])
    return - 1 / N * np.sum(energies_p_modes)


exact_energy = ground_state_energy_per_site(g, L)
print(exact_energy)

# Set up hamiltonian
hamiltonian = jVMC.operator.BranchFreeOperator()
for l in range(L):
    hamiltonian.add(jVMC.operator.scal_opstr(-1., (jVMC.operator.Sz(l), jVMC.operator.Sz((l + 1) % L))))
    hamiltonian.add(jVMC.operator.scal_opstr(g, (jVMC.operator.Sx(l), )))

# Set up sampler
sampler = jVMC.sampler.MCSampler(psi, (L,), 4321, updateProposer=jVMC.sampler.propose_spin_flip_Z2,
                                 numChains=100, sweepSteps=L,
                                 numSamples=n_Samples, thermalizationSweeps=25)

# Set up TDVP
tdvpEquation = jVMC.util.tdvp.TDVP(sampler, rhsPrefactor=1.,
                                   svdTol=1e-8, diagonalShift=10, makeReal='real')

stepper = jVMC.util.stepper.Euler(timeStep=1e-2)  # ODE integrator

res = []
for n in range(n_steps):

    dp, _ = stepper.step(0, tdvpEquation, psi.get_parameters(), hamiltonian=hamiltonian, psi=psi, numSamples=None)
    psi.set_parameters(dp)

    e, v = float(tdvpEquation.ElocMean0.real) / L, float(tdvpEquation.ElocVar0) / L
    if n % 25 == 0 or n == n_steps - 1:
        print(n, e, v)
    res.append([n, e, v])

res = np.array(res)
print("final (E - E0)/L = %.3e, Var(E)/L = %.3e" % (res[-1, 1] - exact_energy, res[-1, 2]))
