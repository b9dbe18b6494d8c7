"""Host-side drivers: measure() and ground_state_search() (mirror of jVMC/util/util.py:91-212)."""
import collections.abc
import time

import torch

from .. import mpi_wrapper as mpi
from . import stepper as jVMCstepper


def get_iterable(x):
    if isinstance(x, collections.abc.Iterable):
        return x
    return (x,)


def measure(observables, psi, sampler, numSamples=None):
    """Expectation values, variances and MC errors of a dict of operators (reference :91-157).
    An operator with extra positional arguments is given as a tuple (operator, *args)."""
    sampleConfigs, sampleLogPsi, p = sampler.sample(numSamples=numSamples)
    result = {}
    for name, ops in observables.items():
        means, variances, errors = [], [], []
        if not isinstance(ops, list):
            ops = [ops]
        for op in ops:
            args = ()
            if isinstance(op, (tuple, list)):
                args = tuple(op[1:])
                op = op[0]
            Oloc = op.get_O_loc(sampleConfigs, psi, sampleLogPsi, *args)
            means.append(mpi.global_mean(Oloc[..., None], p)[0])
            variances.append(mpi.global_variance(Oloc[..., None], p)[0])
            errors.append(torch.sqrt(variances[-1]) / (sampler.get_last_number_of_samples() ** 0.5))
        result[name] = {"mean": torch.stack(means).real, "variance": torch.stack(variances).real,
                        "MC_error": torch.stack(errors).real}
    return result


def ground_state_search(psi, ham, tdvpEquation, sampler, numSteps=200, varianceTol=1e-10, stepSize=1e-2,
                        observables=None, outp=None):
    """Ground-state search by stochastic reconfiguration (reference :160-212): Euler steps, the diagonal
    shift decays by 0.95 per step, stops when the energy variance falls below varianceTol."""
    delta = getattr(tdvpEquation, "diagonalShift", None)
    stepper = jVMCstepper.Euler(timeStep=stepSize)
    n = 0
    if outp is not None and observables is not None:
        outp.write_observables(n, **measure(observables, psi, sampler))
    varE = 1.0
    while n < numSteps and varE > varianceTol:
        tic = time.perf_counter()
        dp, _ = stepper.step(0, tdvpEquation, psi.get_parameters(), hamiltonian=ham, psi=psi, numSamples=None,
                             outp=outp)
        psi.set_parameters(dp)
        n += 1
        varE = float(tdvpEquation.get_energy_variance())
        if outp is not None and observables is not None:
            outp.write_observables(n, **measure(observables, psi, sampler))
        if hasattr(tdvpEquation, "set_diagonal_shift"):
            delta = 0.95 * delta
            tdvpEquation.set_diagonal_shift(delta)
        if outp is not None:
            outp.print(" STEP %d" % (n))
            outp.print("   Energy mean: %f" % (float(tdvpEquation.get_energy_mean())))
            outp.print("   Energy variance: %f" % (varE))
            outp.print_timings(indent="   ")
            outp.print("   == Time for step: %fs" % (time.perf_counter() - tic))
