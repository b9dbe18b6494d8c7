"""-m gpu: the jVMC-compatible Python API on top of the CUDA kernels, written like the reference's own
tests (tests/tdvp_test.py, minsr_test.py, sampler_test.py, operator_test.py, stats_test.py, vqs_test.py)
and checked against the reference's golden values and the CPU oracle."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import jVMC  # noqa: E402  (alias of vmc_jax_b200)
import jVMC.nets as nets  # noqa: E402
import jVMC.operator as op  # noqa: E402
import jVMC.sampler as sampler  # noqa: E402
import jVMC.util.stepper as jVMCstepper  # noqa: E402
from jVMC.util import measure, ground_state_search  # noqa: E402
from jVMC.vqs import NQS  # noqa: E402
from jVMC.stats import SampledObs, RBMGradientObs  # noqa: E402
import jVMC.mpi_wrapper as mpi  # noqa: E402

from oracle import rbm as orbm, bfo as obfo, stats as ostats, solve as osolve  # noqa: E402

with open(os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")) as f:
    REFG = json.load(f)
WEIGHTS = torch.tensor(REFG["rbm_weights"], dtype=torch.float64)


def tfim(L, J, hx):
    h = op.BranchFreeOperator()
    for l in range(L):
        h.add(op.scal_opstr(J, (op.Sz(l), op.Sz((l + 1) % L))))
        h.add(op.scal_opstr(hx, (op.Sx(l),)))
    return h


def host(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------ tests/tdvp_test.py
@pytest.mark.parametrize("k", [0, 1])
def test_gs_search_cpx(k):
    """reference tests/tdvp_test.py:27-56."""
    L, J = 4, -1.0
    hx, exE = REFG["gs_hx"][k], REFG["gs_energies"][k]
    psi = NQS(nets.CpxRBM(numHidden=6, bias=False))
    H = tfim(L, J, hx)
    exactSampler = sampler.ExactSampler(psi, L)
    tdvpEquation = jVMC.util.TDVP(exactSampler, snrTol=1, pinvTol=0.0, pinvCutoff=1e-8, rhsPrefactor=1.,
                                  diagonalShift=2, makeReal='real')
    ground_state_search(psi, H, tdvpEquation, exactSampler, numSteps=100, stepSize=5e-2)
    obs = measure({"energy": H}, psi, exactSampler)
    assert float(torch.max(torch.abs((obs['energy']['mean'] - exE) / exE))) < 1e-3


def test_time_evolution_golden():
    """reference tests/tdvp_test.py:60-128: <ZZ>(t) golden trajectory + energy conservation."""
    L, J, hx = 4, -1.0, -0.3
    psi = NQS(nets.CpxRBM(numHidden=2, bias=False))
    psi(torch.tensor([[[1, 1, 1, 1]]], dtype=torch.int32))
    psi.set_parameters(WEIGHTS)
    hamiltonian = tfim(L, J, hx)
    ZZ = op.BranchFreeOperator()
    for l in range(L):
        ZZ.add((op.Sz(l), op.Sz((l + 1) % L)))
    exactSampler = sampler.ExactSampler(psi, L)
    stepper = jVMCstepper.AdaptiveHeun(timeStep=1e-3, tol=1e-5)
    tdvpEquation = jVMC.util.TDVP(exactSampler, snrTol=1, pinvTol=0.0, pinvCutoff=1e-8, rhsPrefactor=1.j,
                                  diagonalShift=0., makeReal='imag')
    t, obs, times = 0, [], [0]
    newMeas = measure({'E': hamiltonian, 'ZZ': ZZ}, psi, exactSampler)
    obs.append([float(newMeas['E']['mean'][0]), float(newMeas['ZZ']['mean'][0])])
    while t < 0.5:
        dp, dt = stepper.step(0, tdvpEquation, psi.get_parameters(), hamiltonian=hamiltonian, psi=psi, numSamples=0)
        psi.set_parameters(dp)
        t += dt
        times.append(t)
        newMeas = measure({'E': [(hamiltonian, t)], 'ZZ': ZZ}, psi, exactSampler)
        obs.append([float(newMeas['E']['mean'][0]), float(newMeas['ZZ']['mean'][0])])
    obs = np.array(obs)
    assert np.max(np.abs((obs[:, 0] - obs[0, 0]) / obs[0, 0])) < 1e-3
    refTimes = np.arange(0, 0.5, 0.05)
    netZZ = np.interp(refTimes, np.array(times), obs[:, 1])
    assert np.max(np.abs(netZZ - np.array(REFG["zz_trajectory"])[:len(netZZ)])) < 1e-3
    err, res = tdvpEquation.get_residuals()
    assert float(err) >= 0 and float(res) >= 0


def test_time_evolution_mc_sampler():
    """reference tests/tdvp_test.py:131-198 (MCSampler, mu=1, crossValidation)."""
    L, J, hx = 4, -1.0, -0.3
    psi = NQS(nets.CpxRBM(numHidden=2, bias=False), batchSize=5000)
    psi(torch.tensor([[[1, 1, 1, 1]]], dtype=torch.int32))
    psi.set_parameters(WEIGHTS)
    hamiltonian = tfim(L, J, hx)
    ZZ = op.BranchFreeOperator()
    for l in range(L):
        ZZ.add((op.Sz(l), op.Sz((l + 1) % L)))
    MCsampler = sampler.MCSampler(psi, (L,), 0, numSamples=50000, updateProposer=sampler.propose_spin_flip, mu=1,
                                  numChains=500)
    stepper = jVMCstepper.AdaptiveHeun(timeStep=1e-3, tol=1e-4)
    tdvpEquation = jVMC.util.TDVP(MCsampler, snrTol=1, pinvTol=1e-8, rhsPrefactor=1.j, diagonalShift=0.,
                                  makeReal='imag', crossValidation=True)
    t, obs, times = 0, [], [0]
    newMeas = measure({'E': hamiltonian, 'ZZ': ZZ}, psi, MCsampler)
    obs.append([float(newMeas['E']['mean'][0]), float(newMeas['ZZ']['mean'][0])])
    while t < 0.2:
        dp, dt = stepper.step(0, tdvpEquation, psi.get_parameters(), hamiltonian=hamiltonian, psi=psi,
                              numSamples=5000)
        psi.set_parameters(dp)
        t += dt
        times.append(t)
        newMeas = measure({'E': [(hamiltonian, t)], 'ZZ': ZZ}, psi, MCsampler)
        obs.append([float(newMeas['E']['mean'][0]), float(newMeas['ZZ']['mean'][0])])
    obs = np.array(obs)
    assert np.max(np.abs((obs[:, 0] - obs[0, 0]) / obs[0, 0])) < 1e-1
    refTimes = np.arange(0, 0.2, 0.05)
    netZZ = np.interp(refTimes, np.array(times), obs[:, 1])
    assert np.max(np.abs(netZZ - np.array(REFG["zz_trajectory"])[:len(netZZ)])) < 2e-2
    assert "tdvp_residual_cross_validation_ratio" in tdvpEquation.get_metadata()


def test_snr_matches_oracle():
    """SNR / rhoVar of the factorised path vs the oracle's covar_data().transform().var() (reference
    tests/tdvp_test.py:201-263 compares the same quantity against its legacy formula)."""
    L = 4
    psi = NQS(nets.CpxRBM(numHidden=2, bias=False))
    psi(torch.tensor([[[1, 1, 1, 1]]], dtype=torch.int32))
    psi.set_parameters(WEIGHTS)
    H = tfim(L, -1.0, -0.3)
    smp = sampler.MCSampler(psi, (L,), 0, numSamples=10, updateProposer=sampler.propose_spin_flip, mu=2,
                            numChains=500)
    s, logPsi, p = smp.sample()
    Eloc = H.get_O_loc(s, psi, logPsi, 0.0)
    td = jVMC.util.TDVP(smp, snrTol=1, pinvTol=1e-8, rhsPrefactor=1.j, makeReal='imag')
    upd, res, cut = td.solve(SampledObs(Eloc, p), RBMGradientObs(psi, s, p))
    W, b = orbm.unflatten_params(np.array(REFG["rbm_weights"]), 4, 2)
    sn, pn, En = host(s)[0], host(p)[0], host(Eloc)[0]
    otd = osolve.TDVP(snrTol=1, pinvTol=1e-8, rhsPrefactor=1.j, makeReal='imag')
    oE, oG = ostats.SampledObs(En, pn), ostats.SampledObs(orbm.gradients_holomorphic(sn, W, b), pn)
    upd_ref, res_ref, cut_ref = otd.solve(oE, oG, mpi.globNumSamples)
    assert np.allclose(host(td.ev), otd.ev, atol=1e-10 * np.abs(otd.ev).max())     # 1e-10 of the scale of S
    # rhoVar summed over each (numerically) degenerate eigenspace is basis independent
    assert np.isclose(host(td.rhoVar).sum(), otd.rhoVar.sum(), rtol=1e-8)
    assert np.isclose(float(td.ElocVar), otd.ElocVar, rtol=1e-10)
    assert np.allclose(host(td.F0), otd.F0, rtol=1e-10, atol=1e-14)


# ------------------------------------------------------------------ tests/minsr_test.py
@pytest.mark.parametrize("k", [0, 1])
def test_minsr_gs_search(k):
    """reference tests/minsr_test.py:15-43."""
    L, J = 4, -1.0
    hx, exE = REFG["gs_hx"][k], REFG["gs_energies"][k]
    psi = NQS(nets.CpxRBM(numHidden=8, bias=False), seed=1234)
    H = tfim(L, J, hx)
    exactSampler = sampler.ExactSampler(psi, L)
    tdvpEquation = jVMC.util.MinSR(exactSampler, pinvTol=1e-6, diagonalShift=0.)
    ground_state_search(psi, H, tdvpEquation, exactSampler, numSteps=200, stepSize=1e-2)
    obs = measure({"energy": H}, psi, exactSampler)
    assert float(torch.max(torch.abs((obs['energy']['mean'] - exE) / exE))) < 1e-3


# ------------------------------------------------------------------ tests/sampler_test.py
def state_to_int(s):
    return (s.to(torch.int64) * (2 ** torch.arange(s.shape[-1], device=s.device))).sum(-1)


@pytest.mark.parametrize("mu,kappa", [(2, 0.5), (1, 0.5), (1, 1.0)])
def test_MCMC_sampling(mu, kappa):
    """reference tests/sampler_test.py:32-182 (bare CpxRBM as in :563-593): histogram vs exact, 2e-3."""
    L = 4
    psi = NQS(nets.CpxRBM(numHidden=2, bias=False))
    exactSampler = sampler.ExactSampler(psi, L, logProbFactor=kappa)
    mcSampler = sampler.MCSampler(psi, (L,), 0, updateProposer=sampler.propose_spin_flip, numChains=777, mu=mu,
                                  logProbFactor=kappa)
    p0 = psi.get_parameters()
    psi.set_parameters(WEIGHTS)
    _, _, pex = exactSampler.sample()
    numSamples = 500000
    smc, _, p = mcSampler.sample(numSamples=numSamples)
    assert smc.shape[1] >= numSamples and mcSampler.get_last_number_of_samples() == smc.shape[1]
    ints = state_to_int(smc.reshape(-1, L))
    pmc = torch.zeros(16, dtype=torch.float64, device=ints.device).index_add_(0, ints, p[0])
    pmc = pmc / pmc.sum()
    assert float(torch.max(torch.abs(pmc - pex.reshape(-1)[:16]))) < 2e-3
    acc = float(mcSampler.acceptance_ratio())
    assert 0.0 < acc <= 1.0
    # sample(parameters=...) returns amplitudes of the temporary parameters and restores the net
    s, psi_s, _ = mcSampler.sample(parameters=p0, numSamples=100)
    assert torch.allclose(psi.get_parameters().cpu(), WEIGHTS)
    psi.set_parameters(p0)
    psi_s1 = psi(s)
    assert float(torch.max(torch.abs((psi_s - psi_s1) / psi_s))) < 1e-14


@pytest.mark.parametrize("mu,kappa", [(2, 0.5), (1, 0.5), (1, 1.0)])
def test_MCMC_sampling_symnet(mu, kappa):
    """reference tests/sampler_test.py:32-182 as written there: the RBM wrapped into SymNet with the
    (translation, reflection, spinflip) orbit of the chain."""
    import warnings
    L = 4
    rbm = nets.CpxRBM(numHidden=2, bias=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        orbit = jVMC.util.symmetries.get_orbit_1D(L, "translation", "reflection", "spinflip")
    net = nets.sym_wrapper.SymNet(net=rbm, orbit=orbit)
    psi = NQS(net)
    exactSampler = sampler.ExactSampler(psi, L, logProbFactor=kappa)
    mcSampler = sampler.MCSampler(psi, (L,), 0, updateProposer=sampler.propose_spin_flip, numChains=777, mu=mu,
                                  logProbFactor=kappa)
    p0 = psi.get_parameters()
    psi.set_parameters(WEIGHTS)
    _, _, pex = exactSampler.sample()
    numSamples = 500000
    smc, _, p = mcSampler.sample(numSamples=numSamples)
    assert smc.shape[1] >= numSamples
    ints = state_to_int(smc.reshape(-1, L))
    pmc = torch.zeros(16, dtype=torch.float64, device=ints.device).index_add_(0, ints, p[0])
    pmc = pmc / pmc.sum()
    assert float(torch.max(torch.abs(pmc - pex.reshape(-1)[:16]))) < 2e-3
    # the symmetrised amplitude is invariant under the orbit
    s = smc[:, :64]
    assert torch.allclose(psi(s), psi(torch.roll(s, 1, dims=-1)), atol=1e-12)
    assert torch.allclose(psi(s).real, psi(1 - s).real, atol=1e-12)
    s, psi_s, _ = mcSampler.sample(parameters=p0, numSamples=100)
    psi.set_parameters(p0)
    psi_s1 = psi(s)
    assert float(torch.max(torch.abs((psi_s - psi_s1) / psi_s))) < 1e-14


def test_symnet_ground_state_search():
    """reference tests/tdvp_test.py:27-56 with the RBM wrapped into SymNet (generic gradient path: dense O from
    jvmc_symrbm_grad, E_loc through s' enumeration): the ground state of the periodic TFIM lies in the fully
    symmetric sector, so the symmetrised ansatz must reach the same golden energies."""
    import warnings
    L, J = 4, -1.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        orbit = jVMC.util.symmetries.get_orbit_1D(L, "translation", "reflection", "spinflip")
    for k in (0, 1):
        hx, exE = REFG["gs_hx"][k], REFG["gs_energies"][k]
        psi = NQS(nets.sym_wrapper.SymNet(net=nets.CpxRBM(numHidden=6, bias=False), orbit=orbit))
        H = tfim(L, J, hx)
        exactSampler = sampler.ExactSampler(psi, L)
        tdvpEquation = jVMC.util.TDVP(exactSampler, snrTol=1, pinvTol=0.0, pinvCutoff=1e-8, rhsPrefactor=1.,
                                      diagonalShift=2, makeReal='real')
        ground_state_search(psi, H, tdvpEquation, exactSampler, numSteps=100, stepSize=5e-2)
        obs = measure({"energy": H}, psi, exactSampler)
        assert float(torch.max(torch.abs((obs['energy']['mean'] - exE) / exE))) < 1e-3


def test_cnn_ground_state_search_and_sampling():
    """nets.CNN through the public API: SR ground-state search with exact sampling reaches the golden TFIM energy
    (positive ground state, so a real log-amplitude suffices), then MCMC with the same net reproduces the exact
    distribution (reference tests/sampler_test.py style, 2e-3 on the histogram)."""
    L, J, hx, exE = 4, -1.0, REFG["gs_hx"][1], REFG["gs_energies"][1]
    psi = NQS(nets.CNN(F=(4,), channels=(3, 2), strides=(1,), bias=True, firstLayerBias=True), seed=3)
    H = tfim(L, J, hx)
    exactSampler = sampler.ExactSampler(psi, L)
    assert not psi.holomorphic and psi.realParams and psi.numParameters == 3 + 4 * 3 + 2 + 4 * 3 * 2
    tdvpEquation = jVMC.util.TDVP(exactSampler, snrTol=1, pinvTol=0.0, pinvCutoff=1e-8, rhsPrefactor=1.,
                                  diagonalShift=1., makeReal='real')
    e_start = float(measure({"energy": H}, psi, exactSampler)["energy"]["mean"][0])
    ground_state_search(psi, H, tdvpEquation, exactSampler, numSteps=200, stepSize=5e-2)
    e = float(measure({"energy": H}, psi, exactSampler)["energy"]["mean"][0])
    assert e < e_start and abs((e - exE) / exE) < 5e-3, (e_start, e, exE)
    # parameter round trip and dict/flat gradient consistency (reference tests/vqs_test.py:272-289)
    p = psi.get_parameters()
    psi.set_parameters(p)
    assert torch.equal(psi.get_parameters(), p)
    s = exactSampler.basis
    g, gd, mp = psi.gradients(s), psi.gradients_dict(s), psi.grad_dict_to_vec_map()
    for m in gd:
        for k in gd[m]:
            assert torch.equal(g[..., mp[m][k]], gd[m][k])
    # MCMC vs exact
    mcSampler = sampler.MCSampler(psi, (L,), 0, updateProposer=sampler.propose_spin_flip, numChains=777)
    _, _, pex = exactSampler.sample()
    smc, lp, pw = mcSampler.sample(numSamples=300000)
    ints = state_to_int(smc.reshape(-1, L))
    pmc = torch.zeros(16, dtype=torch.float64, device=ints.device).index_add_(0, ints, pw[0])
    assert float(torch.max(torch.abs(pmc / pmc.sum() - pex.reshape(-1)[:16]))) < 3e-3
    assert torch.allclose(lp, psi(smc))


def test_cnn_heisenberg_exchange_sampling_config4_style():
    """Config-4-shaped run at small size: 2-D Heisenberg (Marshall-rotated so that the ground state is positive),
    CNN ansatz, magnetisation-conserving exchange proposer, SR steps with MC sampling and batched local energies."""
    L = 4
    H = op.BranchFreeOperator(ElocBatchSize=500)
    for x in range(L):
        for y in range(L):
            i = x * L + y
            for j in (x * L + (y + 1) % L, ((x + 1) % L) * L + y):
                H.add(op.scal_opstr(-0.25, (op.Sx(i), op.Sx(j))))
                H.add(op.scal_opstr(-0.25, (op.Sy(i), op.Sy(j))))
                H.add(op.scal_opstr(0.25, (op.Sz(i), op.Sz(j))))
    psi = NQS(nets.CNN(F=(2, 2), channels=(4, 2), strides=(1, 1), bias=True, firstLayerBias=False), seed=5)
    neel = np.indices((L, L)).sum(0) % 2
    mc = sampler.MCSampler(psi, (L, L), 123, updateProposer=sampler.propose_spin_flip_zeroMag, numChains=200,
                           numSamples=4000, thermalizationSweeps=10, sweepSteps=L * L, initState=neel)
    tdvpEquation = jVMC.util.TDVP(mc, snrTol=1, pinvTol=1e-8, pinvCutoff=1e-8, rhsPrefactor=1., diagonalShift=0.1,
                                  makeReal='real')
    s, lp, p = mc.sample()
    assert s.shape[2:] == (L, L) and bool((s.reshape(s.shape[1], -1).sum(1) == L * L // 2).all())   # sector conserved
    e0 = float(measure({"energy": H}, psi, mc)["energy"]["mean"][0])
    ground_state_search(psi, H, tdvpEquation, mc, numSteps=40, stepSize=2e-2)
    e1 = float(measure({"energy": H}, psi, mc)["energy"]["mean"][0])
    # exact ground-state energy of the 4x4 periodic Heisenberg model: -0.70178 per site (= -11.2285)
    assert e1 < e0 - 0.5 and e1 > -11.2285 - 0.3, (e0, e1)


def test_sampler_output_contract():
    """time-major / chain-minor order, rounding up per chain, persistent chains."""
    L = 6
    psi = NQS(nets.CpxRBM(numHidden=4, bias=True))
    smp = sampler.MCSampler(psi, (L,), 123, updateProposer=sampler.propose_spin_flip_Z2, numChains=500,
                            sweepSteps=L, thermalizationSweeps=25, numSamples=4096)
    s, lp, p = smp.sample()
    assert s.shape == (1, 4500, L) and s.dtype == torch.int32 and lp.shape == (1, 4500) and p.shape == (1, 4500)
    assert smp.get_last_number_of_samples() == 4500
    assert torch.allclose(p.sum(), torch.tensor(1.0, dtype=torch.float64, device=p.device))
    assert torch.allclose(p, torch.full_like(p, 1 / 4500))
    # the last emitted sweep is the persistent chain state
    assert torch.equal(s[0, -500:], smp.states[0])
    with pytest.raises(ValueError):
        sampler.MCSampler(psi, (L,), 0, updateProposer=sampler.propose_spin_flip, mu=2.5)


def test_exact_sampler():
    """reference tests/sampler_test.py:563-593."""
    L = 4
    psi = NQS(nets.CpxRBM(numHidden=2, bias=False))
    exactSampler = sampler.ExactSampler(psi, L)
    p0 = psi.get_parameters()
    psi.set_parameters(WEIGHTS)
    s, psi_s, pex = exactSampler.sample()
    assert float(torch.max(torch.abs((psi(s) - psi_s) / psi_s))) < 1e-14
    assert abs(float(pex.sum()) - 1) < 1e-14
    s, psi_s, pex = exactSampler.sample(parameters=p0)
    psi.set_parameters(p0)
    assert float(torch.max(torch.abs((psi(s) - psi_s) / psi_s))) < 1e-14


# ------------------------------------------------------------------ tests/operator_test.py
def test_nonzeros_and_prefactor_arguments():
    """reference tests/operator_test.py:48-93."""
    L = 4
    s = torch.as_tensor(np.random.default_rng(3).integers(0, 2, (1, 24, L)).astype(np.int32)).cuda()
    h = op.BranchFreeOperator()
    h += 2. * op.Sp(0)
    h += 2. * op.Sp(1)
    h += 2. * op.Sp(2)
    sp, matEl = h.get_s_primes(s)
    logPsi = torch.ones(s.shape[:-1], dtype=torch.complex128, device="cuda")
    logPsiSP = torch.ones(sp.shape[:-1], dtype=torch.complex128, device="cuda")
    tmp = h.get_O_loc_unbatched(logPsi, logPsiSP)
    assert float(torch.sum(torch.abs(tmp - 2. * torch.sum(-(s[..., :3] - 1), dim=-1).to(torch.float64)))) < 1e-7

    def f(t):
        return 2.0 * t
    ht = op.BranchFreeOperator()
    for i in range(3):
        ht.add(op.scal_opstr(f, (op.Sp(i),)))
    for t in [0.5, 2, 13.9]:
        sp, matEl = ht.get_s_primes(s, t)
        logPsiSP = torch.ones(sp.shape[:-1], dtype=torch.complex128, device="cuda")
        tmp = ht.get_O_loc_unbatched(logPsi, logPsiSP)
        assert float(torch.sum(torch.abs(tmp - f(t) * torch.sum(-(s[..., :3] - 1), dim=-1).to(torch.float64)))) < 1e-7


def test_op_2d_shapes():
    """reference tests/operator_test.py:95-110."""
    L = 4
    s = torch.as_tensor(np.random.default_rng(3).integers(0, 2, (1, 24, L, L)).astype(np.int32)).cuda()
    h = op.BranchFreeOperator()
    h += 0.3 * op.Sp(0)
    h += 1.1 * op.Sp(1)
    h += 0.15 * op.Sp(4)
    sp, matEl = h.get_s_primes(s)
    assert tuple(sp.shape[2:]) == (L, L) and sp.shape[1] == 24 * matEl.shape[2]


def test_batched_Oloc():
    """reference tests/operator_test.py:112-162: batched (13) == unbatched == fused."""
    L = 4
    h, hb = op.BranchFreeOperator(), op.BranchFreeOperator(ElocBatchSize=13)
    for i in range(L):
        for o in (h, hb):
            o.add(op.scal_opstr(2., (op.Sx(i),)))
            o.add(op.scal_opstr(2., (op.Sy(i), op.Sz((i + 1) % L))))
    psi = NQS(nets.CpxRBM(numHidden=2, bias=False))
    mcSampler = sampler.MCSampler(psi, (L,), 0, updateProposer=sampler.propose_spin_flip, numChains=1)
    s, logPsi, _ = mcSampler.sample(numSamples=100)
    sp, matEl = h.get_s_primes(s)
    Oloc1 = h.get_O_loc_unbatched(logPsi, psi(sp))
    Oloc2 = h.get_O_loc_batched(s, psi, logPsi, 13)
    Oloc3 = h.get_O_loc(s, psi, logPsi)     # fused kernel
    assert float(torch.abs(torch.sum(Oloc1) - torch.sum(Oloc2))) < 1e-5
    assert torch.allclose(Oloc1, Oloc2, rtol=1e-12, atol=1e-14) and torch.allclose(Oloc1, Oloc3, rtol=1e-10, atol=1e-13)


def test_fermionic_operators_generic_path():
    """reference tests/operator_test.py:174-212: {c_i, c_j^dagger} = delta_ij (generic s' path, JW signs)."""
    L = 2
    psi = NQS(nets.CpxRBM(numHidden=2, bias=True))
    smp = sampler.ExactSampler(psi, (L,))

    def commutator(i, j):
        Comm = op.BranchFreeOperator()
        Comm.add(op.scal_opstr(1., (op.creation(j), op.annihilation(i))))
        Comm.add(op.scal_opstr(1., (op.annihilation(i), op.creation(j))))
        return Comm
    out = measure({"same_site": [commutator(0, 0), commutator(1, 1)],
                   "distinct_site": [commutator(0, 1), commutator(1, 0)]}, psi, smp)
    vals = torch.cat((out["same_site"]['mean'], out["distinct_site"]['mean']))
    assert torch.allclose(vals, torch.tensor([1., 1., 0., 0.], dtype=torch.float64, device=vals.device), atol=1e-14)
    var = torch.cat((out["same_site"]['variance'], out["distinct_site"]['variance']))
    assert float(var.abs().max()) < 1e-14


# ------------------------------------------------------------------ tests/stats_test.py
def test_sampled_obs():
    """reference tests/stats_test.py:15-38."""
    Obs1Loc = torch.tensor([[1., 2., 3.]], device="cuda", dtype=torch.float64)
    Obs2Loc = torch.tensor([[[1., 4.], [2., 5.], [3., 7.]]], device="cuda", dtype=torch.float64)
    p = torch.ones((1, 3), device="cuda", dtype=torch.float64) / 3
    obs1, obs2 = SampledObs(Obs1Loc, p), SampledObs(Obs2Loc, p)
    assert abs(float(obs1.mean()[0]) - 2.) < 1e-12 and abs(float(obs1.var()[0]) - 2. / 3) < 1e-12
    dd = dict(device="cuda", dtype=torch.float64)
    assert torch.allclose(obs2.covar(), torch.tensor([[2. / 3, 1], [1., 14. / 9]], **dd))
    assert torch.allclose(obs2.mean(), torch.tensor([2, 16. / 3], **dd))
    assert torch.allclose(obs1.covar(obs2), torch.tensor([2. / 3, 1.], **dd))
    assert torch.allclose(obs1.covar(obs2), obs1.covar_data(obs2).mean())
    assert torch.allclose(obs1.covar_var(obs2), obs1.covar_data(obs2).var())
    O = obs2._data.reshape(-1, 2)
    assert torch.allclose(obs2.tangent_kernel(), O @ O.conj().T)


def test_subset_function():
    """reference tests/stats_test.py:85-106."""
    N = 10
    Obs1 = torch.arange(N, dtype=torch.float64, device="cuda").reshape(1, N, 1)
    p = torch.rand((1, N), dtype=torch.float64, device="cuda")
    p = p / mpi.global_sum(p)
    obs1 = SampledObs(Obs1, p)
    obs2 = obs1.subset(0, N // 2)
    assert torch.allclose(obs1.mean()[0], mpi.global_sum(Obs1.reshape(1, N) * p))
    assert torch.allclose(obs2.mean(), mpi.global_sum(Obs1.reshape(1, N)[:, :N // 2] * p[:, :N // 2])
                          / mpi.global_sum(p[:, :N // 2]))
    obs3 = SampledObs(Obs1[:, :N // 2, :], p[:, :N // 2] / mpi.global_sum(p[:, :N // 2]))
    assert torch.allclose(obs3.covar(), obs2.covar())


@pytest.mark.parametrize("bias", [False, True])
def test_rbm_gradient_obs_equals_generic(bias):
    """The factorised observable reproduces the generic SampledObs on the materialised gradients."""
    L, M = 5, 6
    psi = NQS(nets.CpxRBM(numHidden=M, bias=bias), seed=3)
    s = torch.as_tensor(np.random.default_rng(1).integers(0, 2, (1, 40, L)).astype(np.int32)).cuda()
    psi(s)
    W, b = orbm.init_o1(L, M, bias, 5)
    psi.set_parameters(torch.as_tensor(orbm.flatten_params(W, b)))
    p = torch.rand((1, 40), dtype=torch.float64, device="cuda")
    p = p / p.sum()
    E = torch.randn(1, 40, dtype=torch.complex128, device="cuda")
    gen, fac, eo = SampledObs(psi.gradients(s), p), RBMGradientObs(psi, s, p), SampledObs(E, p)
    assert torch.allclose(fac.mean(), gen.mean(), rtol=1e-12, atol=1e-15)
    S_ref = gen.covar()
    assert float((fac.covar() - S_ref).abs().max()) < 1e-10 * float(S_ref.abs().max())   # norm-wise 1e-10 (north_star)
    assert torch.allclose(fac.covar(eo), gen.covar(eo), rtol=1e-10, atol=1e-14)
    assert torch.allclose(eo.covar(fac), eo.covar(gen), rtol=1e-10, atol=1e-14)
    assert torch.allclose(fac.var(), gen.var().reshape(-1), rtol=1e-10, atol=1e-14)
    assert torch.allclose(fac._data, gen._data, rtol=1e-12, atol=1e-15)
    assert torch.allclose(fac.tangent_kernel(), gen.tangent_kernel(), rtol=1e-10, atol=1e-14)
    sub_f, sub_g = fac.subset(start=1, step=2), gen.subset(start=1, step=2)
    assert float((sub_f.covar() - sub_g.covar()).abs().max()) < 1e-9 * float(sub_g.covar().abs().max())
    x = torch.randn(40, dtype=torch.complex128, device="cuda")
    ref = -gen._data.reshape(40, -1).conj().T @ x
    assert torch.allclose(fac.minsr_contract(x), ref, rtol=1e-10, atol=1e-13)
    u = torch.randn(gen._data.shape[-1], dtype=torch.float64, device="cuda")
    S0 = gen.covar()
    assert torch.allclose(fac.quad_form(u), (u.to(S0.dtype) @ (S0 @ u.to(S0.dtype))).real, rtol=1e-9)


# ------------------------------------------------------------------ tests/vqs_test.py
@pytest.mark.parametrize("net", ["cpx", "real"])
def test_gradients_finite_difference(net):
    """reference tests/vqs_test.py:99-164: every flat index."""
    L = 3
    psiC = NQS(nets.CpxRBM(numHidden=2, bias=True) if net == "cpx" else nets.RBM(numHidden=2, bias=True))
    s = torch.zeros((1, 4, L), dtype=torch.int32, device="cuda")
    s[..., 0, 1] = 1
    s[..., 2, 2] = 1
    psi0 = psiC(s)
    assert psiC.holomorphic == (net == "cpx")
    G = psiC.gradients(s)
    delta = 1e-6
    params = psiC.get_parameters()
    assert G.shape[-1] == params.shape[0]
    for j in range(G.shape[-1]):
        u = torch.zeros(G.shape[-1], dtype=torch.float64, device="cuda")
        u[j] = 1
        psiC.update_parameters(delta * u)
        psi1 = psiC(s)
        psiC.set_parameters(params)
        Gfd = (psi1 - psi0) / delta
        assert float(torch.max(torch.abs(Gfd - G[..., j]))) < 1e-4


def test_gradient_dict_and_param_roundtrip():
    """reference tests/vqs_test.py:272-289 + parameter (un)flattening (jVMC/vqs.py:430-466)."""
    psi = NQS(nets.CpxRBM(numHidden=8, bias=False), seed=1234)
    s = torch.zeros((1, 3, 4), dtype=torch.int32, device="cuda")
    psi(s)
    g1 = psi.gradients(s)
    g2 = psi.gradients_dict(s)["Dense_0"]["kernel"]
    assert float(torch.linalg.norm(g1 - g2)) == 0.0
    P = psi.get_parameters()
    assert P.shape == (2 * 4 * 8,) and psi.numParameters == 32
    psi.set_parameters(P * 2)
    assert torch.allclose(psi.get_parameters(), 2 * P)
    psi.set_parameters(WEIGHTS.repeat(4))
    W = psi.params["Dense_0"]["kernel"]
    w = WEIGHTS.repeat(4).to(W.device)
    assert torch.allclose(W, (w[:32] + 1j * w[32:]).reshape(4, 8))
    psi2 = NQS(nets.CpxRBM(numHidden=8))
    with pytest.raises(RuntimeError):
        psi2.set_parameters(P)


# ------------------------------------------------------------------ tests/nets_test.py, vqs_test.py, tdvp_test.py with wrappers
def _orbit_1d(L, *args):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return jVMC.util.symmetries.get_orbit_1D(L, *args)


def test_cnn_translation_invariance_1d_2d():
    """reference tests/nets_test.py:16-46: a periodic CNN gives the same output on all translates."""
    cnn = NQS(nets.CNN(F=(4,), channels=[3, 2, 5]), seed=0)
    S0 = np.pad(np.array([1, 0, 1, 1, 0]), (0, 4), 'wrap')
    S = torch.tensor(np.array([S0[i:i + 5] for i in range(5)]), dtype=torch.int32)[None]
    psiS = cnn(S)
    assert float(torch.max(torch.abs(psiS - psiS[0, 0]))) < 1e-12
    cnn = NQS(nets.CNN(F=(3, 3), channels=[3, 2, 5], strides=[1, 1]), seed=0)
    S0 = np.pad(np.array([[1, 0, 1, 1], [0, 1, 1, 1], [0, 0, 1, 0], [1, 0, 0, 1]]), [(0, 3), (0, 3)], 'wrap')
    S = torch.tensor(np.array([S0[i:i + 4, j:j + 4] for i in range(4) for j in range(4)]), dtype=torch.int32)[None]
    psiS = cnn(S)
    assert float(torch.max(torch.abs(psiS - psiS[0, 0]))) < 1e-12


def test_sym_net_translation_invariance():
    """reference tests/nets_test.py:51-65: real RBM wrapped into SymNet with the translation orbit."""
    L = 5
    psi = NQS(nets.SymNet(net=nets.RBM(numHidden=5), orbit=_orbit_1d(L, "translation")), seed=0)
    S0 = np.pad(np.array([1, 0, 1, 1, 0]), (0, 4), 'wrap')
    S = torch.tensor(np.array([S0[i:i + 5] for i in range(5)]), dtype=torch.int32)[None]
    psiS = psi(S)
    assert not psi.holomorphic
    assert float(torch.max(torch.abs(psiS - psiS[0, 0]))) < 1e-12


def test_holomorphicity_recognition_and_gradients_symnet():
    """reference tests/vqs_test.py:86-131 as written there (the RBM sits inside SymNet with the trivial orbit)."""
    L = 3
    for k in range(6):
        net = nets.sym_wrapper.SymNet(net=nets.CpxRBM(numHidden=2 ** k, bias=True), orbit=_orbit_1d(L))
        psiC = NQS(net)
        psiC(torch.zeros((1, 4, 3), dtype=torch.int32))
        assert psiC.holomorphic
    psiC = NQS(nets.sym_wrapper.SymNet(net=nets.CpxRBM(numHidden=2, bias=True), orbit=_orbit_1d(L, "translation")))
    s = torch.zeros((1, 4, L), dtype=torch.int32, device="cuda")
    s[..., 0, 1] = 1
    s[..., 2, 2] = 1
    psi0 = psiC(s)
    G = psiC.gradients(s)
    delta = 1e-6
    params = psiC.get_parameters()
    assert G.shape[-1] == params.shape[0]
    for j in range(G.shape[-1]):
        u = torch.zeros(G.shape[-1], dtype=torch.float64, device="cuda")
        u[j] = 1
        psiC.update_parameters(delta * u)
        psi1 = psiC(s)
        psiC.set_parameters(params)
        assert float(torch.max(torch.abs((psi1 - psi0) / delta - G[..., j]))) < 1e-4


def test_snr_consistency_symnet():
    """reference tests/tdvp_test.py:201-263: legacy SNR formula vs SampledObs.covar_data().transform().var() for the
    translation-symmetrised RBM sampled by MCMC."""
    L = 4
    psi = NQS(nets.sym_wrapper.SymNet(net=nets.CpxRBM(numHidden=2, bias=False), orbit=_orbit_1d(L, "translation")),
              batchSize=5000)
    psi(torch.tensor([[[1, 1, 1, 1]]], dtype=torch.int32))
    psi.set_parameters(WEIGHTS)
    H = tfim(L, -1.0, -0.3)
    smp = sampler.MCSampler(psi, (L,), 0, numSamples=10, updateProposer=sampler.propose_spin_flip, mu=2, numChains=500)
    s, logPsi, p = smp.sample()
    Eloc_old = H.get_O_loc(s, psi, logPsi, 0.0)
    Eloc = SampledObs(Eloc_old, p)
    grads_old = psi.gradients(s)
    grads = SampledObs(grads_old, p)
    S = jVMC.util.imagFun(grads.covar())
    ev, V = torch.linalg.eigh(S)
    # legacy formula
    E0 = Eloc_old - mpi.global_mean(Eloc_old, p)
    G0 = grads_old - mpi.global_mean(grads_old, p)
    EOdata = (p * E0)[..., None] * torch.conj(G0) * mpi.globNumSamples
    EOdata = jVMC.util.imagFun(EOdata).to(V.dtype) @ torch.conj(V)
    rhoVar_old = mpi.global_variance(EOdata, torch.ones(EOdata.shape[:2], dtype=torch.float64, device=EOdata.device)
                                     / mpi.globNumSamples)
    EO = grads.covar_data(Eloc).transform(linearFun=torch.conj(V).T, nonLinearFun=lambda x: jVMC.util.imagFun(x))
    rhoVar_new = EO.var().ravel()
    assert torch.allclose(rhoVar_old.ravel().real.to(torch.float64), rhoVar_new.real.to(torch.float64), rtol=1e-8,
                          atol=1e-12 * float(rhoVar_new.abs().max()))


def test_two_nets_gradients_and_sampling():
    """reference tests/vqs_test.py:133-164 (finite-difference gradients of NQS((rbm1, rbm2))) and
    tests/sampler_test.py:184-228 (MCMC with the two-network ansatz: the amplitude network alone is sampled)."""
    L = 3
    psi = NQS((nets.RBM(numHidden=2, bias=True), nets.RBM(numHidden=3, bias=True)))
    s = torch.zeros((1, 4, L), dtype=torch.int32, device="cuda")
    s[..., 0, 1] = 1
    s[..., 2, 2] = 1
    psi0 = psi(s)
    assert not psi.holomorphic and psi.realParams
    G = psi.gradients(s)
    delta = 1e-5
    params = psi.get_parameters()
    assert G.shape[-1] == params.shape[0] == (2 + 3 * 2) + (3 + 3 * 3)
    for j in range(G.shape[-1]):
        u = torch.zeros(G.shape[-1], dtype=torch.float64, device="cuda")
        u[j] = 1
        psi.update_parameters(delta * u)
        psi1 = psi(s)
        psi.set_parameters(params)
        assert float(torch.max(torch.abs((psi1 - psi0) / delta - G[..., j]))) < 1e-3
    # sampler
    L = 4
    psi = NQS(nets.two_nets_wrapper.TwoNets((nets.RBM(numHidden=2, bias=False), nets.RBM(numHidden=2, bias=False))))
    exactSampler = sampler.ExactSampler(psi, L)
    mcSampler = sampler.MCSampler(psi, (L,), 0, updateProposer=sampler.propose_spin_flip, numChains=777)
    psi.set_parameters(torch.cat([WEIGHTS[:8], WEIGHTS[:8]]))
    _, _, pex = exactSampler.sample()
    smc, lp, p = mcSampler.sample(numSamples=500000)
    assert smc.shape[1] >= 500000
    ints = state_to_int(smc.reshape(-1, L))
    pmc = torch.zeros(16, dtype=torch.float64, device=ints.device).index_add_(0, ints, p[0])
    assert float(torch.max(torch.abs(pmc / pmc.sum() - pex.reshape(-1)[:16]))) < 2e-3
    assert float(lp.imag.abs().max()) > 0        # the phase network contributes


@pytest.mark.parametrize("makeReal,x,shift,bias", [('real', 1.0, 0.0, False), ('real', 1.0, 3.0, True),
                                                  ('imag', 1.j, 0.0, False), ('imag', 1.j, 0.0, True)])
def test_tdvp_pc_level_solve_matches_reference_layout(makeReal, x, shift, bias):
    """TDVP.solve through the P_c x P_c complex eigen-decomposition vs the reference's doubled P x P layout on the
    same samples: spectrum, forces, solver outputs; eigenvector matrix V (built on demand) diagonalises q(S0)."""
    L = 6
    psi = NQS(nets.CpxRBM(numHidden=5, bias=bias), seed=7)
    psi(torch.zeros((1, 1, L), dtype=torch.int32))
    rng = np.random.default_rng(5)
    n = psi.get_parameters().shape[0]
    psi.set_parameters(torch.as_tensor(0.4 * rng.standard_normal(n)))
    H = tfim(L, -1.0, -0.8)
    for exact in (True, False):
        smp = sampler.ExactSampler(psi, L) if exact else \
            sampler.MCSampler(psi, (L,), 3, numSamples=4000, updateProposer=sampler.propose_spin_flip, numChains=200)
        s, logPsi, p = smp.sample()
        Eloc = SampledObs(H.get_O_loc(s, psi, logPsi), p)
        out = []
        for pc in (True, False):
            td = jVMC.util.TDVP(smp, snrTol=2, pinvTol=1e-8, pinvCutoff=1e-8, rhsPrefactor=x, makeReal=makeReal,
                                diagonalShift=shift)
            td.pcLevel = pc
            upd, res, cut = td.solve(Eloc, RBMGradientObs(psi, s, p))
            out.append((td, upd, res, cut))
        (a, ua, ra, ca), (b, ub, rb, cb) = out
        scale = float(b.ev.abs().max())
        assert torch.allclose(a.ev, b.ev, atol=1e-10 * scale)
        assert torch.allclose(a.F0, b.F0, rtol=1e-12, atol=1e-14)
        assert torch.allclose(a.S.to(torch.complex128), b.S.to(torch.complex128), atol=1e-12 * scale)
        # the doubled eigenbasis reconstructed from the eigenvectors of A is orthonormal and diagonalises q(S0)
        V = a.V.to(torch.complex128)
        S = b.S.to(torch.complex128)
        assert torch.allclose(V.conj().T @ V, torch.eye(V.shape[0], dtype=V.dtype, device=V.device), atol=1e-10)
        assert torch.allclose(V.conj().T @ S @ V, torch.diag(a.ev).to(V.dtype), atol=1e-9 * scale)
        # basis-independent quantities: |VtF|^2 and rhoVar summed over each (possibly degenerate) eigenvalue cluster
        assert np.isclose(float((a.VtF.abs() ** 2).sum()), float((b.VtF.abs() ** 2).sum()), rtol=1e-9)
        assert np.isclose(float(a.rhoVar.sum()), float(b.rhoVar.sum()), rtol=1e-8)
        if makeReal == 'imag' or exact:
            # unique eigenvectors ('imag': +-lambda) or no SNR weighting (exact sampler): identical update
            assert float(cb) == float(ca)
            assert torch.allclose(ua, ub, rtol=1e-6, atol=1e-8 * float(ub.abs().max()))
            assert np.isclose(float(ra), float(rb), rtol=1e-6, atol=1e-10)
        else:
            # 'real' with an MC sampler: every eigenvalue of q(S0) is doubly degenerate and the SNR weighting is active.
            # The P_c-level solve regularises with the eigenspace-invariant SNR; the same rule applied to the
            # reference-layout decomposition (whatever basis cuSOLVER returned inside the pairs) gives the same update.
            from vmc_jax_b200 import kernels as K, mpi_wrapper as mpi
            vtf2 = (b.VtF.abs() ** 2).reshape(-1, 2).sum(1)
            rvp = b.rhoVar.reshape(-1, 2).sum(1)
            snr_pair = torch.sqrt(torch.abs(mpi.globNumSamples * vtf2 / rvp))
            ok = b.ev.reshape(-1, 2)[:, 0] > 1e-9 * scale          # 0/0 in the numerical null space of S
            assert int(ok.sum()) > 4
            assert torch.allclose(a.snr.reshape(-1, 2)[:, 0][ok], snr_pair[ok], rtol=1e-5)
            assert torch.allclose(a.snr.reshape(-1, 2)[:, 1][ok], snr_pair[ok], rtol=1e-5)
            F = realFunOrImag(b, makeReal)
            pinvEv, scal = K.tdvp_regularize(b.ev, b.VtF, torch.repeat_interleave(snr_pair, 2), F.to(torch.complex128),
                                             1e-8, 1e-8, 2.0)
            uref = torch.mv(b.V.to(torch.complex128), pinvEv * b.VtF).real
            assert float(scal[1]) == float(ca)
            assert torch.allclose(ua, uref, rtol=1e-6, atol=1e-8 * float(uref.abs().max()))
            assert np.isclose(float(ra), float(scal[0]), rtol=1e-6, atol=1e-10)


def realFunOrImag(td, makeReal):
    """q(F0) of a solved TDVP object"""
    return td.makeReal(td.F0)


def test_tdvp_snr_second_moments_by_gram_kernel():
    """'imag' mode with N_s >= 2 P_c >= 512: the SNR second moments come from a second Gram matrix
    (RBMGradientObs.weighted_second_moment) instead of per-sample projections; same rhoVar / SNR / update as the
    reference-layout path."""
    L, M = 10, 26
    psi = NQS(nets.CpxRBM(numHidden=M, bias=False), seed=3)
    psi(torch.zeros((1, 1, L), dtype=torch.int32))
    rng = np.random.default_rng(11)
    psi.set_parameters(torch.as_tensor(0.3 * rng.standard_normal(psi.get_parameters().shape[0])))
    H = tfim(L, -1.0, -0.8)
    smp = sampler.MCSampler(psi, (L,), 3, numSamples=3000, updateProposer=sampler.propose_spin_flip, numChains=300)
    s, logPsi, p = smp.sample()
    assert s.shape[1] >= 2 * L * M
    Eloc = SampledObs(H.get_O_loc(s, psi, logPsi), p)
    out = []
    for pc in (True, False):
        td = jVMC.util.TDVP(smp, snrTol=2, pinvTol=1e-8, pinvCutoff=1e-8, rhsPrefactor=1.j, makeReal='imag')
        td.pcLevel = pc
        upd, res, cut = td.solve(Eloc, RBMGradientObs(psi, s, p))
        out.append((td, upd, res, cut))
    (a, ua, ra, ca), (b, ub, rb, cb) = out
    assert torch.allclose(a.ev, b.ev, atol=1e-10 * float(b.ev.abs().max()))
    big = b.rhoVar > 1e-8 * b.rhoVar.max()
    assert torch.allclose(a.rhoVar[big], b.rhoVar[big], rtol=1e-6)
    assert torch.allclose(a.snr[big], b.snr[big], rtol=1e-5)
    assert float(ca) == float(cb)
    assert torch.allclose(ua, ub, rtol=1e-5, atol=1e-7 * float(ub.abs().max()))


def test_nqs_orbit_argument_wraps_symnet():
    """NQS(net, orbit=...) wraps the net into SymNet itself (reference vqs.py:170-173)."""
    L = 4
    orbit = _orbit_1d(L, "translation", "reflection")
    a = NQS(nets.CpxRBM(numHidden=3, bias=True), orbit=orbit, seed=5)
    b = NQS(nets.sym_wrapper.SymNet(net=nets.CpxRBM(numHidden=3, bias=True), orbit=orbit), seed=5)
    s = torch.randint(0, 2, (1, 9, L), dtype=torch.int32, device="cuda")
    assert a.kind == "symrbm" and torch.equal(a(s), b(s)) and torch.equal(a.gradients(s), b.gradients(s))
    assert torch.equal(a.get_parameters(), b.get_parameters())
