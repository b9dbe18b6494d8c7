"""Pins the CPU oracle against the reference's own golden vectors / known answers
(SURVEY 8c).  CPU only; the oracle is the checker for the CUDA path in the -m gpu tests."""
import json
import os

import numpy as np
import pytest

from oracle import rbm, bfo, stats, solve, sampling, stepper

GOLD = os.path.join(os.path.dirname(__file__), "golden")
with open(os.path.join(GOLD, "reference_goldens.json")) as f:
    REFG = json.load(f)


def tfim(L, J, hx):
    return bfo.Tables(bfo.tfim_strings((L,), hx, J))


def zz_op(L):
    return bfo.Tables([(1.0, [bfo.Sz(l), bfo.Sz((l + 1) % L)]) for l in range(L)])


def measure(tab, W, b, basis, p, *args):
    f = lambda s: rbm.cpx_rbm_logpsi(s, W, b)
    O = bfo.get_O_loc(tab, basis, f, *args)
    return np.real(np.sum(p * O))


def exact_state(W, b, N):
    basis = sampling.basis_states(N)
    lp = rbm.cpx_rbm_logpsi(basis, W, b)
    p, _ = sampling.exact_probabilities(lp)
    return basis, lp, p


def test_param_layout_discriminator():
    # reference tests/tdvp_test.py:123: <ZZ>(t=0) = 0.88276213 only with W=(w[:8]+i w[8:]).reshape(N,M)
    W, b = rbm.unflatten_params(np.array(REFG["rbm_weights"]), 4, 2)
    basis, lp, p = exact_state(W, b, 4)
    assert abs(measure(zz_op(4), W, b, basis, p) - REFG["zz_trajectory"][0]) < 1e-8
    assert np.allclose(rbm.flatten_params(W, b), REFG["rbm_weights"])


def _tdvp_rhs_exact(tdvp, ham, N, M, bias):
    basis = sampling.basis_states(N)

    def f(y, t, intStep=0):
        W, b = rbm.unflatten_params(y, N, M, bias)
        lp = rbm.cpx_rbm_logpsi(basis, W, b)
        p, _ = sampling.exact_probabilities(lp)
        E = bfo.get_O_loc(ham, basis, lambda s: rbm.cpx_rbm_logpsi(s, W, b), t, logPsiS=lp)
        upd, _, _ = solve.tdvp_rhs(W, b, basis, p, E, tdvp, 2 ** N, rbm.gradients_holomorphic)
        if intStep == 0:
            f.E0, f.var0 = np.real(tdvp.ElocMean), tdvp.ElocVar
        return upd
    return f


def test_time_evolution_golden():
    """reference tests/tdvp_test.py:60-128: 21 <ZZ>(t) values to 1e-3, energy drift < 1e-3."""
    L, J, hx = 4, -1.0, -0.3
    y = np.array(REFG["rbm_weights"])
    ham, ZZ = tfim(L, J, hx), zz_op(L)
    tdvp = solve.TDVP(snrTol=1, pinvTol=0.0, pinvCutoff=1e-8, rhsPrefactor=1.j, diagonalShift=0.,
                      makeReal='imag', exact_sampler=True)
    f = _tdvp_rhs_exact(tdvp, ham, L, 2, False)
    st = stepper.AdaptiveHeun(timeStep=1e-3, tol=1e-5)
    t, times, obs = 0.0, [0.0], []

    def meas(y, t):
        W, b = rbm.unflatten_params(y, L, 2)
        basis, lp, p = exact_state(W, b, L)
        return [measure(ham, W, b, basis, p, t), measure(ZZ, W, b, basis, p)]
    obs.append(meas(y, t))
    while t < 0.5:
        y, dt = st.step(f, y, 0)
        t += dt
        times.append(t)
        obs.append(meas(y, t))
    obs = np.array(obs)
    assert np.max(np.abs((obs[:, 0] - obs[0, 0]) / obs[0, 0])) < 1e-3
    refT = np.arange(0, 0.5, 0.05)
    zz = np.interp(refT, np.array(times), obs[:, 1])
    assert np.max(np.abs(zz - np.array(REFG["zz_trajectory"])[:len(zz)])) < 1e-3


@pytest.mark.parametrize("k", [0, 1])
def test_sr_ground_state_golden(k):
    """reference tests/tdvp_test.py:27-56 (ground_state_search: Euler 5e-2, shift*=0.95)."""
    L, hx, exE = 4, REFG["gs_hx"][k], REFG["gs_energies"][k]
    W, b = rbm.init_cpx_rbm(L, 6, False, 1234)
    y = rbm.flatten_params(W, b)
    ham = tfim(L, -1.0, hx)
    tdvp = solve.TDVP(snrTol=1, pinvTol=0.0, pinvCutoff=1e-8, rhsPrefactor=1., diagonalShift=2,
                      makeReal='real', exact_sampler=True)
    f = _tdvp_rhs_exact(tdvp, ham, L, 6, False)
    delta = 2.0
    for _ in range(100):
        y, _ = stepper.euler_step(f, y, 0, 5e-2)
        delta *= 0.95
        tdvp.diagonalShift = delta
    W, b = rbm.unflatten_params(y, L, 6)
    basis, lp, p = exact_state(W, b, L)
    assert abs((measure(ham, W, b, basis, p) - exE) / exE) < 1e-3


@pytest.mark.parametrize("k", [0, 1])
def test_minsr_ground_state_golden(k):
    """reference tests/minsr_test.py:15-43."""
    L, M, hx, exE = 4, 8, REFG["gs_hx"][k], REFG["gs_energies"][k]
    W, b = rbm.init_cpx_rbm(L, M, False, 1234)
    y = rbm.flatten_params(W, b).astype(np.complex128)
    ham = tfim(L, -1.0, hx)
    basis = sampling.basis_states(L)
    for _ in range(200):
        W, b = rbm.unflatten_params(y, L, M)
        lp = rbm.cpx_rbm_logpsi(basis, W, b)
        p, _ = sampling.exact_probabilities(lp)
        E = bfo.get_O_loc(ham, basis, lambda s: rbm.cpx_rbm_logpsi(s, W, b), 0, logPsiS=lp)
        upd = solve.minsr_solve(stats.SampledObs(E, p), stats.SampledObs(rbm.gradients_holomorphic(basis, W, b), p),
                                True, pinvTol=1e-6)
        y = y + 1e-2 * upd
        # set_parameters(P) then get_parameters(): complex flat vector collapses to real (vqs.py:437,462)
        y = rbm.flatten_params(*rbm.unflatten_params(y, L, M)).astype(np.complex128)
    W, b = rbm.unflatten_params(y, L, M)
    basis, lp, p = exact_state(W, b, L)
    assert abs((measure(ham, W, b, basis, p) - exE) / exE) < 1e-3


def test_fermion_hubbard_golden():
    """reference tests/operator_test.py:215-254: <H> = -9.95314531 on the stored target state."""
    t, mu, V, L, fl = -1.0, -2.0, 4.0, 4, 2
    n = fl * L
    up, do = 0, n - 1
    S = []
    for i in range(L):
        S.append((V, [bfo.number(up + i), bfo.number(do - i)]))
        S.append((mu, [bfo.number(up + i)]))
        S.append((mu, [bfo.number(do - i)]))
        if i == L - 1:
            continue
        S.append((t, [bfo.creation(up + i + 1), bfo.annihilation(up + i)]))
        S.append((t, [bfo.creation(up + i), bfo.annihilation(up + i + 1)]))
        S.append((t, [bfo.creation(do - i - 1), bfo.annihilation(do - i)]))
        S.append((t, [bfo.creation(do - i), bfo.annihilation(do - i - 1)]))
    tab = bfo.Tables(S)
    kernel = np.load(os.path.join(GOLD, "fermion_ref.npy"))

    def target(s):  # Target module, tests/operator_test.py:19-39 (state reversed for the index)
        idx = (s[:, ::-1].astype(np.int64) * (2 ** np.arange(n))[None, :]).sum(1)
        k = kernel[idx]
        return np.log(np.abs(k + 1e-15)) + 1j * np.angle(k)
    basis = sampling.basis_states(n)
    lp = target(basis)
    p, _ = sampling.exact_probabilities(lp)
    O = bfo.get_O_loc(tab, basis, target, logPsiS=lp)
    assert np.allclose(np.sum(p * O), REFG["fermion_energy"])


def test_anticommutators():
    """reference tests/operator_test.py:174-212."""
    W, b = rbm.init_o1(2, 2, True, 7)
    basis, lp, p = exact_state(W, b, 2)
    f = lambda s: rbm.cpx_rbm_logpsi(s, W, b)

    def comm(i, j):
        return bfo.Tables([(1., [bfo.creation(j), bfo.annihilation(i)]), (1., [bfo.annihilation(i), bfo.creation(j)])])
    vals = [np.sum(p * bfo.get_O_loc(comm(i, j), basis, f, logPsiS=lp)) for i, j in [(0, 0), (1, 1), (0, 1), (1, 0)]]
    assert np.allclose(np.real(vals), [1., 1., 0., 0.], rtol=1e-15, atol=1e-15)


def test_splus_known_answer_and_prefactors():
    """reference tests/operator_test.py:48-93."""
    rng = np.random.default_rng(3)
    s = rng.integers(0, 2, (24, 4)).astype(np.int32)
    tab = bfo.Tables([(2., [bfo.Sp(i)]) for i in range(3)])
    sp, m, cnt = bfo.get_s_primes(tab, s)
    O = bfo.o_loc(m, np.ones(24), np.ones(sp.shape[0]))
    assert np.sum(np.abs(O - 2. * np.sum(1 - s[:, :3], axis=-1))) < 1e-7
    tabt = bfo.Tables([(lambda t: 2.0 * t, [bfo.Sp(i)]) for i in range(3)])
    for t in [0.5, 2, 13.9]:
        sp, m, cnt = bfo.get_s_primes(tabt, s, t)
        O = bfo.o_loc(m, np.ones(24), np.ones(sp.shape[0]))
        assert np.sum(np.abs(O - 2.0 * t * np.sum(1 - s[:, :3], axis=-1))) < 1e-7


def test_compaction_padding_semantics():
    """jVMC/operator/base.py:91-114: ascending nonzero ops, padded with the last op's s', m=0."""
    tab = bfo.Tables([(2., [bfo.Sp(0)]), (3., [bfo.Sp(1)]), (5., [bfo.Sp(2)])])
    s = np.array([[0, 0, 0], [1, 0, 1], [1, 1, 1]], np.int32)
    sp, m, cnt = bfo.get_s_primes(tab, s)
    assert list(cnt) == [3, 1, 0] and m.shape == (3, 3)
    sp = sp.reshape(3, 3, 3)
    assert np.array_equal(m[1], [3., 0., 0.])
    assert np.array_equal(sp[1, 0], [1, 1, 1])
    # padding = last operator's s' (Sp(2) applied to [1,0,1] maps site 2: 1 -> 0)
    assert np.array_equal(sp[1, 1], [1, 0, 0]) and np.array_equal(sp[1, 2], [1, 0, 0])


def test_sampled_obs_integers():
    """reference tests/stats_test.py:15-38."""
    O1 = np.array([1., 2., 3.])
    O2 = np.array([[1., 4.], [2., 5.], [3., 7.]])
    p = np.ones(3) / 3
    o1, o2 = stats.SampledObs(O1, p), stats.SampledObs(O2, p)
    assert np.isclose(o1.mean()[0], 2.) and np.isclose(o1.var()[0], 2. / 3)
    assert np.allclose(o2.covar(), [[2. / 3, 1.], [1., 14. / 9]])
    assert np.allclose(o2.mean(), [2, 16. / 3])
    assert np.allclose(o1.covar(o2), [2. / 3, 1.])
    assert np.allclose(o1.covar(o2), o1.covar_data(o2).mean())
    assert np.allclose(o1.covar_var(o2), o1.covar_data(o2).var())
    assert np.allclose(o2.tangent_kernel(), o2._data @ o2._data.conj().T)


def test_subset_consistency():
    """reference tests/stats_test.py:85-106."""
    N = 10
    O = np.arange(N, dtype=np.float64)
    p = np.random.default_rng(123).uniform(size=N)
    p /= p.sum()
    o1 = stats.SampledObs(O, p)
    o2 = o1.subset(0, N // 2)
    assert np.allclose(o2.mean(), np.sum(O[:N // 2] * p[:N // 2]) / np.sum(p[:N // 2]))
    o3 = stats.SampledObs(O[:N // 2], p[:N // 2] / np.sum(p[:N // 2]))
    assert np.allclose(o3.covar(), o2.covar())


@pytest.mark.parametrize("bias", [False, True])
def test_gradients_finite_difference(bias):
    """reference tests/vqs_test.py:99-131: every flat index, holomorphic CpxRBM."""
    N, M = 3, 2
    W, b = rbm.init_o1(N, M, bias, 11)
    s = np.zeros((4, N), np.int32)
    s[0, 1] = 1
    s[2, 2] = 1
    G = rbm.gradients_holomorphic(s, W, b)
    y0 = rbm.flatten_params(W, b)
    assert G.shape[1] == y0.shape[0]
    psi0 = rbm.cpx_rbm_logpsi(s, W, b)
    d = 1e-6
    for j in range(G.shape[1]):
        y = y0.copy()
        y[j] += d
        W1, b1 = rbm.unflatten_params(y, N, M, bias)
        fd = (rbm.cpx_rbm_logpsi(s, W1, b1) - psi0) / d
        assert np.max(np.abs(fd - G[:, j])) < 1e-5


def test_gradients_real_rbm_finite_difference():
    N, M = 3, 2
    W, b = rbm.init_o1(N, M, True, 5, real=True)
    s = np.array([[0, 1, 0], [0, 0, 0], [0, 0, 1], [1, 1, 0]], np.int32)
    G = rbm.gradients_real(s, W, b)
    y0 = rbm.flatten_params(W, b)
    psi0 = rbm.real_rbm_logpsi(s, W, b)
    for j in range(G.shape[1]):
        y = y0.copy()
        y[j] += 1e-6
        W1, b1 = rbm.unflatten_params(y, N, M, True, real=True)
        assert np.max(np.abs((rbm.real_rbm_logpsi(s, W1, b1) - psi0) / 1e-6 - G[:, j])) < 1e-5


def test_mcmc_matches_exact_distribution():
    """reference tests/sampler_test.py:61-75 (bare CpxRBM instead of SymNet): histogram of the
    reference-faithful Metropolis sampler vs exact probabilities."""
    L = 4
    W, b = rbm.unflatten_params(np.array(REFG["rbm_weights"]), L, 2)
    basis, lp, pex = exact_state(W, b, L)
    smp = sampling.MCSampler(lambda s: np.real(rbm.cpx_rbm_logpsi(s, W, b)), L, numChains=777, seed=0)
    cfg, glob = smp.sample(200000)
    assert cfg.shape[0] == glob >= 200000
    ints = (cfg * (2 ** np.arange(L))[None, :]).sum(1)
    pmc = np.bincount(ints, minlength=16) / cfg.shape[0]
    assert np.max(np.abs(pmc - pex)) < 4e-3


def test_distribute_sampling_rounding():
    """jVMC/mpi_wrapper.py:87-95: config 1, 4096 requested, 500 chains -> 9 per chain, 4500."""
    assert sampling.distribute_sampling(4096, localDevices=1, numChainsPerDevice=500) == (9, 4500)
    assert sampling.distribute_sampling(10, commSize=3, rank=0) == (4, 10)
    assert sampling.distribute_sampling(10, commSize=3, rank=2) == (3, 10)


def test_adaptive_heun_linear_ode():
    """reference tests/stepper_test.py:14-39."""
    from scipy.linalg import expm
    rng = np.random.default_rng(0)
    A = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    y0 = rng.normal(size=4) + 0j
    st = stepper.AdaptiveHeun(timeStep=1e-2, tol=1e-7)
    y, t = y0.copy(), 0.0
    while t < 0.3:
        y, dt = st.step(lambda y, t, intStep=0: A @ y, y, t)
        t += dt
    assert np.linalg.norm(y - expm(A * t) @ y0) < 1e-5


def test_symnet_oracle_properties():
    """SymNet restatement (jVMC/nets/sym_wrapper.py:8-66): orbit sizes of the reference's test case, invariance of
    the symmetrised amplitude under every orbit element, analytic gradients vs finite differences."""
    from oracle import symmetries as osym
    o, f = osym.orbit_1d(4, "translation", "reflection", "spinflip")
    assert o.shape == (16, 4, 4) and np.all(f == 1.0)          # 4 translations x reflection x spin flip
    o2, _ = osym.orbit_2d_square(3, "translation", "reflection", "rotation")
    assert o2.shape == (72, 9, 9)                                # |Z3 x Z3| x |D4|
    W, b = rbm.init_o1(4, 3, True, 3)
    s = np.random.default_rng(0).integers(0, 2, (9, 4))
    lp = osym.symnet_logpsi(s, W, b, o, f)
    x = 2 * s - 1
    for O in o:
        sp = ((x @ O.T) + 1) // 2
        assert np.max(np.abs(np.exp(osym.symnet_logpsi(sp, W, b, o, f) - lp) - 1)) < 1e-12
    gb, gW = osym.symnet_gradients(s, W, b, o, f)
    eps = 1e-6
    for (i, j) in [(0, 0), (1, 2), (3, 1)]:
        Wp = W.copy(); Wp[i, j] += eps
        Wm = W.copy(); Wm[i, j] -= eps
        fd = (osym.symnet_logpsi(s, Wp, b, o, f) - osym.symnet_logpsi(s, Wm, b, o, f)) / (2 * eps)
        assert np.max(np.abs(fd - gW[:, i, j])) < 1e-8
    bp = b.copy(); bp[1] += eps
    bm = b.copy(); bm[1] -= eps
    fd = (osym.symnet_logpsi(s, W, bp, o, f) - osym.symnet_logpsi(s, W, bm, o, f)) / (2 * eps)
    assert np.max(np.abs(fd - gb[:, 1])) < 1e-8
