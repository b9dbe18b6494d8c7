"""-m gpu: BASELINE.json full-size checks through size-independent properties (no oracle run at this size):
config 2 (2D TFIM 10x10, CpxRBM alpha=4, N=100, M=400, 2^16 samples) and the config-5 lattice (20x20, M=1600)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import jVMC  # noqa: E402
import jVMC.operator as op  # noqa: E402
from jVMC.stats import SampledObs, RBMGradientObs  # noqa: E402
from vmc_jax_b200 import kernels as K  # noqa: E402

import gpu_checks as G  # noqa: E402
from oracle import rbm as orbm, bfo as obfo  # noqa: E402


def tfim2d(Lx, Ly, g, parts="all"):
    H = op.BranchFreeOperator()
    for x in range(Lx):
        for y in range(Ly):
            l = x * Ly + y
            if parts in ("all", "zz"):
                H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz(x * Ly + (y + 1) % Ly))))
                H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz(((x + 1) % Lx) * Ly + y))))
            if parts in ("all", "x"):
                H.add(op.scal_opstr(g, (op.Sx(l),)))
    return H


@pytest.fixture(scope="module")
def cfg2():
    shape, M = (10, 10), 400
    N = 100
    psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=M, bias=False), seed=1)
    psi(torch.zeros((1, 1) + shape, dtype=torch.int32, device="cuda"))
    W, b = orbm.init_o1(N, M, False, 4321)
    psi.set_parameters(torch.as_tensor(orbm.flatten_params(W, b)))
    smp = jVMC.sampler.MCSampler(psi, shape, 4321, updateProposer=jVMC.sampler.propose_spin_flip, numChains=2368,
                                 sweepSteps=N, thermalizationSweeps=25, numSamples=2 ** 16)
    s, logPsi, p = smp.sample()
    return dict(psi=psi, smp=smp, s=s, logPsi=logPsi, p=p, W=W, N=N, M=M)


def test_cfg2_sampler_and_logpsi(cfg2):
    s, logPsi, p, psi = cfg2["s"], cfg2["logPsi"], cfg2["p"], cfg2["psi"]
    assert s.shape == (1, 66304, 10, 10) and cfg2["smp"].get_last_number_of_samples() == 66304
    assert int(s.min()) == 0 and int(s.max()) == 1
    assert abs(float(p.sum()) - 1.0) < 1e-12
    acc = float(cfg2["smp"].acceptance_ratio())
    assert 0.02 < acc < 0.98
    # Z2 property: without bias logpsi is invariant under a global spin flip (cosh is even)
    lp2 = psi(1 - s)
    assert float(torch.max(torch.abs(lp2.real - logPsi.real))) < 1e-10
    # spot parity with the oracle on 256 of the 66304 samples
    idx = torch.arange(0, 66304, 259, device="cuda")
    ref = orbm.cpx_rbm_logpsi(G.host(s[0, idx]).reshape(len(idx), -1), cfg2["W"])
    assert np.max(np.abs(G.host(logPsi[0, idx]) - ref) / np.abs(ref)) < 1e-10


def test_cfg2_eloc_linearity_and_spot_parity(cfg2):
    s, logPsi, psi = cfg2["s"], cfg2["logPsi"], cfg2["psi"]
    g = 3.04
    E = tfim2d(10, 10, g).get_O_loc(s, psi, logPsi)
    Ezz = tfim2d(10, 10, g, "zz").get_O_loc(s, psi, logPsi)
    Ex = tfim2d(10, 10, g, "x").get_O_loc(s, psi, logPsi)
    E2 = tfim2d(10, 10, 2 * g).get_O_loc(s, psi, logPsi)
    scale = float(E.abs().max())
    assert float((E - (Ezz + Ex)).abs().max()) < 1e-10 * scale          # additivity over operator strings
    assert float((E2 - E - Ex).abs().max()) < 1e-10 * scale             # linearity in the prefactor
    assert float(Ezz.imag.abs().max()) == 0.0                           # diagonal part is real
    idx = torch.arange(0, 66304, 1031, device="cuda")
    sub = G.host(s[0, idx]).reshape(len(idx), -1)
    tab = obfo.Tables(obfo.tfim_strings((10, 10), g, -1.0))
    ref = obfo.get_O_loc(tab, sub, lambda x: orbm.cpx_rbm_logpsi(x, cfg2["W"]))
    assert G.relerr(G.host(E[0, idx]), ref) < 1e-10


def test_cfg2_s_primes_structure(cfg2):
    """get_s_primes at full size on a 4096-sample slice (2.6 GB would be needed for all): structural invariants."""
    s = cfg2["s"][:, :4096].contiguous()
    H = tfim2d(10, 10, 3.04)
    sp, matEl = H.get_s_primes(s)
    Kmax = matEl.shape[2]
    assert Kmax in (100, 101) and sp.shape == (1, 4096 * Kmax, 10, 10)
    spv = sp.reshape(4096, Kmax, 100)
    flips = (spv != s.reshape(4096, 1, 100)).sum(-1)
    cnt = H.numNonzero.reshape(-1)
    # slot 0 = merged diagonal (no flip) when non-zero, then one single-site flip per Sx string in site order
    has_diag = cnt == 101
    assert bool(((flips[:, 0] == 0) == has_diag).all())
    assert bool((flips[has_diag][:, 1:101] == 1).all())
    assert torch.allclose(matEl.reshape(4096, Kmax)[has_diag][:, 1:101], torch.full((1,), 3.04, dtype=torch.complex128,
                                                                                   device="cuda"))
    zz = (2.0 * s.reshape(4096, 10, 10) - 1)
    ediag = -(zz * torch.roll(zz, -1, 2)).sum((1, 2)) - (zz * torch.roll(zz, -1, 1)).sum((1, 2))
    assert torch.allclose(matEl.reshape(4096, Kmax)[has_diag][:, 0].real, ediag[has_diag].to(torch.float64))


def test_cfg2_gram_matvec_identity(cfg2):
    """S at P_c = 40 000 (25.6 GB): exactly Hermitian, and S.v equals the matrix-free Khatri-Rao product
    sum_n p_n conj(dO_n) (dO_n . v) built from the moment / mat-vec kernels (validated against the oracle at small
    sizes) -- 1e-10 relative."""
    psi, s, p = cfg2["psi"], cfg2["s"], cfg2["p"]
    Gobs = RBMGradientObs(psi, s, p)
    A = Gobs.gram_A()
    assert A.shape == (40000, 40000)
    blk = A[:4000, 36000:]
    assert torch.equal(blk, A[36000:, :4000].conj().T)
    assert float(torch.diagonal(A).imag.abs().max()) == 0.0 and float(torch.diagonal(A).real.min()) > 0.0
    gen = torch.Generator(device="cuda").manual_seed(7)
    v = torch.randn(40000, dtype=torch.complex128, device="cuda", generator=gen)
    mu = Gobs.kr_mean()
    u = K.rbm_krmatvec(Gobs._s, Gobs._tau, v.reshape(mu.shape), False) - (mu.reshape(-1) * v).sum()
    w = K.rbm_moments(Gobs._s, Gobs._tau, Gobs._p * u, False, 1) - mu.conj() * (Gobs._p * u).sum()
    Av = A @ v
    assert float((Av - w.reshape(-1)).abs().max() / Av.abs().max()) < 1e-10
    # trace identity: tr S = sum_c Var(O_c)
    assert abs(float(torch.diagonal(A).real.sum() / Gobs.var()[:40000].sum()) - 1.0) < 1e-10
    del A


def test_cfg5_lattice_sampler_eloc_minsr():
    """20x20 TFIM, CpxRBM alpha=4 (N=400, M=1600, P_c=640 000): sampler + fused E_loc + Khatri-Rao MinSR kernel
    on 4096 samples; T checked through T.x = Obar (Obar^dagger x) computed matrix-free."""
    shape, N, M = (20, 20), 400, 1600
    psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=M, bias=False), seed=1)
    psi(torch.zeros((1, 1) + shape, dtype=torch.int32, device="cuda"))
    W, b = orbm.init_o1(N, M, False, 4321)
    psi.set_parameters(torch.as_tensor(orbm.flatten_params(W, b)))
    smp = jVMC.sampler.MCSampler(psi, shape, 11, updateProposer=jVMC.sampler.propose_spin_flip, numChains=1024,
                                 sweepSteps=N, thermalizationSweeps=5, numSamples=4096)
    s, logPsi, p = smp.sample()
    assert s.shape == (1, 4096, 20, 20)
    H = tfim2d(20, 20, 3.04)
    E = H.get_O_loc(s, psi, logPsi)
    idx = torch.arange(0, 4096, 257, device="cuda")
    tab = obfo.Tables(obfo.tfim_strings((20, 20), 3.04, -1.0))
    ref = obfo.get_O_loc(tab, G.host(s[0, idx]).reshape(len(idx), -1), lambda x: orbm.cpx_rbm_logpsi(x, W))
    assert G.relerr(G.host(E[0, idx]), ref) < 1e-10
    Gobs = RBMGradientObs(psi, s, p)
    T = Gobs.tangent_kernel()
    assert T.shape == (4096, 4096) and torch.equal(T, T.conj().T)
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(4096, dtype=torch.complex128, device="cuda", generator=gen)
    # Obar^dagger x (Khatri-Rao order), then Obar (.)  ; factor 2 = doubled holomorphic layout
    mu = Gobs.kr_mean()
    wgt = torch.sqrt(Gobs._p) * x
    y = K.rbm_moments(Gobs._s, Gobs._tau, wgt, False, 1) - mu.conj() * wgt.sum()       # conj(Obar)^T x
    Tx = 2.0 * torch.sqrt(Gobs._p) * (K.rbm_krmatvec(Gobs._s, Gobs._tau, y, False) - (mu * y).sum())
    ref = T @ x
    assert float((Tx - ref).abs().max() / ref.abs().max()) < 1e-10
    upd = jVMC.util.MinSR(smp, pinvTol=1e-8).solve(SampledObs(E, p), Gobs, holomorphic=True)
    assert upd.shape == (2 * 2 * N * M // 2 * 1,) or upd.numel() == 2 * N * M
    assert bool(torch.isfinite(upd.real).all())


def test_cfg4_cnn_12x12_parity_on_sampled_subset():
    """BASELINE configs[3] lattice: 2D Heisenberg J1 12x12 (Marshall-rotated), real CNN F=(3,3) channels=(6,4), exchange
    proposer.  Sampled configurations stay in the zero-magnetisation sector; log psi, the per-sample gradients (central
    finite differences of the oracle, the reference's own method for real nets, tests/vqs_test.py:133-164) and E_loc
    (the reference's s' -> psi(s') algorithm) are compared with oracle/cnn.py on a subset of the samples."""
    from oracle import cnn as ocnn
    L = 12
    H = op.BranchFreeOperator(ElocBatchSize=512)
    strings = []
    for x in range(L):
        for y in range(L):
            i = x * L + y
            for j in (x * L + (y + 1) % L, ((x + 1) % L) * L + y):
                H.add(op.scal_opstr(-0.25, (op.Sx(i), op.Sx(j))))
                H.add(op.scal_opstr(-0.25, (op.Sy(i), op.Sy(j))))
                H.add(op.scal_opstr(0.25, (op.Sz(i), op.Sz(j))))
                strings += [(-0.25, [obfo.Sx(i), obfo.Sx(j)]), (-0.25, [obfo.Sy(i), obfo.Sy(j)]),
                            (0.25, [obfo.Sz(i), obfo.Sz(j)])]
    kw = dict(F=(3, 3), channels=(6, 4), strides=(1, 1), bias=True, firstLayerBias=False)
    psi = jVMC.vqs.NQS(jVMC.nets.CNN(**kw), seed=7)
    neel = np.indices((L, L)).sum(0) % 2
    smp = jVMC.sampler.MCSampler(psi, (L, L), 99, updateProposer=jVMC.sampler.propose_spin_flip_zeroMag, numChains=296,
                                 numSamples=592, thermalizationSweeps=4, sweepSteps=L * L, initState=neel)
    s, logPsi, p = smp.sample()
    assert s.shape == (1, 592, L, L)
    assert bool((s.reshape(592, -1).sum(1) == L * L // 2).all())          # exchange moves conserve the magnetisation
    assert 0.05 < float(smp.acceptance_ratio()) <= 1.0
    theta = G.host(psi.get_parameters())
    idx = np.arange(0, 592, 37)
    sn = G.host(s[0])[idx]
    okw = dict(F=(3, 3), channels=(6, 4), strides=(1, 1), actFun=("elu",), bias=True, firstLayerBias=False)
    ref = ocnn.cnn_logpsi(sn, theta, **okw)
    got = G.host(logPsi[0])[idx]
    assert np.max(np.abs(got.real - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-10 and np.all(got.imag == 0)
    g = G.host(psi.gradients(s[:, torch.as_tensor(idx, device="cuda")]))[0]
    eps = 1e-6
    for k in range(0, theta.size, 13):
        tp, tm = theta.copy(), theta.copy()
        tp[k] += eps
        tm[k] -= eps
        fd = (ocnn.cnn_logpsi(sn, tp, **okw) - ocnn.cnn_logpsi(sn, tm, **okw)) / (2 * eps)
        assert np.max(np.abs(g[:, k].real - fd)) < 1e-7 * max(1.0, np.max(np.abs(fd)))
    E = H.get_O_loc(s, psi, logPsi)
    tab = obfo.Tables(strings)
    refE = obfo.get_O_loc(tab, sn.reshape(len(idx), -1),
                          lambda x: ocnn.cnn_logpsi(x.reshape((-1, L, L)), theta, **okw).astype(np.complex128))
    assert G.relerr(G.host(E[0])[idx], refE) < 1e-10


def test_cfg3_chain_L40_bias_eloc_F_parity():
    """BASELINE configs[2] shape: 1D TFIM L=40, CpxRBM alpha=2 with bias (P = 6560), 2^16 samples: E_loc on a subset vs the
    reference's algorithm, F = <conj(dO) dE> vs the dense oracle on a subset re-weighted consistently, and the
    matrix-free identity S.v for the full-sample Gram matrix."""
    L, M = 40, 80
    psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=M, bias=True), seed=1)
    psi(torch.zeros((1, 1, L), dtype=torch.int32, device="cuda"))
    W, b = orbm.init_o1(L, M, True, 4321)
    psi.set_parameters(torch.as_tensor(orbm.flatten_params(0.3 * W, 0.3 * b)))
    W, b = 0.3 * W, 0.3 * b
    H = op.BranchFreeOperator()
    for l in range(L):
        H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz((l + 1) % L))))
        H.add(op.scal_opstr(-0.3, (op.Sx(l),)))
    smp = jVMC.sampler.MCSampler(psi, (L,), 4321, updateProposer=jVMC.sampler.propose_spin_flip, numChains=2368,
                                 sweepSteps=L, thermalizationSweeps=25, numSamples=2 ** 16)
    s, logPsi, p = smp.sample()
    B = s.shape[1]
    assert B == 66304
    E = H.get_O_loc(s, psi, logPsi)
    idx = np.arange(0, B, 259)
    tab = obfo.Tables(obfo.tfim_strings((L,), -0.3, -1.0))
    sn = G.host(s[0])[idx]
    refE = obfo.get_O_loc(tab, sn, lambda x: orbm.cpx_rbm_logpsi(x, W, b))
    assert G.relerr(G.host(E[0])[idx], refE) < 1e-10
    assert G.relerr(G.host(logPsi[0])[idx].real, orbm.cpx_rbm_logpsi(sn, W, b).real) < 1e-10
    Gobs = RBMGradientObs(psi, s, p)
    Eobs = SampledObs(E, p)
    A = Gobs.gram_A()
    assert torch.equal(A, A.conj().T)
    # S.v matrix-free: A v = sum_n p_n conj(O_n - mu) ((O_n - mu).v) in Khatri-Rao order
    gen = torch.Generator(device="cuda").manual_seed(5)
    v = torch.randn(A.shape[0], dtype=torch.complex128, device="cuda", generator=gen)
    mu = Gobs.kr_mean()
    t = K.rbm_krmatvec(Gobs._s, Gobs._tau, v.reshape(mu.shape), True) - (mu.reshape(-1) * v).sum()
    Av = K.rbm_moments(Gobs._s, Gobs._tau, Gobs._p * t, True, 1).reshape(-1) - mu.conj().reshape(-1) * (Gobs._p * t).sum()
    ref = A @ v
    assert float((Av - ref).abs().max() / ref.abs().max()) < 1e-10
    # F on the subset, weights renormalised: same estimator through the oracle's dense route
    sub = torch.as_tensor(idx, device="cuda")
    ps = p[:, sub] / p[:, sub].sum()
    Fk = RBMGradientObs(psi, s[:, sub], ps).covar(SampledObs(E[:, sub], ps)).reshape(-1)
    from oracle import stats as ostats
    pn = G.host(ps)[0]
    oF = ostats.SampledObs(orbm.gradients_holomorphic(sn, W, b), pn).covar(ostats.SampledObs(G.host(E[0])[idx], pn)).ravel()
    assert G.relerr(G.host(Fk), oF) < 1e-10


def test_tdvp_solve_parity_6x6_alpha4():
    """TDVP solve at 6x6, alpha = 4 (P_c = 5184, P = 10 368) on 4144 sampled configurations against the oracle's
    reference-layout solve (reference jVMC/util/tdvp.py:153-213; NumPy eigh of the 10 368 x 10 368 matrix, ~30 s):
    spectrum, F, Var E, the eigenspace-summed SNR ingredients, and -- with the SNR weighting switched off on both sides,
    where the update is unique -- residual, cutoff and update."""
    from oracle import solve as osolve, stats as ostats
    L, alpha = 6, 4
    N, M = L * L, alpha * L * L
    psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=M, bias=False), seed=1)
    psi(torch.zeros((1, 1, L, L), dtype=torch.int32, device="cuda"))
    W, b = orbm.init_o1(N, M, False, 4321)
    psi.set_parameters(torch.as_tensor(orbm.flatten_params(W, b)))
    H = tfim2d(L, L, 3.04)
    smp = jVMC.sampler.MCSampler(psi, (L, L), 4321, updateProposer=jVMC.sampler.propose_spin_flip, numChains=296,
                                 sweepSteps=N, thermalizationSweeps=10, numSamples=4096)
    s, logPsi, p = smp.sample()
    B = s.shape[1]
    Eloc = H.get_O_loc(s, psi, logPsi)
    import jVMC.mpi_wrapper as mpi
    res = {}
    for snrTol in (0.0, 2.0):
        td = jVMC.util.TDVP(smp, snrTol=snrTol, pinvTol=1e-8, pinvCutoff=1e-8, rhsPrefactor=1., diagonalShift=1e-3,
                            makeReal='real')
        upd, r, c = td.solve(SampledObs(Eloc, p), RBMGradientObs(psi, s, p))
        res[snrTol] = (td, upd, float(r), float(c))
    sn, pn, En = G.host(s[0]).reshape(B, -1), G.host(p[0]), G.host(Eloc[0])
    oE, oG = ostats.SampledObs(En, pn), ostats.SampledObs(orbm.gradients_holomorphic(sn, W, b), pn)
    otd = osolve.TDVP(snrTol=2, pinvTol=1e-8, pinvCutoff=1e-8, rhsPrefactor=1., diagonalShift=1e-3, makeReal='real',
                      exact_sampler=True)                      # exact_sampler: SNR computed but not applied (tdvp.py:203)
    upd_ref, res_ref, cut_ref = otd.solve(oE, oG, mpi.globNumSamples)
    td0, upd0, r0, c0 = res[0.0]
    scale = np.abs(otd.ev).max()
    assert np.allclose(G.host(td0.ev), otd.ev, atol=1e-9 * scale)
    assert np.allclose(G.host(td0.F0), otd.F0, rtol=1e-10, atol=1e-13)
    assert np.isclose(float(td0.ElocVar), otd.ElocVar, rtol=1e-10)
    assert c0 == cut_ref and np.isclose(r0, res_ref, rtol=1e-5, atol=1e-9)
    assert np.allclose(G.host(upd0), upd_ref, rtol=1e-5, atol=1e-6 * np.abs(upd_ref).max())
    # SNR ingredients summed over each doubly degenerate eigenspace (basis independent)
    td2 = res[2.0][0]
    vtf2 = (np.abs(otd.VtF) ** 2).reshape(-1, 2).sum(1)
    rv = otd.rhoVar.reshape(-1, 2).sum(1)
    ok = otd.ev.reshape(-1, 2)[:, 0] > 1e-9 * scale
    snr_pair = np.sqrt(np.abs(mpi.globNumSamples * vtf2 / rv))
    assert np.allclose(G.host(td2.snr).reshape(-1, 2)[:, 0][ok], snr_pair[ok], rtol=1e-4)
    assert np.isfinite(res[2.0][2]) and bool(torch.isfinite(res[2.0][1]).all())
