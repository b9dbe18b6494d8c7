"""Kernel-level parity checks: CUDA path (through the C ABI) vs the CPU oracle on identical seeded inputs.
Used by tests/test_gpu_kernels.py (-m gpu) and by __graft_entry__.smoke()."""
import numpy as np
import torch

from oracle import rbm as orbm, bfo as obfo, stats as ostats, solve as osolve, sampling as osamp, symmetries as osym, cnn as ocnn
from vmc_jax_b200 import kernels as K

DEV = "cuda:0"
RTOL = 1e-10   # north_star: logpsi, E_loc, F and S within 1e-10 relative (fp64)


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def host(t):
    return t.detach().cpu().numpy()


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    den = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / den) if a.size else 0.0


def rand_configs(B, N, seed):
    return np.random.default_rng(seed).integers(0, 2, (B, N)).astype(np.int32)


# ---------------------------------------------------------------- logpsi / tau / tables
def check_logpsi(N=20, M=40, B=333, bias=True, seed=1, real=False):
    W, b = orbm.init_o1(N, M, bias, seed, real=real)
    s = rand_configs(B, N, seed + 1)
    Wc = np.asarray(W, np.complex128)
    bc = None if b is None else np.asarray(b, np.complex128)
    lp, tau = K.rbm_logpsi(dev(s), dev(Wc), None if bc is None else dev(bc))
    ref = orbm.real_rbm_logpsi(s, W, b) if real else orbm.cpx_rbm_logpsi(s, Wc, bc)
    e1 = np.max(np.abs(host(lp) - ref) / np.maximum(np.abs(ref), 1e-3))
    e2 = relerr(host(tau), orbm.tau(s, Wc, bc))
    assert e1 < RTOL and e2 < RTOL, (e1, e2)
    return e1, e2


def check_tables(N=12, M=24, bias=True, seed=3):
    W, b = orbm.init_o1(N, M, bias, seed)
    t = host(K.rbm_tables(dev(W), None if b is None else dev(b)))
    T = t[:N * M].reshape(N, M)
    lc = t[N * M:N * M + N]
    tb2 = t[N * M + N:N * M + N + M]
    lcb = t[-1]
    assert relerr(T, np.tanh(2 * W)) < RTOL
    assert relerr(np.exp(lc), np.prod(np.cosh(2 * W), axis=1)) < RTOL
    if bias:
        assert relerr(tb2, np.tanh(2 * b)) < RTOL
        assert relerr(np.exp(lcb), np.prod(np.cosh(2 * b))) < RTOL
    else:
        assert np.all(tb2 == 0) and lcb == 0


# ---------------------------------------------------------------- operators
def op_tables_to_device(tab):
    isDiag = np.zeros(tab.numOps, np.uint8)
    isDiag[tab.diag] = 1
    return K.OpTables(tab.idx, tab.map, tab.matEls, tab.fermi, isDiag, DEV)


def operator_zoo(N):
    """Operators covering the reference's test cases (tests/operator_test.py) and BASELINE configs."""
    zoo = {}
    zoo["tfim1d"] = obfo.tfim_strings((N,), -0.7, -1.0)
    zoo["tfim_field"] = [x for l in range(N) for x in
                         [(-1., [obfo.Sz(l), obfo.Sz((l + 1) % N)]), (-0.7, [obfo.Sx(l)]), (0.1, [obfo.Sz(l)])]]
    zoo["sx_sysz"] = [x for i in range(N) for x in
                      [(2., [obfo.Sx(i)]), (2., [obfo.Sy(i), obfo.Sz((i + 1) % N)])]]
    zoo["splus"] = [(2., [obfo.Sp(i)]) for i in range(3)]
    zoo["heis"] = [x for l in range(N) for x in
                   [(1., [obfo.Sx(l), obfo.Sx((l + 1) % N)]), (1., [obfo.Sy(l), obfo.Sy((l + 1) % N)]),
                    (1., [obfo.Sz(l), obfo.Sz((l + 1) % N)])]]
    zoo["single_diag"] = [(0.5, [obfo.Sz(0)]), (1.5, [obfo.Sx(1)])]
    zoo["same_site"] = [(1.0, [obfo.Sz(2), obfo.Sx(2)]), (0.25, [obfo.Sx(1), obfo.Sx(1)]), (1.0, [obfo.Sm(0)])]
    return zoo


def hubbard_strings(L=4):
    t, mu, V, fl = -1.0, -2.0, 4.0, 2
    n = fl * L
    up, do = 0, n - 1
    S = []
    for i in range(L):
        S.append((V, [obfo.number(up + i), obfo.number(do - i)]))
        S.append((mu, [obfo.number(up + i)]))
        S.append((mu, [obfo.number(do - i)]))
        if i == L - 1:
            continue
        S.append((t, [obfo.creation(up + i + 1), obfo.annihilation(up + i)]))
        S.append((t, [obfo.creation(up + i), obfo.annihilation(up + i + 1)]))
        S.append((t, [obfo.creation(do - i - 1), obfo.annihilation(do - i)]))
        S.append((t, [obfo.creation(do - i), obfo.annihilation(do - i - 1)]))
    return S


def check_s_primes(strings, N, B=97, seed=5, args=()):
    """Bit-exact s' / matEl / counts against the oracle (north_star)."""
    tab = obfo.Tables(strings)
    s = rand_configs(B, N, seed)
    sp_ref, m_ref, cnt_ref = obfo.get_s_primes(tab, s, *args)
    dt = op_tables_to_device(tab)
    sp, m, cnt = K.bfo_s_primes(dev(s), dt, dev(tab.eval_prefactors(*args)))
    assert np.array_equal(host(cnt), cnt_ref)
    assert host(sp).shape == sp_ref.shape and np.array_equal(host(sp), sp_ref)
    assert np.array_equal(host(m), m_ref), np.max(np.abs(host(m) - m_ref))


def check_eloc(strings, N=10, M=20, B=203, bias=False, seed=7, args=()):
    """Fused RBM E_loc and the generic (s' -> logpsi -> reduce) path vs the reference's algorithm."""
    tab = obfo.Tables(strings)
    W, b = orbm.init_o1(N, M, bias, seed)
    s = rand_configs(B, N, seed + 1)
    f = lambda x: orbm.cpx_rbm_logpsi(x, W, b)
    ref = obfo.get_O_loc(tab, s, f, *args)
    dW, db, ds = dev(W), (None if b is None else dev(b)), dev(s)
    lp, tau = K.rbm_logpsi(ds, dW, db)
    dt = op_tables_to_device(tab)
    pref = dev(tab.eval_prefactors(*args))
    # generic path
    sp, m, cnt = K.bfo_s_primes(ds, dt, pref)
    lpsp, _ = K.rbm_logpsi(sp, dW, db, want_tau=False)
    e_gen = host(K.oloc_reduce(m, lp, lpsp))
    # fused path
    tables = K.rbm_tables(dW, db)
    e_fused, err = K.rbm_eloc(ds, tau, tables, dt, pref)
    assert int(err.item()) == 0
    r1, r2 = relerr(e_gen, ref), relerr(host(e_fused), ref)
    assert r1 < RTOL and r2 < RTOL, (r1, r2)
    return r1, r2


# ---------------------------------------------------------------- gradients / moments / Gram
def kr_gradients(s, W, b):
    """Oracle gradients in Khatri-Rao complex order c = r*M + j (bias pseudo-site first)."""
    G = orbm.gradients_holomorphic(s, W, b)
    N, M = W.shape
    if b is None:
        return G[:, :N * M]
    return np.concatenate([G[:, :M], G[:, 2 * M:2 * M + N * M]], axis=1)


def check_grad(N=6, M=5, B=17, bias=True, seed=9):
    W, b = orbm.init_o1(N, M, bias, seed)
    s = rand_configs(B, N, seed + 1)
    ds = dev(s)
    _, tau = K.rbm_logpsi(ds, dev(W), None if b is None else dev(b))
    g = host(K.rbm_grad(ds, tau, bias, 0))
    assert relerr(g, orbm.gradients_holomorphic(s, W, b)) < RTOL
    Wr, br = orbm.init_o1(N, M, bias, seed, real=True)
    _, taur = K.rbm_logpsi(ds, dev(Wr.astype(np.complex128)), None if br is None else dev(br.astype(np.complex128)))
    gr = host(K.rbm_grad(ds, taur, bias, 1))
    assert relerr(gr, orbm.gradients_real(s, Wr, br)) < RTOL


def check_moments_gram(N=7, M=24, B=211, bias=True, seed=11, uniform=True, tile=0, backend="dmma"):
    """mu, F-kernel and the centred Gram A = <conj(O) O>_c vs the oracle's SampledObs.covar."""
    W, b = orbm.init_o1(N, M, bias, seed)
    s = rand_configs(B, N, seed + 1)
    rng = np.random.default_rng(seed + 2)
    p = np.ones(B) / B if uniform else rng.uniform(0.5, 1.5, B)
    p = p / p.sum()
    E = rng.normal(size=B) + 1j * rng.normal(size=B)
    O = kr_gradients(s, W, b)
    obsO, obsE = ostats.SampledObs(O, p), ostats.SampledObs(E, p)
    A_ref = obsO.covar()
    F_ref = obsO.covar(obsE).ravel()
    ds = dev(s)
    _, tau = K.rbm_logpsi(ds, dev(W), None if b is None else dev(b))
    dp = dev(p)
    mu = K.rbm_moments(ds, tau, dp.to(torch.complex128), bias, 0)
    assert relerr(host(mu).ravel(), obsO.mean()) < RTOL
    Ebar = np.sum(p * E)
    Fk = K.rbm_moments(ds, tau, dev(p * (E - Ebar)), bias, 1)
    assert relerr(host(Fk).ravel(), F_ref) < RTOL
    sigT = K.pack_sigma(ds, bias)
    gram = (lambda Y, sg, m, al, ka: K.rbm_gram_S_i8(Y, sg, m, al, ka)) if backend == "i8" else \
        (lambda Y, sg, m, al, ka: K.rbm_gram_S(Y, sg, m, al, ka, tile=tile))
    if uniform:
        A = gram(tau, sigT, mu, p[0], 1.0)
    else:
        A = gram(tau * torch.sqrt(dp)[:, None], sigT, mu, 1.0, 1.0)
    A = host(A)
    # scale: the uncentred second moment (A is a difference of two such terms)
    scale = max(np.max(np.abs(A_ref)), np.max(np.abs(obsO.mean())) ** 2)
    r = float(np.max(np.abs(A - A_ref)) / scale)
    assert r < RTOL, r
    assert np.array_equal(A, A.conj().T)   # exactly Hermitian by construction
    return r


def check_gram_heavy_tail(N=5, M=40, B=4096, seed=31, tail=3000.0):
    """tau columns whose maximum is far above their typical size (tanh(theta) next to a pole), and a uniformly tiny one.
    The check is entry-wise, |dA_jl| <= 1e-10 sqrt(A_jj A_ll): stricter than the max|A| normalisation of
    check_moments_gram exactly where it matters (small entries of A may not hide behind a large one).  The int8 digits
    resolve 2^-41 of the column MAXIMUM, so entries BETWEEN two heavy-tailed columns lose accuracy in this metric
    (error ~ 0.3 * 4 * 2^-40 r_z r_z' / sqrt(B), r = max / rms); the automatic choice must send the few outlier samples of
    such a matrix through the fp64 DMMA kernel and the bulk (and a benign matrix as a whole) through the int8 kernel, both
    within the bar; the error model behind the choice is checked too."""
    rng = np.random.default_rng(seed)
    Y = 0.5 * (rng.standard_normal((B, M)) + 1j * rng.standard_normal((B, M)))
    Yb = Y.copy()
    Y[rng.integers(0, B, 3), 7] *= tail                       # three outliers in column 7 (Re and Im)
    Y[rng.integers(0, B), 21] = tail * 0.3                    # one in column 21
    Y[:, 30] *= 1e-6                                          # and a uniformly tiny column
    Yb[:, 30] *= 1e-6
    s = rand_configs(B, N, seed + 1)
    sig = 2.0 * s - 1.0
    ds = dev(s)
    sigT = K.pack_sigma(ds, False)
    al = 1.0 / B
    res = {}
    for tag, Yc in (("heavy", Y), ("benign", Yb)):
        O = (sig[:, :, None] * Yc[:, None, :]).reshape(B, N * M)   # Khatri-Rao order
        A_ref = al * (O.conj().T @ O)
        nat = np.sqrt(np.outer(np.real(np.diag(A_ref)), np.real(np.diag(A_ref))))
        dY = dev(Yc)
        mu0 = torch.zeros((N, M), dtype=torch.complex128, device=dY.device)
        Zr = np.stack([Yc.real, Yc.imag], axis=2).reshape(B, 2 * M)
        r_ref = np.max(np.abs(Zr), axis=0) / np.sqrt(np.mean(Zr ** 2, axis=0))
        assert np.allclose(host(K.i8_tail_ratios(dY)), r_ref, rtol=1e-10)
        err = {}
        for name, fn in (("auto", lambda *a_: K.rbm_gram_S_auto(*a_, s=ds, hasBias=False)), ("i8", K.rbm_gram_S_i8),
                         ("dmma", K.rbm_gram_S)):
            A = host(fn(dY, sigT, mu0, al, 0.0))
            err[name] = float(np.max(np.abs(A - A_ref) / nat))
            assert np.array_equal(A, A.conj().T)
            if name == "auto":
                chosen, nout = K.LAST_GRAM["backend"], K.LAST_GRAM["outlier_rows"]
                r1, r2 = K.LAST_GRAM["tail_ratios"]
        assert chosen == ("i8+dmma" if tag == "heavy" else "i8"), (tag, K.LAST_GRAM)
        assert (1 <= nout <= 4) if tag == "heavy" else nout == 0
        assert err["auto"] < RTOL and err["dmma"] < RTOL, (tag, err)
        pred = K.i8_predicted_error(r1, r2, B)
        assert err["i8"] < 10.0 * pred, (tag, err, pred)       # the model bounds the measured int8 error
        res[tag] = (err, pred)
    assert res["benign"][0]["i8"] < RTOL
    return res


def check_eigh_large(n=700, cplx=True, kind="gram", seed=3):
    """kernels.eigh_large (hetrd -> Cuppen tearing -> secular equation -> unmtr; the path for n > 32768) against the
    direct cuSOLVER decomposition at a size where both run: spectrum, residual, orthonormality, on spectra that exercise
    the deflation (clusters, numerically rank-deficient Gram matrices, decoupled halves)."""
    rng = np.random.default_rng(seed)
    if kind == "gram":                     # S-like: Gram matrix of fewer samples than parameters, decaying column scales
        X = rng.standard_normal((n // 2, n)) + (1j * rng.standard_normal((n // 2, n)) if cplx else 0.0)
        X = X * np.logspace(0, -7, n)[None, :]
        A = X.conj().T @ X / (n // 2)
    elif kind == "clustered":
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0.0))
        ev = np.concatenate([np.ones(n // 3), 1.0 + 1e-9 * rng.standard_normal(n // 3), np.linspace(2, 3, n - 2 * (n // 3))])
        A = (Q * ev[None, :]) @ Q.conj().T
    elif kind == "blockdiag":               # beta = 0 at the tearing point
        A = np.zeros((n, n), complex if cplx else float)
        for lo, hi in ((0, n // 2), (n // 2, n)):
            Bk = rng.standard_normal((hi - lo, hi - lo)) + (1j * rng.standard_normal((hi - lo, hi - lo)) if cplx else 0.0)
            A[lo:hi, lo:hi] = Bk + Bk.conj().T
    else:
        Bk = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0.0)
        A = Bk + Bk.conj().T
    A = 0.5 * (A + A.conj().T)
    dA = dev(A, torch.complex128 if cplx else torch.float64)
    St = dA.T.clone(memory_format=torch.contiguous_format)
    w_ref, Vt_ref, _ = K.eigh_inplace(St.clone())
    w, Vt = K.eigh_large(St.clone())
    scale = float(w_ref.abs().max())
    assert float((w - w_ref).abs().max()) < 1e-12 * scale * n ** 0.5, float((w - w_ref).abs().max()) / scale
    V = Vt.T
    resid = float((dA @ V - V * w[None, :].to(V.dtype)).abs().max()) / scale
    orth = float((V.conj().T @ V - torch.eye(n, dtype=V.dtype, device=V.device)).abs().max())
    assert resid < 1e-11 and orth < 1e-11, (resid, orth)
    return resid, orth


def check_gram_T(N=7, M=24, B=211, bias=True, seed=21, uniform=False):
    """Khatri-Rao tangent kernel T = 2 Obar Obar^dagger and O.x mat-vec vs the dense oracle (stats.py:332-336)."""
    W, b = orbm.init_o1(N, M, bias, seed)
    s = rand_configs(B, N, seed + 1)
    rng = np.random.default_rng(seed + 2)
    p = np.ones(B) / B if uniform else rng.uniform(0.5, 1.5, B)
    p = p / p.sum()
    G = ostats.SampledObs(orbm.gradients_holomorphic(s, W, b), p)
    T_ref = G.tangent_kernel()
    ds, dp = dev(s), dev(p)
    _, tau = K.rbm_logpsi(ds, dev(W), None if b is None else dev(b))
    mu = K.rbm_moments(ds, tau, dp.to(torch.complex128), bias, 0)
    T = host(K.rbm_gram_T(ds, tau, dp, mu, bias, 2.0))
    # scale: the uncentred kernel (T is a difference of such terms; exactly 0 for a single sample)
    scale = 2.0 * np.max(p) * np.max(np.sum(np.abs(kr_gradients(s, W, b)) ** 2, axis=1))
    r = float(np.max(np.abs(T - T_ref)) / scale)
    assert r < RTOL, r
    assert np.array_equal(T, T.conj().T)
    x = rng.normal(size=mu.numel()) + 1j * rng.normal(size=mu.numel())
    u = host(K.rbm_krmatvec(ds, tau, dev(x).reshape(mu.shape), bias))
    assert relerr(u, kr_gradients(s, W, b) @ x) < RTOL
    return r


def check_minsr_solve(N=4, M=8, bias=False, seed=23):
    """MinSR update from the factorised path vs oracle minsr_solve (minsr.py:53-80) on the exact basis."""
    from vmc_jax_b200.util.minsr import pinv_hermitian
    W, b = orbm.init_o1(N, M, bias, seed)
    basis = osamp.basis_states(N)
    lp = orbm.cpx_rbm_logpsi(basis, W, b)
    p, _ = osamp.exact_probabilities(lp)
    ham = obfo.Tables(obfo.tfim_strings((N,), -0.9, -1.0))
    E = obfo.get_O_loc(ham, basis, lambda x: orbm.cpx_rbm_logpsi(x, W, b), logPsiS=lp)
    oE, oG = ostats.SampledObs(E, p), ostats.SampledObs(orbm.gradients_holomorphic(basis, W, b), p)
    upd_ref = osolve.minsr_solve(oE, oG, True, pinvTol=1e-6)
    ds, dp = dev(basis), dev(p)
    _, tau = K.rbm_logpsi(ds, dev(W), None if b is None else dev(b))
    mu = K.rbm_moments(ds, tau, dp.to(torch.complex128), bias, 0)
    T = K.rbm_gram_T(ds, tau, dp, mu, bias, 2.0)
    Tinv = pinv_hermitian(T, 1e-6)
    assert relerr(host(Tinv), osolve.pinv_hermitian(oG.tangent_kernel(), 1e-6)) < 1e-7
    e = dev(oE._data.reshape(-1))
    x = Tinv @ e
    wgt = torch.sqrt(dp) * x
    Fk = (K.rbm_moments(ds, tau, wgt, bias, 1) - mu.conj() * wgt.sum()).reshape(-1)
    Mb = M if bias else 0
    parts = ([Fk[:Mb], -1j * Fk[:Mb]] if bias else []) + [Fk[Mb:], -1j * Fk[Mb:]]
    upd = -torch.cat(parts)
    r = relerr(host(upd), upd_ref)
    assert r < 1e-7, r
    return r


# ---------------------------------------------------------------- solve
def check_tdvp_solve(N=4, M=3, bias=True, makeReal='real', rhsPrefactor=1.0, shift=2.0, seed=13):
    """S expansion + eigh + regulariser vs oracle TDVP.solve on the exact basis (no SNR cut, unique update)."""
    W, b = orbm.init_o1(N, M, bias, seed)
    basis = osamp.basis_states(N)
    lp = orbm.cpx_rbm_logpsi(basis, W, b)
    p, _ = osamp.exact_probabilities(lp)
    ham = obfo.Tables(obfo.tfim_strings((N,), -0.9, -1.0))
    E = obfo.get_O_loc(ham, basis, lambda x: orbm.cpx_rbm_logpsi(x, W, b), logPsiS=lp)
    td = osolve.TDVP(snrTol=1, pinvTol=0.0, pinvCutoff=1e-8, rhsPrefactor=rhsPrefactor, diagonalShift=shift,
                     makeReal=makeReal, exact_sampler=True)
    upd_ref, res_ref, cut_ref = osolve.tdvp_rhs(W, b, basis, p, E, td, 2 ** N, orbm.gradients_holomorphic)
    # device
    ds, dp = dev(basis), dev(p)
    _, tau = K.rbm_logpsi(ds, dev(W), None if b is None else dev(b))
    mu = K.rbm_moments(ds, tau, dp.to(torch.complex128), bias, 0)
    Ebar = np.sum(p * E)
    Fc = K.rbm_moments(ds, tau, dev(p * (E - Ebar)), bias, 1).reshape(-1)
    sigT = K.pack_sigma(ds, bias)
    A = K.rbm_gram_S(tau * torch.sqrt(dp)[:, None], sigT, mu, 1.0, 1.0)
    mode = 0 if makeReal == 'real' else 1
    St = K.expand_S(A, M, N, bias, mode, shift)
    S_ref = td.S
    assert relerr(host(St).T, S_ref) < RTOL
    # F0 = -x [Fc, -i Fc] per leaf (flat layout), F = q(F0)
    Mb = M if bias else 0
    parts = []
    for lo, hi in ([(0, Mb)] if bias else []) + [(Mb, Mb + N * M)]:
        parts += [Fc[lo:hi], -1j * Fc[lo:hi]]
    F0 = (-rhsPrefactor) * torch.cat(parts)
    assert relerr(host(F0), td.F0) < RTOL
    F = F0.real.to(torch.complex128) if mode == 0 else (1j * F0.imag)
    ev, Vt, info = K.eigh_inplace(St)
    assert int(info.item()) == 0
    assert relerr(host(ev), td.ev) < 1e-9
    VtF = torch.mv(Vt.conj().to(torch.complex128), F)
    pinvEv, scal = K.tdvp_regularize(ev, VtF, None, F, 0.0, 1e-8, 1.0)
    upd = torch.mv(Vt.to(torch.complex128).T, pinvEv * VtF).real
    r = relerr(host(upd), upd_ref)
    assert r < 1e-6, r
    assert abs(scal[1].item() - cut_ref) < 1e-15
    assert abs(scal[0].item() - res_ref) < 1e-8 * max(1.0, res_ref)
    return r


# ---------------------------------------------------------------- sampler
def chi2_pvalue(counts, pex):
    from scipy.stats import chi2
    n = counts.sum()
    exp = pex * n
    keep = exp >= 5
    c = np.concatenate([counts[keep], [counts[~keep].sum()]]) if (~keep).any() else counts[keep]
    e = np.concatenate([exp[keep], [exp[~keep].sum()]]) if (~keep).any() else exp[keep]
    ok = e > 0
    stat = np.sum((c[ok] - e[ok]) ** 2 / e[ok])
    return float(chi2.sf(stat, ok.sum() - 1)), stat


def run_sampler(W, b, C, numSamples, proposer="spin_flip", mu=2.0, seed=4321, K_sweep=None, therm=20, init=None,
                refreshEvery=1):
    N, M = W.shape
    K_sweep = N if K_sweep is None else K_sweep
    dW, db = dev(W), (None if b is None else dev(b))
    tables = K.rbm_tables(dW, db)
    states = torch.zeros((C, N), dtype=torch.int32, device=DEV)
    if init is not None:
        states[:] = dev(np.asarray(init, np.int32))
    counters = torch.zeros(2, dtype=torch.int64, device=DEV)
    spc = (numSamples + C - 1) // C
    cfg = K.rbm_mcmc(states, dW, db, tables, seed, 0, 0, proposer, mu, K_sweep, therm * K_sweep, spc, counters,
                     refreshEvery)
    torch.cuda.synchronize()
    return host(cfg), host(counters), host(states)


def check_sampler_chi2(N=4, M=2, weights=None, proposer="spin_flip", mu=2.0, C=592, numSamples=1_000_000,
                       bias=False, seed=4321, sector=False, refreshEvery=1, amp=None):
    """Sampled distribution vs ExactSampler probabilities, chi-squared p-value > 1e-3 (north_star).
    amp: amplitude of the uniform weight distribution (default 1/sqrt(N))."""
    if weights is None:
        W, b = orbm.init_o1(N, M, bias, 17)
        if amp is not None:
            W = W * (amp * np.sqrt(N))
            b = None if b is None else b * (amp * np.sqrt(N))
    else:
        W, b = orbm.unflatten_params(np.asarray(weights), N, M, bias)
    basis = osamp.basis_states(N)
    lp = orbm.cpx_rbm_logpsi(basis, W, b)
    pex = np.exp(mu * np.real(lp))
    init = None
    if sector:
        mask = basis.sum(1) == N // 2
        pex = np.where(mask, pex, 0.0)
        init = np.array([1] * (N // 2) + [0] * (N - N // 2), np.int32)
    pex /= pex.sum()
    # 8N Metropolis steps between emitted samples: autocorrelation is negligible, so the plain
    # chi-squared statistic applies
    cfg, counters, _ = run_sampler(W, b, C, numSamples, proposer, mu, seed, K_sweep=8 * N, init=init,
                                   refreshEvery=refreshEvery)
    ints = (cfg.astype(np.int64) * (2 ** np.arange(N))[None, :]).sum(1)
    counts = np.bincount(ints, minlength=2 ** N).astype(np.float64)
    if sector:
        assert counts[pex == 0].sum() == 0
    pval, stat = chi2_pvalue(counts, pex)
    # chains are autocorrelated between emitted samples only weakly (K=N steps per sweep); allow p > 1e-3
    assert pval > 1e-3, (pval, stat)
    assert counters[0] > 0 and 0 < counters[1] <= counters[0]
    return pval


# ---------------------------------------------------------------- orbit-averaged RBM (SymNet)
def _orbit(kind, L, args, **fac):
    import warnings
    from vmc_jax_b200.util import symmetries as psym
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if kind == "1d":
            return psym.get_orbit_1D(L, *args, **fac), osym.orbit_1d(L, *args, **fac)
        return psym.get_orbit_2D_square(L, *args, **fac), osym.orbit_2d_square(L, *args, **fac)


def check_symrbm(kind="1d", L=4, M=3, args=("translation", "reflection", "spinflip"), bias=False, B=77, seed=5, **fac):
    """SymNet(orbit, CpxRBM): log Psi and gradients (reference flat layout) vs the oracle restatement."""
    ls, (orb, f) = _orbit(kind, L, args, **fac)
    N = orb.shape[1]
    W, b = orbm.init_o1(N, M, bias, seed)
    s = rand_configs(B, N, seed + 1)
    st = K.SymTables(ls, DEV)
    dW, db, ds = dev(W), (None if b is None else dev(b)), dev(s)
    lp, wts = K.symrbm_logpsi(ds, dW, db, st, want_weights=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        ref = osym.symnet_logpsi(s, W, b, orb, f)
        # amplitudes (Im log Psi is defined modulo 2 pi); with non-trivial quantum numbers the symmetrised amplitude
        # of a configuration can vanish identically (log = -inf): compare amplitudes on the scale of the largest one
        aref, agpu = np.exp(ref - np.max(ref.real)), np.exp(host(lp) - np.max(ref.real))
    aref, agpu = np.nan_to_num(aref), np.nan_to_num(agpu)
    e1 = np.max(np.abs(agpu - aref))
    assert e1 < RTOL, e1
    ok = np.abs(aref) > 1e-3            # gradients of log Psi are only defined away from the nodes
    assert ok.sum() >= 3
    assert np.max(np.abs(host(wts)[ok].sum(1) - 1.0)) < 1e-10
    g = host(K.symrbm_grad(ds, dW, db, st, wts, 0))
    with np.errstate(divide="ignore", invalid="ignore"):
        gb, gW = osym.symnet_gradients(s, W, b, orb, f)
    parts = ([gb, 1j * gb] if bias else []) + [gW.reshape(B, -1), 1j * gW.reshape(B, -1)]
    e2 = relerr(g[ok], np.concatenate(parts, axis=1)[ok])
    assert e2 < 1e-8, e2
    return e1, e2


def check_symrbm_sampler(L=4, M=2, args=("translation", "reflection", "spinflip"), weights=None, C=592,
                         numSamples=600_000, mu=2.0, seed=4321, **fac):
    """jvmc_symrbm_mcmc vs exact probabilities of the orbit-averaged RBM, chi-squared p > 1e-3."""
    ls, (orb, f) = _orbit("1d", L, args, **fac)
    if weights is None:
        W, b = orbm.init_o1(L, M, False, 17)
    else:
        W, b = orbm.unflatten_params(np.asarray(weights), L, M, False)
    basis = osamp.basis_states(L)
    pex = np.exp(mu * np.real(osym.symnet_logpsi(basis, W, b, orb, f)))
    pex /= pex.sum()
    st = K.SymTables(ls, DEV)
    dW = dev(W)
    tables = K.rbm_tables(dW, None)
    states = torch.zeros((C, L), dtype=torch.int32, device=DEV)
    counters = torch.zeros(2, dtype=torch.int64, device=DEV)
    spc = (numSamples + C - 1) // C
    cfg = host(K.symrbm_mcmc(states, dW, None, st, tables, seed, 0, 0, mu, 8 * L, 20 * 8 * L, spc, counters, 1))
    ints = (cfg.astype(np.int64) * (2 ** np.arange(L))[None, :]).sum(1)
    counts = np.bincount(ints, minlength=2 ** L).astype(np.float64)
    pval, stat = chi2_pvalue(counts, pex)
    assert pval > 1e-3, (pval, stat)
    c = host(counters)
    assert c[0] > 0 and 0 < c[1] <= c[0]
    return pval


# ---------------------------------------------------------------- CNN (real parameters)
def _cnn(shape, F, channels, strides, act, bias, firstLayerBias, seed):
    from vmc_jax_b200.nets.cnn import CNN
    net = CNN(F=F, channels=channels, strides=strides, actFun=act, bias=bias, firstLayerBias=firstLayerBias)
    cd = K.CnnDesc(net, shape)
    P = ocnn.num_parameters(F, channels, bias, firstLayerBias)
    assert cd.P == P
    theta = np.random.default_rng(seed).normal(size=P) * 0.4
    kw = dict(F=F, channels=channels, strides=strides, actFun=act, bias=bias, firstLayerBias=firstLayerBias)
    return cd, theta, kw


def check_cnn(shape=(6,), F=(3,), channels=(3, 2), strides=(1,), act=("elu",), bias=True, firstLayerBias=False, B=23,
              seed=3):
    """CNN log psi vs the oracle restatement; per-sample gradients vs central finite differences of the oracle."""
    cd, theta, kw = _cnn(shape, F, channels, strides, act, bias, firstLayerBias, seed)
    s = np.random.default_rng(seed + 1).integers(0, 2, (B,) + tuple(shape)).astype(np.int32)
    flat = dev(s.reshape(B, -1))
    dth = dev(theta)
    lp = host(K.cnn_logpsi(flat, dth, cd))
    ref = ocnn.cnn_logpsi(s, theta, **kw)
    e1 = np.max(np.abs(lp.real - ref) / np.maximum(np.abs(ref), 1e-3))
    assert e1 < RTOL and np.all(lp.imag == 0), e1
    g = host(K.cnn_grad(flat, dth, cd))
    assert np.all(g.imag == 0)
    eps, fd = 1e-6, np.zeros((B, cd.P))
    for k in range(cd.P):
        tp, tm = theta.copy(), theta.copy()
        tp[k] += eps
        tm[k] -= eps
        fd[:, k] = (ocnn.cnn_logpsi(s, tp, **kw) - ocnn.cnn_logpsi(s, tm, **kw)) / (2 * eps)
    e2 = np.max(np.abs(g.real - fd)) / max(np.max(np.abs(fd)), 1e-3)
    assert e2 < 1e-7, e2       # finite-difference truncation, not kernel accuracy
    return e1, e2


def check_cnn_sampler(shape=(6,), F=(3,), channels=(2,), proposer="spin_flip", C=296, numSamples=400_000, mu=2.0,
                      sector=False, seed=4321):
    cd, theta, kw = _cnn(shape, F, channels, (1,) * len(shape), ("elu",), True, False, 11)
    N = int(np.prod(shape))
    basis = osamp.basis_states(N)
    pex = np.exp(mu * ocnn.cnn_logpsi(basis.reshape((-1,) + tuple(shape)), theta, **kw))
    init = None
    if sector:
        pex = np.where(basis.sum(1) == N // 2, pex, 0.0)
        init = np.array([1] * (N // 2) + [0] * (N - N // 2), np.int32)
    pex /= pex.sum()
    states = torch.zeros((C, N), dtype=torch.int32, device=DEV)
    if init is not None:
        states[:] = dev(init)
    counters = torch.zeros(2, dtype=torch.int64, device=DEV)
    spc = (numSamples + C - 1) // C
    cfg = host(K.cnn_mcmc(states, dev(theta), cd, seed, 0, 0, proposer, mu, 8 * N, 20 * 8 * N, spc, counters))
    ints = (cfg.astype(np.int64) * (2 ** np.arange(N))[None, :]).sum(1)
    counts = np.bincount(ints, minlength=2 ** N).astype(np.float64)
    if sector:
        assert counts[pex == 0].sum() == 0
    pval, stat = chi2_pvalue(counts, pex)
    assert pval > 1e-3, (pval, stat)
    return pval


def check_cnn_inc_sampler(shape=(6, 6), F=(3, 3), channels=(3, 2), act=("elu",), proposer="spin_flip_zeroMag", C=37,
                          sweeps=6, seed=77, mu=2.0):
    """Incremental CNN sampler (csrc/cnn_inc.cu, one warp per chain, cached activations) against the generic kernel
    (full forward pass per proposal, the reference's algorithm sampler.py:327-356): same Philox counters and proposal
    rules, so the chains coincide bit for bit unless an acceptance probability sits within rounding of its uniform
    number (probability ~1e-13 per step)."""
    from vmc_jax_b200 import _lib
    cd, theta, kw = _cnn(shape, F, channels, (1,) * len(shape), act, True, False, 11)
    N = int(np.prod(shape))
    init = np.array(([1, 0] * N)[:N], np.int32)
    res = []
    for generic in (0, 1, 2):        # incremental (two warps per chain), generic, incremental with four warps per chain
        _lib.load().jvmc_cnn_set_generic(generic)
        try:
            states = torch.zeros((C, N), dtype=torch.int32, device=DEV)
            states[:] = dev(init)
            counters = torch.zeros(2, dtype=torch.int64, device=DEV)
            cfg = host(K.cnn_mcmc(states, dev(theta), cd, seed, 5, 3, proposer, mu, N, 2 * N, sweeps, counters))
            res.append((cfg, host(counters), host(states)))
        finally:
            _lib.load().jvmc_cnn_set_generic(0)
    for other in res[1:]:
        assert np.array_equal(res[0][0], other[0])
        assert np.array_equal(res[0][1], other[1]) and np.array_equal(res[0][2], other[2])
    assert res[0][1][0] == C * (2 + sweeps) * N and 0 < res[0][1][1] <= res[0][1][0]
    return res[0][1]


def check_cnn_eloc(strings, shape=(4, 4), F=(3, 3), channels=(3, 2), act=("elu",), bias=True, B=41, seed=5, args=()):
    """Fused incremental CNN local energy (jvmc_cnn_eloc_bfo) vs the reference's algorithm on the oracle
    (s' -> psi(s'), operator/base.py:166-192) and vs the generic device route."""
    from vmc_jax_b200 import _lib
    cd, theta, kw = _cnn(shape, F, channels, (1,) * len(shape), act, bias, False, seed)
    N = int(np.prod(shape))
    tab = obfo.Tables(strings)
    s = rand_configs(B, N, seed + 1)
    f = lambda x: ocnn.cnn_logpsi(x.reshape((-1,) + tuple(shape)), theta, **kw).astype(np.complex128)
    ref = obfo.get_O_loc(tab, s, f, *args)
    ds, dth = dev(s), dev(theta)
    dt = op_tables_to_device(tab)
    pref = dev(tab.eval_prefactors(*args))
    res = K.cnn_eloc(ds, dth, cd, dt, pref)
    assert res is not None, "the net is inside the incremental kernel's scope"
    e_fused, err = res
    assert int(err.item()) == 0
    sp, m, cnt = K.bfo_s_primes(ds, dt, pref)
    e_gen = host(K.oloc_reduce(m, K.cnn_logpsi(ds, dth, cd), K.cnn_logpsi(sp, dth, cd).reshape(m.shape)))
    r1, r2 = relerr(host(e_fused), ref), relerr(e_gen, ref)
    assert r1 < RTOL and r2 < RTOL, (r1, r2)
    _lib.load().jvmc_cnn_set_generic(1)
    try:
        assert K.cnn_eloc(ds, dth, cd, dt, pref) is None
    finally:
        _lib.load().jvmc_cnn_set_generic(0)
    return r1


def smoke_check():
    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    check_logpsi(N=8, M=16, B=64)
    check_s_primes(obfo.tfim_strings((8,), -0.7), 8, B=32)
    check_eloc(obfo.tfim_strings((8,), -0.7), N=8, M=16, B=64)
    check_moments_gram(N=5, M=16, B=100)
    check_tdvp_solve()
    check_gram_T(N=5, M=16, B=70)
    check_sampler_chi2(N=4, M=2, numSamples=200_000, C=148)
