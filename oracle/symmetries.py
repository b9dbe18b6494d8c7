"""CPU restatement of the reference's lattice-symmetry orbits and of the SymNet wrapper around an RBM.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/, never by the product.

Follows jVMC/util/symmetries.py:40-259 (orbit matrices O with O[i, map[i]] = 1, products in the order the
reference multiplies them, optional global sign for "spinflip", np.unique ordering with the factors of the first
occurrences) and jVMC/nets/sym_wrapper.py:8-66 (x -> O (2x - 1) -> (x + 1) // 2, net on every image, logsumexp with
the symmetry factors as weights).  Parity: no golden vector of the reference involves SymNet amplitudes; pinned by
the invariance / exact-sampling properties the reference tests assert (tests/sampler_test.py:32-75).
"""
import numpy as np

from . import rbm as orbm


def _perm_matrix(mp):
    n = mp.size
    O = np.zeros((n, n), dtype=np.int64)
    O[np.arange(n), mp.ravel()] = 1
    return O


def orbit_1d(L, *args, translation_factor=1., reflection_factor=1., spinflip_factor=1.):
    """jVMC/util/symmetries.py:227-259 -> (orbit int[G, L, L], factor[G])."""
    idx = np.arange(L)
    refl = [(idx, 1.)]
    trans = [(idx, 1.)]
    if "reflection" in args:
        refl.append((idx[::-1], reflection_factor))
    if "translation" in args:
        for x in range(1, L):
            trans.append((np.roll(idx, x), translation_factor ** x))
    orbits, factors = [], []
    for mt, ft in trans:
        for mr, fr in refl:
            orbits.append(_perm_matrix(mt) @ _perm_matrix(mr))
            factors.append(ft * fr)
    return _finish(np.array(orbits), np.array(factors), "spinflip" in args, spinflip_factor, L)


def orbit_2d_square(L, *args, translation_factor=1., reflection_factor=1., rotation_factor=1., spinflip_factor=1.):
    """jVMC/util/symmetries.py:113-143.  The reference passes (translation, reflection, rotation) to a helper whose
    parameters are named (translation, rotation, reflection) (:133 vs :40), so the product is T @ Rot @ Ref."""
    idx = np.arange(L * L).reshape(L, L)
    rot = [(idx, 1.)]
    refl = [(idx, 1.)]
    trans = [(idx, 1.)]
    if "rotation" in args:
        rot += [(idx[::-1, :].T, rotation_factor), (idx[::-1, ::-1], rotation_factor ** 2), (idx[:, ::-1].T, rotation_factor ** 3)]
    if "reflection" in args:
        refl += [(idx[::-1, :], reflection_factor), (idx[:, ::-1], reflection_factor), (idx[::-1, ::-1], reflection_factor ** 2)]
    if "translation" in args:
        for x in range(L):
            for y in range(L):
                if x == 0 and y == 0:
                    continue
                trans.append((np.roll(idx, (y, x), axis=(0, 1)), translation_factor ** (x + y)))
    orbits, factors = [], []
    for mt, ft in trans:
        for m2, f2 in rot:          # the helper's "reflection" loop receives the rotation maps
            for m3, f3 in refl:     # ... and its "rotation" loop the reflection maps
                orbits.append(_perm_matrix(mt) @ _perm_matrix(m2) @ _perm_matrix(m3))
                factors.append(ft * f2 * f3)
    return _finish(np.array(orbits), np.array(factors), "spinflip" in args, spinflip_factor, L * L)


def _finish(orbits, factors, spinflip, sf, n):
    if spinflip:
        orbits = np.concatenate([orbits, -orbits], axis=0)
        factors = np.concatenate([factors, sf * factors])
    uniq, indices = np.unique(orbits.reshape(-1, n * n), return_index=True, axis=0)
    return uniq.reshape(-1, n, n), factors[indices]


def symnet_logpsi(s, W, b, orbit, factor):
    """SymNet(orbit, CpxRBM).__call__ with avgFun_Coefficients_Exp: s int[B, N] -> complex[B]."""
    s = np.asarray(s)
    B = s.shape[0]
    x = 2 * s.reshape(B, -1) - 1
    coeffs = []
    for O in orbit:
        xs = (x @ O.T + 1) // 2
        coeffs.append(orbm.cpx_rbm_logpsi(xs, W, b))
    coeffs = np.array(coeffs)                       # [G, B]
    f = np.asarray(factor).astype(complex)[:, None]
    m = np.max(np.where(f != 0, coeffs.real, -np.inf), axis=0)
    return np.log(np.sum(f * np.exp(coeffs - m), axis=0)) + m


def symnet_gradients(s, W, b, orbit, factor):
    """d log Psi / d(b, W) per sample, complex parameters: returns (gb [B, M] or None, gW [B, N, M])."""
    s = np.asarray(s)
    B = s.shape[0]
    x = 2 * s.reshape(B, -1) - 1
    f = np.asarray(factor).astype(complex)
    amps, taus, xs = [], [], []
    for O in orbit:
        xg = x @ O.T
        th = xg @ W + (0 if b is None else b)
        amps.append(np.sum(orbm.log_cosh(th), axis=1))
        taus.append(np.tanh(th))
        xs.append(xg)
    amps = np.array(amps)
    m = np.max(amps.real, axis=0)
    w = f[:, None] * np.exp(amps - m)
    w = w / np.sum(w, axis=0)                       # [G, B]
    gW = sum(w[g][:, None, None] * xs[g][:, :, None] * taus[g][:, None, :] for g in range(len(orbit)))
    gb = None if b is None else sum(w[g][:, None] * taus[g] for g in range(len(orbit)))
    return gb, gW
