"""Lattice symmetry orbits (mirror of jVMC/util/symmetries.py: LatticeSymmetry :9-38, get_orbit_2D_square :113-143,
get_orbit_1D :227-259).

Every symmetry operation is a signed permutation of the lattice sites, (O x)_i = sign_i * x[perm_i] on spins
x = 2s - 1.  The device kernels work on the (perm, sign) index form; `orbit` exposes the dense matrices in the
reference's order (rows sorted lexicographically by numpy.unique, factors of the first occurrences), so code that
indexes `LatticeSymmetry.orbit` / `.factor` sees the same arrays."""
import warnings

import numpy as np


class LatticeSymmetry:
    """orbit: int array [G, N, N] of signed permutation matrices; factor: [G] phase factors."""

    def __init__(self, orbit, factor):
        orbit = np.asarray(orbit)
        if orbit.ndim != 3 or orbit.shape[1] != orbit.shape[2]:
            raise ValueError("orbit must have shape [G, N, N]")
        nz = (orbit != 0)
        if not (nz.sum(axis=2) == 1).all() or not np.isin(orbit[nz], (-1, 1)).all():
            raise NotImplementedError("only signed permutation matrices have a device representation")
        self._orbit = orbit
        self._factor = np.asarray(factor)
        self.perm = np.argmax(nz, axis=2).astype(np.int32)                     # [G, N]: column of the entry of row i
        self.sign = np.take_along_axis(orbit, self.perm[:, :, None], axis=2)[:, :, 0].astype(np.int32)
        G, N = self.perm.shape
        self.inv = np.empty_like(self.perm)                                    # row i that reads site k
        rows = np.broadcast_to(np.arange(N, dtype=np.int32), (G, N))
        np.put_along_axis(self.inv, self.perm, rows, axis=1)

    @property
    def factor(self):
        return self._factor

    @property
    def orbit(self):
        return self._orbit

    @property
    def shape(self):
        return self._orbit.shape

    @property
    def dtype(self):
        return self._orbit.dtype


def _compose(first, then):
    """site map of (O_first @ O_then): row i reads then[first[i]]."""
    return then.ravel()[first.ravel()]


def _collect(maps, factors, n, spinflip, spinflip_factor):
    maps = np.array(maps, dtype=np.int64)
    factors = np.array(factors)
    signs = np.ones(len(maps), dtype=np.int64)
    if spinflip:
        warnings.warn("Symmetry 'spinflip' assumes that spin up/down is encoded with numerical values +1/-1.")
        maps = np.concatenate([maps, maps])
        signs = np.concatenate([signs, -signs])
        factors = np.concatenate([factors, spinflip_factor * factors])
    dense = np.zeros((len(maps), n, n), dtype=np.int32)
    dense[np.arange(len(maps))[:, None], np.arange(n)[None, :], maps] = signs[:, None]
    # the reference keeps one representative per distinct matrix, in numpy.unique's lexicographic row order
    uniq, first = np.unique(dense.reshape(len(maps), n * n), return_index=True, axis=0)
    return LatticeSymmetry(uniq.reshape(-1, n, n), factors[first])


def get_orbit_1D(L, *args, translation_factor=1., reflection_factor=1., spinflip_factor=1.):
    """Translations / reflection / global spin flip of a chain of L sites (reference :227-259)."""
    idx = np.arange(L)
    reflections = [(idx, 1.)] + ([(idx[::-1], reflection_factor)] if "reflection" in args else [])
    translations = [(idx, 1.)]
    if "translation" in args:
        translations += [(np.roll(idx, x), translation_factor ** x) for x in range(1, L)]
    maps, factors = [], []
    for mt, ft in translations:
        for mr, fr in reflections:
            maps.append(_compose(mt, mr))
            factors.append(ft * fr)
    return _collect(maps, factors, L, "spinflip" in args, spinflip_factor)


def get_orbit_2D_square(L, *args, translation_factor=1., reflection_factor=1., rotation_factor=1., spinflip_factor=1.):
    """Space group of the L x L square lattice (reference :113-143; composition order translation, rotation,
    reflection, as the reference's helper multiplies them)."""
    idx = np.arange(L * L).reshape(L, L)
    rotations = [(idx, 1.)]
    if "rotation" in args:
        rotations += [(idx[::-1, :].T, rotation_factor), (idx[::-1, ::-1], rotation_factor ** 2),
                      (idx[:, ::-1].T, rotation_factor ** 3)]
    reflections = [(idx, 1.)]
    if "reflection" in args:
        reflections += [(idx[::-1, :], reflection_factor), (idx[:, ::-1], reflection_factor),
                        (idx[::-1, ::-1], reflection_factor ** 2)]
    translations = [(idx, 1.)]
    if "translation" in args:
        translations += [(np.roll(idx, (y, x), axis=(0, 1)), translation_factor ** (x + y))
                         for x in range(L) for y in range(L) if (x, y) != (0, 0)]
    maps, factors = [], []
    for mt, ft in translations:
        for mo, fo in rotations:
            for mr, fr in reflections:
                maps.append(_compose(_compose(mt, mo), mr))
                factors.append(ft * fo * fr)
    return _collect(maps, factors, L * L, "spinflip" in args, spinflip_factor)
