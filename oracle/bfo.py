"""Oracle: branch-free operators -- tables, s' / matrix-element enumeration, compaction,
local estimator.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

An operator is a list of strings; a string is ``(prefactor, [localop, ...])`` with the
operators written left to right as in the reference (the right-most acts first).
``prefactor`` is a number or a callable ``f(*args)``.
"""
import numpy as np


def _lop(idx, mp, me, diag, fermionic=False):
    return dict(idx=int(idx), map=np.asarray(mp, np.int32), matEls=np.asarray(me, np.complex128),
                diag=bool(diag), fermionic=bool(fermionic))


# Local operator tables, jVMC/operator/branch_free.py:17-190
def Id(idx=0, lDim=2):
    return _lop(idx, range(lDim), [1.0] * lDim, True)


def Sx(idx):
    return _lop(idx, [1, 0], [1.0, 1.0], False)


def Sy(idx):
    return _lop(idx, [1, 0], [1.0j, -1.0j], False)


def Sz(idx):
    return _lop(idx, [0, 1], [-1.0, 1.0], True)


def Sp(idx):
    return _lop(idx, [1, 0], [1.0, 0.0], False)


def Sm(idx):
    return _lop(idx, [0, 0], [0.0, 1.0], False)


def number(idx):
    return _lop(idx, [0, 1], [0.0, 1.0], True, False)


def creation(idx):
    return _lop(idx, [1, 0], [1.0, 0.0], False, True)


def annihilation(idx):
    return _lop(idx, [1, 0], [0.0, 1.0], False, True)


class Tables:
    """Result of BranchFreeOperator.compile (jVMC/operator/branch_free.py:348-441)."""

    def __init__(self, strings, lDim=2):
        self.numOps = len(strings)
        self.maxLen = max(len(ops) for _, ops in strings)
        L = self.maxLen
        self.idx = np.zeros((self.numOps, L), np.int32)
        self.map = np.zeros((self.numOps, L, lDim), np.int32)
        self.matEls = np.zeros((self.numOps, L, lDim), np.complex128)
        self.fermi = np.zeros((self.numOps, L), np.int32)
        self.prefactors = []
        diag = []
        ident = Id(lDim=lDim)
        for o, (pref, ops) in enumerate(strings):
            self.prefactors.append(pref)
            isdiag = True
            for k in range(L):
                # positions are stored reversed: right-most operator first (:381-387);
                # short strings are padded with Id(idx=0) (:402-406)
                op = ops[len(ops) - 1 - k] if k < len(ops) else ident
                if k < len(ops) and not op["diag"]:
                    isdiag = False
                self.idx[o, k] = op["idx"]
                self.map[o, k] = op["map"]
                self.matEls[o, k] = op["matEls"]
                self.fermi[o, k] = 1 if op.get("fermionic", False) else 0
            if isdiag:
                diag.append(o)
        self.diag = np.asarray(diag, np.int32)

    def eval_prefactors(self, *args):
        """arg_fun (jVMC/operator/branch_free.py:416-438): complex128[numOps]."""
        out = np.ones(self.numOps, np.complex128)
        for o, f in enumerate(self.prefactors):
            out[o] = f(*args) if callable(f) else f
        return out


def s_primes_all(tab, s, pref):
    """_get_s_primes (jVMC/operator/branch_free.py:443-487) for a batch.

    s: int [B, N] (flattened configs).  Returns sp int32 [B, numOps, N] and
    matEl complex128 [B, numOps] *before* compaction.  The diagonal merge uses an
    ascending-index sequential sum."""
    s = np.asarray(s, np.int32)
    B, N = s.shape
    sp = np.repeat(s[:, None, :], tab.numOps, axis=1).copy()
    m = np.repeat(np.asarray(pref, np.complex128)[None, :], B, axis=0).copy()
    ar = np.arange(B)
    for o in range(tab.numOps):
        c = sp[:, o, :]
        for k in range(tab.maxLen):
            i = tab.idx[o, k]
            cur = c[:, i]
            m[:, o] = m[:, o] * tab.matEls[o, k][cur]
            if tab.fermi[o, k]:
                # Jordan-Wigner sign prod_{j>i} (1-2 c_j)   (:463-466)
                sign = np.prod(1 - 2 * c[:, i + 1:], axis=1)
                m[:, o] = m[:, o] * sign
            c[ar, i] = tab.map[o, k][cur]
    if len(tab.diag) > 1:
        acc = np.zeros(B, np.complex128)
        for o in tab.diag:
            acc = acc + m[:, o]
        m[:, tab.diag[1:]] = 0.0
        m[:, tab.diag[0]] = acc
    return sp, m


def compact(sp, m):
    """Operator.get_s_primes post-processing (jVMC/operator/base.py:91-114,156-158):
    keep |m|>1e-6 in ascending op order, pad to Kmax=max count with the last op's s'
    and matEl 0.  Returns sp [B*Kmax, N], matEl [B, Kmax], numNonzero [B]."""
    B, numOps, N = sp.shape
    nz = np.abs(m) > 1e-6
    cnt = nz.sum(axis=1)
    Kmax = int(cnt.max()) if B > 0 else 0
    choice = np.where(nz, np.arange(numOps)[None, :], numOps - 1)
    choice = np.sort(choice, axis=1)[:, :Kmax]
    ar = np.arange(B)[:, None]
    mm = m[ar, choice]
    mm = np.where(np.arange(Kmax)[None, :] < cnt[:, None], mm, 0.0)
    spc = sp[ar, choice]
    return spc.reshape(B * Kmax, N), mm, cnt


def get_s_primes(tab, s, *args):
    pref = tab.eval_prefactors(*args)
    sp, m = s_primes_all(tab, s, pref)
    return compact(sp, m)


def o_loc(matEl, logPsiS, logPsiSP):
    """Operator._get_O_loc (jVMC/operator/base.py:162-164)."""
    return np.sum(matEl * np.exp(logPsiSP.reshape(matEl.shape) - logPsiS[:, None]), axis=1)


def get_O_loc(tab, s, logpsi_fn, *args, logPsiS=None):
    """Operator.get_O_loc (jVMC/operator/base.py:166-192), unbatched branch: the
    reference's algorithm -- enumerate s', full forward pass on every s'."""
    if logPsiS is None:
        logPsiS = logpsi_fn(s)
    sp, m, _ = get_s_primes(tab, s, *args)
    return o_loc(m, logPsiS, logpsi_fn(sp))


def tfim_strings(shape, g, J=-1.0, pbc=True):
    """TFIM strings in ex0 order (examples/ex0_ground_state_search.py:57-59): for each
    site J*Sz(l)Sz(l+1) then g*Sx(l); 2D: right bond, down bond, then Sx."""
    strings = []
    if len(shape) == 1:
        L = shape[0]
        for l in range(L):
            if pbc or l + 1 < L:
                strings.append((J, [Sz(l), Sz((l + 1) % L)]))
            strings.append((g, [Sx(l)]))
    else:
        Lx, Ly = shape
        for x in range(Lx):
            for y in range(Ly):
                l = x * Ly + y
                strings.append((J, [Sz(l), Sz(x * Ly + (y + 1) % Ly)]))
                strings.append((J, [Sz(l), Sz(((x + 1) % Lx) * Ly + y)]))
                strings.append((g, [Sx(l)]))
    return strings
