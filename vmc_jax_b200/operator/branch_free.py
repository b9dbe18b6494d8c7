"""Mirror of jVMC/operator/branch_free.py: local operators, operator strings and the
BranchFreeOperator whose s'/matrix-element enumeration runs in CUDA (csrc/bfo.cu)."""
import copy
import functools

import numpy as np
import torch

from .. import global_defs
from .. import kernels as K
from .base import Operator, opDtype

__all__ = ["Id", "Sx", "Sy", "Sz", "Sp", "Sm", "number", "creation", "annihilation", "OpStr", "scal_opstr",
           "LocalOp", "BranchFreeOperator"]


class LocalOp(dict):
    """Operator on one local Hilbert space: keys "idx", "map", "matEls", "diag" [, "fermionic"]
    (reference :279-308)."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    def __mul__(self, other):
        if isinstance(other, dict):
            return OpStr(self, LocalOp(**other))
        if isinstance(other, OpStr):
            return OpStr(self, *other)
        return OpStr(self, other)

    def __rmul__(self, other):
        return other * OpStr(self)


def _lop(idx, mp, me, diag, **kw):
    return LocalOp(idx=idx, map=np.asarray(mp, dtype=np.int32), matEls=np.asarray(me, dtype=opDtype), diag=diag, **kw)


# reference :17-190 -- tables of the common local operators
def Id(idx=0, lDim=2):
    return _lop(idx, list(range(lDim)), [1.] * lDim, True)


def Sx(idx):
    return _lop(idx, [1, 0], [1.0, 1.0], False)


def Sy(idx):
    return _lop(idx, [1, 0], [1.j, -1.j], False)


def Sz(idx):
    return _lop(idx, [0, 1], [-1.0, 1.0], True)


def Sp(idx):
    return _lop(idx, [1, 0], [1.0, 0.0], False)


def Sm(idx):
    return _lop(idx, [0, 0], [0.0, 1.0], False)


def number(idx):
    return _lop(idx, [0, 1], [0., 1.], True, fermionic=False)


def creation(idx):
    return _lop(idx, [1, 0], [1., 0.], False, fermionic=True)


def annihilation(idx):
    return _lop(idx, [1, 0], [0., 1.], False, fermionic=True)


def _const(*args, val=1.0, **kwargs):
    return val


def _product(*args, f1, f2, **kwargs):
    return f1(*args) * f2(*args)


def _as_factor(a):
    return a if callable(a) else functools.partial(_const, val=a)


class OpStr(tuple):
    """Operator string: optional (callable) prefactor followed by LocalOps (reference :203-254)."""

    def __new__(cls, *args):
        factors, ops = [], []
        for o in args:
            if isinstance(o, (LocalOp, dict)):
                ops.append(o)
            else:
                factors.append(_as_factor(o))
        while len(factors) > 1:
            factors[0] = functools.partial(_product, f1=factors[0], f2=factors.pop())
        return super().__new__(cls, tuple(factors + ops))

    def __mul__(self, other):
        if not isinstance(other, (tuple, OpStr)):
            other = OpStr(other)
        if len(other) > 0 and callable(other[0]):
            return OpStr(*(other[0] * self), *(other[1:]))
        return OpStr(*self, *other)

    def __rmul__(self, a):
        if isinstance(a, dict):
            return OpStr(LocalOp(**a), *self)
        new = [copy.deepcopy(o) for o in self]
        a = _as_factor(a)
        if len(new) > 0 and callable(new[0]):
            new[0] = functools.partial(_product, f1=a, f2=new[0])
        else:
            new = [a] + new
        return OpStr(*new)


def scal_opstr(a, op):
    """Add a prefactor (scalar or callable) to an operator string (reference :257-276)."""
    if not isinstance(op, (tuple, OpStr)):
        raise RuntimeError("Can add prefactors only to OpStr or tuple objects.")
    if not isinstance(op, OpStr):
        op = OpStr(*op)
    return a * op


class CompiledTables:
    """Host + device tables of a compiled BranchFreeOperator (reference compile(), :348-441)."""

    def __init__(self, ops, lDim):
        self.lDim = lDim
        self.numOps = len(ops)
        self.maxOpStrLength = max((len(op) - (1 if callable(op[0]) else 0)) for op in ops)
        L = self.maxOpStrLength
        self.idx = np.zeros((self.numOps, L), np.int32)
        self.map = np.zeros((self.numOps, L, lDim), np.int32)
        self.matEls = np.zeros((self.numOps, L, lDim), opDtype)
        self.fermionic = np.zeros((self.numOps, L), np.int32)
        self.isDiag = np.zeros(self.numOps, np.uint8)
        self.prefactor = []
        self.nondiag_sites = []
        ident = Id(lDim=lDim)
        for o, op in enumerate(ops):
            k0 = 0
            if callable(op[0]):
                self.prefactor.append((o, op[0]))
                k0 = 1
            isDiagonal = True
            touched = set()
            for k in range(L):
                kRev = len(op) - k - 1          # right-most operator first (reference :381-387)
                cur = op[kRev] if kRev >= k0 else ident
                if kRev >= k0 and not cur["diag"]:
                    isDiagonal = False
                    touched.add(int(cur["idx"]))
                self.idx[o, k] = cur["idx"]
                self.map[o, k] = np.asarray(cur["map"])
                self.matEls[o, k] = np.asarray(cur["matEls"])
                self.fermionic[o, k] = 1 if (kRev >= k0 and cur.get("fermionic", False)) else 0
            self.isDiag[o] = 1 if isDiagonal else 0
            self.nondiag_sites.append(touched)
        self.diag = np.nonzero(self.isDiag)[0].astype(np.int32)
        self._dev = None
        self._static_pref = None

    def eval_prefactors(self, *args):
        """arg_fun of the reference (:416-438): complex128[numOps], ones where no prefactor is given.
        Callables are evaluated on the host at every call (as the reference does)."""
        res = np.ones(self.numOps, dtype=opDtype)
        for i, f in self.prefactor:
            v = f(*args)
            if isinstance(v, torch.Tensor):
                v = v.item()
            res[i] = complex(v)
        return torch.as_tensor(res).to(global_defs.myDevice)

    def device_tables(self):
        if self._dev is None:
            self._dev = K.OpTables(self.idx, self.map, self.matEls, self.fermionic, self.isDiag, global_defs.myDevice)
        return self._dev

    def fused_ok(self, psi):
        """The fused (Cpx)RBM local-energy kernel covers lDim=2, non-fermionic strings flipping <= 2 sites."""
        return (getattr(psi, "khatri_rao", False) and hasattr(psi, "flip_tables") and getattr(psi, "logarithmic", False)
                and self.lDim == 2 and self.maxOpStrLength <= 16 and not self.fermionic.any()
                and all(len(t) <= 2 for t in self.nondiag_sites))

    def cnn_fused_ok(self, psi):
        """The incremental CNN local-energy kernel covers lDim=2, non-fermionic strings changing <= 2 sites."""
        return (getattr(psi, "kind", "") == "cnn" and getattr(psi, "logarithmic", False)
                and self.lDim == 2 and self.maxOpStrLength <= 16 and not self.fermionic.any()
                and all(len(t) <= 2 for t in self.nondiag_sites))


class BranchFreeOperator(Operator):
    """Operators whose strings map every basis state to at most one basis state (reference :311-487)."""

    def __init__(self, lDim=2, **kwargs):
        self.ops = []
        self.lDim = lDim
        super().__init__(**kwargs)

    def add(self, opDescr):
        """Add another operator string (reference :330-342)."""
        if not isinstance(opDescr, tuple):
            opDescr = (opDescr,)
        self.ops.append(opDescr)
        self.compiled = False

    def __iadd__(self, opDescr):
        self.add(opDescr)
        return self

    def compile(self):
        """Builds the index / map / matrix-element tables (reference :348-441)."""
        tab = CompiledTables(self.ops, self.lDim)
        self.idxC, self.mapC, self.matElsC = tab.idx, tab.map, tab.matEls
        self.fermionicC, self.diag = tab.fermionic, tab.diag
        self.maxOpStrLength, self.prefactor = tab.maxOpStrLength, tab.prefactor
        return tab
