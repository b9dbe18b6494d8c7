"""Tensor-level wrappers around the C ABI (one function per entry point of include/jvmc_b200.h).

torch is used for device memory and streams only; all arithmetic of the hot path happens in the
kernels of libjvmc_b200.so.  Complex tensors are complex128, configurations int32."""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import ptr, call

CPX = torch.complex128
F64 = torch.float64
I32 = torch.int32

PROPOSER_IDS = {"spin_flip": 0, "spin_flip_Z2": 1, "spin_flip_zeroMag": 2}


def _c(t, dtype=None):
    if t is None:
        return None
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if t.is_complex():
        t = t.resolve_conj()      # torch conjugates lazily: kernels read raw memory
    return t.contiguous()


def rbm_logpsi(s, W, b=None, want_tau=True):
    """s int32[B,N]; W complex128[N,M]; b complex128[M]|None -> logpsi[B], tau[B,M]|None."""
    s = _c(s, I32)
    W = _c(W, CPX)
    b = _c(b, CPX)
    B, N = s.shape
    M = W.shape[1]
    assert W.shape[0] == N
    logpsi = torch.empty(B, dtype=CPX, device=s.device)
    tau = torch.empty((B, M), dtype=CPX, device=s.device) if want_tau else None
    call("jvmc_rbm_logpsi", ptr(s), B, N, M, ptr(W), ptr(b), ptr(logpsi), ptr(tau))
    return logpsi, tau


def rbm_tables(W, b=None):
    W = _c(W, CPX)
    b = _c(b, CPX)
    N, M = W.shape
    n = _lib.load().jvmc_rbm_tables_elems(N, M)
    tables = torch.empty(n, dtype=CPX, device=W.device)
    call("jvmc_rbm_tables", N, M, ptr(W), ptr(b), ptr(tables))
    return tables


def rbm_mcmc(states, W, b, tables, seed, step0, chain0, proposer, mu, sweepSteps, thermSteps, numSamplesPerChain,
             counters, refreshEvery=1):
    """states int32[C,N] updated in place; returns configs int32[numSamplesPerChain*C, N]."""
    assert states.dtype == I32 and states.is_contiguous()
    C, N = states.shape
    W = _c(W, CPX)
    b = _c(b, CPX)
    M = W.shape[1]
    out = torch.empty((numSamplesPerChain * C, N), dtype=I32, device=states.device)
    pid = PROPOSER_IDS[proposer] if isinstance(proposer, str) else int(proposer)
    call("jvmc_rbm_mcmc", ptr(states), C, N, M, ptr(W), ptr(b), ptr(tables), ctypes.c_ulonglong(seed),
         ctypes.c_ulonglong(step0), chain0, pid, float(mu), int(sweepSteps), int(thermSteps),
         int(numSamplesPerChain), int(refreshEvery), ptr(out), ptr(counters))
    return out


class OpTables:
    """Device copy of the compiled operator tables (BranchFreeOperator.compile)."""

    def __init__(self, idx, mp, matEls, fermi, isDiag, device):
        self.numOps, self.len = idx.shape
        self.lDim = mp.shape[2]
        self.idx = torch.as_tensor(idx, dtype=I32).contiguous().to(device)
        self.map = torch.as_tensor(mp, dtype=I32).contiguous().to(device)
        self.matEls = torch.as_tensor(matEls, dtype=CPX).contiguous().to(device)
        self.fermi = torch.as_tensor(fermi, dtype=I32).contiguous().to(device)
        self.isDiag = torch.as_tensor(isDiag, dtype=torch.uint8).contiguous().to(device)
        self.numDiag = int(torch.as_tensor(isDiag).sum())

    def args(self):
        return (self.numOps, self.len, self.lDim, ptr(self.idx), ptr(self.map), ptr(self.matEls), ptr(self.fermi),
                ptr(self.isDiag))


def bfo_s_primes(s, tab, pref):
    """Operator.get_s_primes on the device: returns sp[B*Kmax,N], matEl[B,Kmax], count[B]."""
    s = _c(s, I32)
    pref = _c(pref, CPX)
    B, N = s.shape
    dev = s.device
    mAll = torch.empty((B, tab.numOps), dtype=CPX, device=dev)
    choice = torch.empty((B, tab.numOps), dtype=I32, device=dev)
    count = torch.empty(B, dtype=I32, device=dev)
    maxCount = torch.zeros(1, dtype=I32, device=dev)
    call("jvmc_bfo_matels", ptr(s), B, N, *tab.args(), tab.numDiag, ptr(pref), ptr(mAll), ptr(choice), ptr(count),
         ptr(maxCount))
    Kmax = int(maxCount.item())   # data-dependent output shape: the one host sync of this API (base.py:157)
    sp = torch.empty((B * Kmax, N), dtype=I32, device=dev)
    matEl = torch.empty((B, Kmax), dtype=CPX, device=dev)
    call("jvmc_bfo_emit", ptr(s), B, N, *tab.args(), ptr(mAll), ptr(choice), ptr(count), Kmax, ptr(sp), ptr(matEl))
    return sp, matEl, count


def oloc_reduce(matEl, logPsiS, logPsiSP):
    matEl = _c(matEl, CPX)
    logPsiS = _c(logPsiS, CPX)
    logPsiSP = _c(logPsiSP, CPX)
    B, K = matEl.shape
    out = torch.empty(B, dtype=CPX, device=matEl.device)
    call("jvmc_oloc_reduce", ptr(matEl), ptr(logPsiS), ptr(logPsiSP), B, K, ptr(out))
    return out


class CnnDesc:
    """Host int descriptor of a CNN on a given lattice (see include/jvmc_b200.h) + its parameter count."""

    def __init__(self, net, sampleShape):
        d = net.descriptor(tuple(sampleShape))
        self.n = len(d)
        self.arr = (ctypes.c_int * self.n)(*d)
        P = ctypes.c_int(0)
        _lib.check(_lib.load().jvmc_cnn_num_parameters(self.arr, self.n, ctypes.byref(P)), "jvmc_cnn_num_parameters")
        self.P = P.value
        self.N = int(np.prod(sampleShape))


def cnn_logpsi(s, theta, cd):
    s = _c(s, I32)
    theta = _c(theta, F64)
    B = s.shape[0]
    out = torch.empty(B, dtype=CPX, device=s.device)
    call("jvmc_cnn_logpsi", cd.arr, cd.n, ptr(theta), ptr(s), B, ptr(out))
    return out


def cnn_grad(s, theta, cd):
    s = _c(s, I32)
    theta = _c(theta, F64)
    B = s.shape[0]
    out = torch.empty((B, cd.P), dtype=CPX, device=s.device)
    call("jvmc_cnn_grad", cd.arr, cd.n, ptr(theta), ptr(s), B, ptr(out))
    return out


def cnn_mcmc(states, theta, cd, seed, step0, chain0, proposer, mu, sweepSteps, thermSteps, numSamplesPerChain, counters):
    assert states.dtype == I32 and states.is_contiguous()
    C, N = states.shape
    theta = _c(theta, F64)
    out = torch.empty((numSamplesPerChain * C, N), dtype=I32, device=states.device)
    pid = PROPOSER_IDS[proposer] if isinstance(proposer, str) else int(proposer)
    call("jvmc_cnn_mcmc", cd.arr, cd.n, ptr(theta), ptr(states), C, ctypes.c_ulonglong(seed), ctypes.c_ulonglong(step0),
         chain0, pid, float(mu), int(sweepSteps), int(thermSteps), int(numSamplesPerChain), ptr(out), ptr(counters))
    return out


def cnn_eloc(s, theta, cd, tab, pref):
    """Fused local energy of a stride-1 CNN (incremental psi(s')/psi(s), csrc/cnn_inc.cu); None when the net or the
    operator is outside the kernel's scope (the caller then takes the generic s' route)."""
    s = _c(s, I32)
    theta = _c(theta, F64)
    pref = _c(pref, CPX)
    B = s.shape[0]
    out = torch.empty(B, dtype=CPX, device=s.device)
    err = torch.zeros(1, dtype=I32, device=s.device)
    _lib.require_cuda()
    rc = _lib.load().jvmc_cnn_eloc_bfo(cd.arr, cd.n, ptr(theta), ptr(s), B, *tab.args(), tab.numDiag, ptr(pref), ptr(out),
                                       ptr(err), _lib.stream())
    if rc == -3:      # JVMC_ERR_UNSUPPORTED
        return None
    _lib.check(rc, "jvmc_cnn_eloc_bfo")
    _lib.LAUNCHES += 1
    return out, err


class SymTables:
    """Device copies of a LatticeSymmetry in index form (perm, sign, inverse map, factors)."""

    def __init__(self, sym, device):
        self.G, self.N = sym.perm.shape
        self.pmap = torch.as_tensor(np.ascontiguousarray(sym.perm), dtype=I32).to(device).contiguous()
        self.esgn = torch.as_tensor(np.ascontiguousarray(sym.sign), dtype=I32).to(device).contiguous()
        self.qmap = torch.as_tensor(np.ascontiguousarray(sym.inv), dtype=I32).to(device).contiguous()
        self.fac = torch.as_tensor(np.asarray(sym.factor).astype(np.complex128)).to(device).contiguous()


def symrbm_logpsi(s, W, b, st, want_weights=False):
    """log Psi of the orbit-averaged RBM; optionally the normalised group weights f_g psi_g / sum [B, G]."""
    s = _c(s, I32)
    W = _c(W, CPX)
    b = _c(b, CPX)
    B, N = s.shape
    M = W.shape[1]
    out = torch.empty(B, dtype=CPX, device=s.device)
    wts = torch.empty((B, st.G), dtype=CPX, device=s.device) if want_weights else None
    call("jvmc_symrbm_logpsi", ptr(s), B, N, M, st.G, ptr(W), ptr(b), ptr(st.pmap), ptr(st.esgn), ptr(st.fac), ptr(out),
         ptr(wts))
    return (out, wts) if want_weights else out


def symrbm_grad(s, W, b, st, wts, layout):
    s = _c(s, I32)
    W = _c(W, CPX)
    b = _c(b, CPX)
    B, N = s.shape
    M = W.shape[1]
    Pc = (M if b is not None else 0) + N * M
    out = torch.empty((B, 2 * Pc if layout == 0 else Pc), dtype=CPX, device=s.device)
    call("jvmc_symrbm_grad", ptr(s), B, N, M, st.G, ptr(W), ptr(b), ptr(st.pmap), ptr(st.esgn), ptr(st.fac), ptr(wts),
         int(layout), ptr(out))
    return out


def symrbm_mcmc(states, W, b, st, tables, seed, step0, chain0, mu, sweepSteps, thermSteps, numSamplesPerChain, counters,
                refreshEvery=1):
    """states int32[C,N] updated in place; returns configs int32[numSamplesPerChain*C, N] (propose_spin_flip)."""
    assert states.dtype == I32 and states.is_contiguous()
    C, N = states.shape
    W = _c(W, CPX)
    b = _c(b, CPX)
    M = W.shape[1]
    out = torch.empty((numSamplesPerChain * C, N), dtype=I32, device=states.device)
    call("jvmc_symrbm_mcmc", ptr(states), C, N, M, st.G, ptr(W), ptr(b), ptr(st.pmap), ptr(st.esgn), ptr(st.qmap),
         ptr(st.fac), ptr(tables), ctypes.c_ulonglong(seed), ctypes.c_ulonglong(step0), chain0, float(mu),
         int(sweepSteps), int(thermSteps), int(numSamplesPerChain), int(refreshEvery), ptr(out), ptr(counters))
    return out


def rbm_eloc(s, tau, tables, tab, pref):
    s = _c(s, I32)
    tau = _c(tau, CPX)
    pref = _c(pref, CPX)
    B, N = s.shape
    M = tau.shape[1]
    out = torch.empty(B, dtype=CPX, device=s.device)
    err = torch.zeros(1, dtype=I32, device=s.device)
    call("jvmc_rbm_eloc_bfo", ptr(s), ptr(tau), B, N, M, ptr(tables), *tab.args(), tab.numDiag, ptr(pref), ptr(out),
         ptr(err))
    return out, err


def rbm_grad(s, tau, hasBias, layout):
    s = _c(s, I32)
    tau = _c(tau, CPX)
    B, N = s.shape
    M = tau.shape[1]
    Mb = M if hasBias else 0
    P = 2 * (Mb + N * M) if layout == 0 else Mb + N * M
    out = torch.empty((B, P), dtype=CPX, device=s.device)
    call("jvmc_rbm_grad", ptr(s), ptr(tau), B, N, M, int(hasBias), int(layout), ptr(out))
    return out


def rbm_moments(s, tau, wgt, hasBias, conjTau):
    """out[r,j] = sum_n wgt_n sigma_nr (conj?)tau_nj, r = [bias pseudo-site,] sites."""
    s = _c(s, I32)
    tau = _c(tau, CPX)
    wgt = _c(wgt, CPX)
    B, N = s.shape
    M = tau.shape[1]
    R = N + (1 if hasBias else 0)
    chunks = _lib.load().jvmc_rbm_moments_chunks(B)
    ws = torch.empty(chunks * R * M, dtype=CPX, device=s.device)
    out = torch.empty((R, M), dtype=CPX, device=s.device)
    call("jvmc_rbm_moments", ptr(s), ptr(tau), ptr(wgt), B, N, M, int(hasBias), int(conjTau), ptr(ws), ptr(out))
    return out


def rbm_krmatvec(s, tau, x, hasBias, conjTau=False):
    """out[n] = sum_{r,j} sigma_nr (conj?)tau_nj x[r,j]  (O_n . x)."""
    s = _c(s, I32)
    tau = _c(tau, CPX)
    x = _c(x, CPX)
    B, N = s.shape
    M = tau.shape[1]
    out = torch.empty(B, dtype=CPX, device=s.device)
    call("jvmc_rbm_krmatvec", ptr(s), ptr(tau), ptr(x), B, N, M, int(hasBias), int(conjTau), ptr(out))
    return out


def pack_sigma(s, hasBias):
    s = _c(s, I32)
    B, N = s.shape
    R = N + (1 if hasBias else 0)
    words = (B + 31) // 32
    sigT = torch.empty((R, max(words, 1)), dtype=I32, device=s.device)
    call("jvmc_pack_sigma", ptr(s), B, N, int(hasBias), ptr(sigT))
    return sigT


def rbm_gram_S(Y, sigT, mu, alpha, kappa, out=None, tile=0):
    """A[(r,j),(r',l)] (row-major [R*M, R*M] complex128)."""
    Y = _c(Y, CPX)
    B, M = Y.shape
    R = sigT.shape[0]
    mu = _c(mu, CPX)
    if out is None:
        out = torch.empty((R * M, R * M), dtype=CPX, device=Y.device)
    call("jvmc_rbm_gram_S", ptr(Y), B, M, R, ptr(sigT), ptr(mu), float(alpha), float(kappa), ptr(out), int(tile))
    return out


_I8_TILES = {}


NARROW_DIAGONAL_TILES = os.environ.get("JVMC_I8_NARROW", "1") != "0"


def i8_tile_list(M, rows=64, cols=40):
    """Tiles (rowGroup, colGroup, NC, jlo, jhi) covering every (j, l <= j) exactly once: `rows` complex rows starting at
    complex row 4 rowGroup, NC real columns starting at real column 8 colGroup (the kernel's full tile is 128 x 80 real
    columns = 64 x 40 complex).

    Full row blocks are anchored at the END of the row range (the block ending at row e needs the columns [0, e)), so
    the ragged remainder sits at rows [0, o) where it needs few columns; the last column tile of every row block is
    only as wide as the diagonal requires (multiple of 16 real columns).  M = 400: 39 tiles, 2912 real columns
    (ideal 2500; full-width tiles 3120; remainder at the bottom 3680)."""
    M4 = (M + 3) // 4 * 4            # row blocks start on 8-real-row (4 complex) boundaries of the digit layout
    nFull = M4 // rows
    o = M4 - rows * nFull
    blocks = ([(0, 0, min(o, M))] if o > 0 else []) + \
        [(o + rows * b, o + rows * b, min(o + rows * (b + 1), M)) for b in range(nFull)]
    tl = []
    for start, lo, hi in blocks:
        need = hi                                    # complex columns [0, hi)
        for J in range((need + cols - 1) // cols):
            c0 = cols * J
            width = min(cols, need - c0)             # complex columns of this tile
            nc = min(2 * cols, (2 * width + 15) // 16 * 16) if NARROW_DIAGONAL_TILES else 2 * cols
            tl.append((start // 4, (2 * c0) // 8, nc, lo, hi))
    return tl


def _i8_tiles(M, device):
    key = (M, str(device))
    if key not in _I8_TILES:
        r, c = ctypes.c_int(0), ctypes.c_int(0)
        _lib.check(_lib.load().jvmc_i8_tile_shape(ctypes.byref(r), ctypes.byref(c)), "jvmc_i8_tile_shape")
        tl = [t + (0, 0, 0) for t in i8_tile_list(M, r.value // 2, c.value // 2)]
        _I8_TILES[key] = torch.tensor(tl, dtype=I32, device=device).contiguous()
    return _I8_TILES[key]


def rbm_gram_S_i8(Y, sigT, mu, alpha, kappa, out=None, accumulate=False, pair_groups=None, after_group=None):
    """A (as rbm_gram_S) on the tcgen05 INT8 tensor cores via error-free splitting of Y (fp64-equivalent).
    accumulate: A += alpha G (no mean correction) into ``out``.
    pair_groups: list of (pair0, npairs) ranges of site pairs (order: r0 ascending, r1 = r0..R-1) launched one after the
    other; ``after_group(k)`` is called when the launches of group k are enqueued -- block rows of the upper block
    triangle of A are final group by group, which lets the caller reduce them over ranks while the next group runs."""
    Y = _c(Y, CPX)
    B, M = Y.shape
    R = sigT.shape[0]
    mu = _c(mu, CPX)
    lib = _lib.load()
    nch, nzg, nbytes = ctypes.c_longlong(0), ctypes.c_int(0), ctypes.c_longlong(0)
    _lib.check(lib.jvmc_i8_layout(B, M, ctypes.byref(nch), ctypes.byref(nzg), ctypes.byref(nbytes)), "jvmc_i8_layout")
    dev = Y.device
    digits = torch.zeros(nbytes.value, dtype=torch.int8, device=dev)
    colmax = torch.empty(2 * M, dtype=torch.int64, device=dev)
    scale = torch.empty(2 * M, dtype=F64, device=dev)
    call("jvmc_i8_slice", ptr(Y), B, M, ptr(colmax), ptr(scale), ptr(digits))
    tiles = _i8_tiles(M, dev)
    if out is None:
        out = torch.empty((R * M, R * M), dtype=CPX, device=dev)
    groups = [(0, 0)] if pair_groups is None else list(pair_groups)
    for k, (p0, npairs) in enumerate(groups):
        call("jvmc_rbm_gram_S_i8", ptr(digits), ptr(scale), B, M, R, ptr(sigT), ptr(tiles), int(tiles.shape[0]), ptr(mu),
             float(alpha), float(kappa), int(bool(accumulate)), int(p0), int(npairs), ptr(out))
        if after_group is not None:
            after_group(k)
    return out


def pairs_before_row(R, r0):
    """number of site pairs (r, r' >= r) with r < r0 in the launch order of jvmc_rbm_gram_S_i8"""
    return r0 * R - r0 * (r0 - 1) // 2


# ---- which Gram kernel.  The int8 digits resolve 2^-41 of each column's MAXIMUM, and the dropped digit pairs (level >= 7)
# are ~2^-40 c_z c_z' per sample; summed over B samples with random signs, the error of A_zz' relative to its natural size
# sqrt(A_zz A_z'z') is ~ 0.3 * 4 * 2^-40 * r_z r_z' / sqrt(B) with r_z = max_n|Z_nz| / rms_n(Z_nz) (prefactor measured on
# B200: 3-8e-14 on uniform test data with r ~ 2, 2e-10 for two columns with r ~ 200 at B = 30000).  tau = tanh(theta) is
# heavy-tailed next to the poles of tanh: at config 2 (synthetic weights) r reaches 40-60 and the prediction 1e-11.  Above
# I8_MAX_PREDICTED_ERROR the few samples that carry the outliers (an entry above I8_OUTLIER_T column rms; ~150 of 66 304 at
# config 2) are taken out and go through the exact fp64 DMMA kernel, the bulk (r <= I8_OUTLIER_T) through the int8 kernel;
# if more than 1/16 of the samples are outliers the whole Gram runs in fp64.
I8_MAX_PREDICTED_ERROR = float(os.environ.get("JVMC_I8_MAX_ERROR", "2e-12"))
I8_OUTLIER_T = 16.0
LAST_GRAM = {"backend": None, "tail_ratios": None, "predicted_error": None, "outlier_rows": 0, "grouped": False}


def i8_tail_ratios(Y, return_scratch=False):
    """max_n |Z_nz| / rms_n(Z_nz) for the 2M real columns of Y (device tensor)."""
    Y = _c(Y, CPX)
    B, M = Y.shape
    scratch = torch.empty(4 * M, dtype=F64, device=Y.device)
    ratios = torch.empty(2 * M, dtype=F64, device=Y.device)
    call("jvmc_i8_tail_ratios", ptr(Y), B, M, ptr(scratch), ptr(ratios))
    return (ratios, scratch) if return_scratch else ratios


def i8_predicted_error(r1, r2, B):
    return 0.3 * 4.0 * 2.0 ** -40 * r1 * r2 / max(B, 1) ** 0.5


def rbm_gram_S_auto(Y, sigT, mu, alpha, kappa, out=None, s=None, hasBias=False, pair_groups=None, after_group=None):
    """A by the int8 tensor-core kernel; heavy-tailed data (see above) is split: outlier samples through the fp64 DMMA
    kernel, the bulk through the int8 kernel, so that the entry-wise error stays below I8_MAX_PREDICTED_ERROR.
    s: configurations [B, N] (needed to re-pack the signs of the outlier samples; without it the whole matrix falls back to
    fp64 when the prediction is too large).  Costs one pass over Y and one small read-back per Gram."""
    Y = _c(Y, CPX)
    B, M = Y.shape
    r, scratch = i8_tail_ratios(Y, return_scratch=True)
    top = torch.topk(r, 2 if r.numel() > 1 else 1).values.tolist()
    r1, r2 = top[0], top[-1]
    err = i8_predicted_error(r1, r2, B)
    LAST_GRAM.update(backend="i8", tail_ratios=(r1, r2), predicted_error=err, outlier_rows=0, grouped=pair_groups is not None)
    if err <= I8_MAX_PREDICTED_ERROR:
        return rbm_gram_S_i8(Y, sigT, mu, alpha, kappa, out, pair_groups=pair_groups, after_group=after_group)
    if s is not None:
        flag = torch.empty(B, dtype=torch.uint8, device=Y.device)
        call("jvmc_i8_outlier_rows", ptr(Y), B, M, ptr(scratch), float(I8_OUTLIER_T), ptr(flag))
        idx = torch.nonzero(flag).reshape(-1)
        nout = int(idx.numel())
        bulk_err = i8_predicted_error(I8_OUTLIER_T, I8_OUTLIER_T, B)
        if nout == 0 and bulk_err <= max(I8_MAX_PREDICTED_ERROR, 2e-11):
            # nothing stands out by more than I8_OUTLIER_T column rms (small B: the prediction is dominated by 1/sqrt(B))
            LAST_GRAM.update(predicted_error=bulk_err)
            return rbm_gram_S_i8(Y, sigT, mu, alpha, kappa, out, pair_groups=pair_groups, after_group=after_group)
        if 0 < nout <= B // 16 and bulk_err <= max(I8_MAX_PREDICTED_ERROR, 2e-11):
            LAST_GRAM.update(backend="i8+dmma", outlier_rows=nout, predicted_error=bulk_err)
            sig_out = pack_sigma(_c(s, I32)[idx].contiguous(), hasBias)
            A = rbm_gram_S(Y[idx].contiguous(), sig_out, mu, alpha, kappa, out)      # exact, with the mean correction
            Yb = Y.clone()
            Yb[idx] = 0
            return rbm_gram_S_i8(Yb, sigT, mu, alpha, kappa, A, accumulate=True, pair_groups=pair_groups,
                                 after_group=after_group)
    LAST_GRAM.update(backend="dmma", grouped=False)
    return rbm_gram_S(Y, sigT, mu, alpha, kappa, out)


def rbm_gram_T(s, tau, p, mu, hasBias, scale=1.0, part=None, v=None):
    """Centred tangent kernel T = scale * Obar Obar^dagger [B,B] from the Khatri-Rao factors (no dense O).
    part = (k, n): only the k-th of n equal ranges of the tile pairs is computed, into a zero-initialised T (the ranks'
    parts add up to the full kernel: one SUM all-reduce).  v = O . conj(mu) per sample, if the caller has it already
    (ranks evaluate it for their own samples and gather it)."""
    s = _c(s, I32)
    tau = _c(tau, CPX)
    p = _c(p, F64)
    mu = _c(mu, CPX)
    B, N = s.shape
    M = tau.shape[1]
    R = N + (1 if hasBias else 0)
    wordsR = (R + 31) // 32
    sigR = torch.empty((B, wordsR), dtype=I32, device=s.device)
    call("jvmc_pack_sigma_rows", ptr(s), B, N, int(hasBias), ptr(sigR))
    v = rbm_krmatvec(s, tau, mu.conj().contiguous(), hasBias) if v is None else _c(v, CPX)
    c = (mu.conj() * mu).real.sum().reshape(1).contiguous()
    if part is None:
        T = torch.empty((B, B), dtype=CPX, device=s.device)
        t0, nt = 0, 0
    else:
        T = torch.zeros((B, B), dtype=CPX, device=s.device)
        tiles = int(_lib.load().jvmc_rbm_gram_T_tiles(B))
        k, n = part
        t0 = tiles * k // n
        nt = tiles * (k + 1) // n - t0
        if nt == 0:
            return T
    call("jvmc_rbm_gram_T", ptr(tau), B, M, R, ptr(sigR), ptr(p), ptr(v), ptr(c), float(scale), int(t0), int(nt), ptr(T))
    return T


def hermitian_pack_blocks(A, M, packed=None, unpack=False):
    """Upper block triangle of A (row suffixes [r M, Pc)) <-> packed 1-d buffer."""
    assert A.dtype == CPX and A.is_contiguous() and A.shape[0] == A.shape[1]
    Pc = int(A.shape[0])
    n = _lib.load().jvmc_hermitian_packed_elems(Pc, int(M))
    if packed is None:
        packed = torch.empty(n, dtype=CPX, device=A.device)
    assert packed.numel() == n
    call("jvmc_hermitian_pack_blocks", ptr(A), Pc, int(M), ptr(packed), int(bool(unpack)))
    return packed


def hermitian_pack_rows(A, M, row0, nrows, packed, unpack=False):
    """rows [row0, row0 + nrows) of the upper block triangle <-> their slice of the packed buffer (packed: the WHOLE buffer)."""
    assert A.dtype == CPX and A.is_contiguous() and A.shape[0] == A.shape[1]
    call("jvmc_hermitian_pack_rows", ptr(A), int(A.shape[0]), int(M), int(row0), int(nrows), ptr(packed), int(bool(unpack)))
    return packed


def hermitian_packed_offset(Pc, M, r):
    """first element of block row r in the packed upper block triangle"""
    return M * (r * Pc - M * r * (r - 1) // 2)


def hermitian_mirror_blocks(A, M):
    """In place: A[i][k] = conj(A[k][i]) below the block diagonal (blocks M x M)."""
    assert A.dtype == CPX and A.is_contiguous() and A.shape[0] == A.shape[1]
    call("jvmc_hermitian_mirror_blocks", ptr(A), int(A.shape[0]), int(M))
    return A


def expand_S(A, M, N, hasBias, mode, shift):
    """Column-major S = q(S0) in the reference's flat layout (returned as a [P,P] tensor holding S^T)."""
    Mb = M if hasBias else 0
    P = 2 * (Mb + N * M)
    out = torch.empty((P, P), dtype=F64 if mode == 0 else CPX, device=A.device)
    call("jvmc_expand_S", ptr(A), M, N, int(hasBias), int(mode), float(shift), ptr(out))
    return out


EIGH_MAX_N = 32768     # largest n accepted by cuSOLVER 11.7 (CUDA 12.9) dense eigensolvers, probed with tools/eigh_probe.cu
# above this size eigh_inplace tears the problem in two (eigh_large); JVMC_EIGH_TEAR_ABOVE lowers it for tests
EIGH_TEAR_ABOVE = int(os.environ.get("JVMC_EIGH_TEAR_ABOVE", str(EIGH_MAX_N)))


def eigh_inplace(St, check=True):
    """Eigen-decomposition of the Hermitian matrix whose column-major image is ``St`` (i.e. St = S^T as
    a row-major tensor; lower triangle of S referenced).  Overwrites St with the eigenvectors: row k of
    the result is eigenvector k.  Returns (ev ascending, Vt, devInfo); raises if cuSOLVER's devInfo is non-zero."""
    n = St.shape[0]
    isC = St.is_complex()
    lib = _lib.load()
    d = ctypes.c_longlong(0)
    h = ctypes.c_longlong(0)
    _lib.require_cuda()
    if n > EIGH_TEAR_ABOVE:
        # cuSOLVER's dense eigensolvers stop at n = 32768: one level of Cuppen tearing on the tridiagonal matrix
        w, Vt = eigh_large(St)
        return w, Vt, torch.zeros(1, dtype=I32, device=St.device)
    _lib.check(lib.jvmc_eigh_workspace(n, int(isC), ctypes.byref(d), ctypes.byref(h)), "jvmc_eigh_workspace")
    work = torch.empty(max(d.value, 8), dtype=torch.uint8, device=St.device)
    w = torch.empty(n, dtype=F64, device=St.device)
    info = torch.zeros(1, dtype=I32, device=St.device)
    call("jvmc_eigh", n, int(isC), ptr(St), ptr(w), ptr(work), d.value, ptr(info))
    if check:
        # cuSOLVER reports non-convergence / illegal arguments only through the device flag: a failed decomposition must
        # not flow into the pseudo-inverse (the reference re-solves on the host, jVMC/util/tdvp.py:156-164; there is no
        # CPU path here, so it is an error).  The callers consume ev / V on the host side right after, so this sync is
        # not an extra one on the step's critical path.
        bad = int(info.item())
        if bad != 0:
            raise RuntimeError("cuSOLVER eigen-decomposition failed (n = %d, devInfo = %d)" % (n, bad))
    return w, St, info


def tdvp_regularize(ev, VtF, snr, F, pinvTol, pinvCutoff, snrTol):
    n = ev.shape[0]
    ev = _c(ev, F64)
    VtF = _c(VtF, CPX)
    F = _c(F, CPX)
    snr = _c(snr, F64)
    pinvEv = torch.empty(n, dtype=F64, device=ev.device)
    scal = torch.empty(2, dtype=F64, device=ev.device)
    call("jvmc_tdvp_regularize", n, ptr(ev), ptr(VtF), ptr(snr), ptr(F), float(pinvTol), float(pinvCutoff),
         float(snrTol), ptr(pinvEv), ptr(scal))
    return pinvEv, scal


def tdvp_solve(St, F, D, e, w, prefactor, numSamplesGlobal, snrTol, pinvTol, pinvCutoff, useSnr=True, comm=None):
    """TDVP.solve (reference jVMC/util/tdvp.py:153-213) as ONE C-ABI call: eigh, V^dagger F, SNR from the per-sample
    projections, cutoff loop, update.  St: column-major image of q(S0) (float64 -> 'real', complex128 -> 'imag'),
    overwritten by the eigenvectors (row k of the returned Vt = eigenvector k).  D [B, n] / e [B] / w [B]: centred data
    of the gradients and local energies (D None: no SNR at all); useSnr False: SNR computed but not applied (ExactSampler).
    Returns a dict."""
    n = St.shape[0]
    mode = 1 if St.is_complex() else 0
    assert St.is_contiguous() and St.dtype in (F64, CPX)
    dev = St.device
    F = _c(F, CPX)
    B = 0
    if D is not None:
        D, e, w = _c(D, CPX), _c(e, CPX), _c(w, F64)
        B = D.shape[0]
    lib = _lib.load()
    nb = ctypes.c_longlong(0)
    _lib.require_cuda()
    _lib.check(lib.jvmc_tdvp_solve_workspace(n, mode, B, ctypes.byref(nb)), "jvmc_tdvp_solve_workspace")
    work = torch.empty(max(nb.value, 8), dtype=torch.uint8, device=dev)
    out = {k: torch.empty(n, dtype=F64, device=dev) for k in ("ev", "rhoVar", "snr", "pinvEv", "update")}
    out["VtF"] = torch.empty(n, dtype=CPX, device=dev)
    out["scal"] = torch.empty(2, dtype=F64, device=dev)
    info = torch.zeros(1, dtype=I32, device=dev)
    x = complex(prefactor)
    call("jvmc_tdvp_solve", n, mode, ptr(St), ptr(F), B, ptr(D), ptr(e), ptr(w), x.real, x.imag, float(numSamplesGlobal),
         int(bool(useSnr)), float(snrTol), float(pinvTol), float(pinvCutoff), comm, ptr(out["ev"]), ptr(out["VtF"]), ptr(out["rhoVar"]),
         ptr(out["snr"]), ptr(out["pinvEv"]), ptr(out["update"]), ptr(out["scal"]), ptr(info), ptr(work), nb.value)
    bad = int(info.item())
    if bad != 0:
        raise RuntimeError("cuSOLVER eigen-decomposition failed inside jvmc_tdvp_solve (n = %d, devInfo = %d)" % (n, bad))
    out["Vt"] = St
    if D is None:
        out["rhoVar"] = out["snr"] = None
    return out


def minsr_solve(Tt, e, rtol):
    """pinv(T, rtol, hermitian=True) @ e (reference jVMC/util/minsr.py:61-62,73-78) as one C-ABI call.  Tt: column-major
    image of T (overwritten by the eigenvectors).  Returns (x complex[n], ev)."""
    n = Tt.shape[0]
    isC = Tt.is_complex()
    assert Tt.is_contiguous() and Tt.dtype in (F64, CPX)
    dev = Tt.device
    e = _c(e, CPX)
    lib = _lib.load()
    nb = ctypes.c_longlong(0)
    _lib.require_cuda()
    _lib.check(lib.jvmc_minsr_solve_workspace(n, int(isC), ctypes.byref(nb)), "jvmc_minsr_solve_workspace")
    work = torch.empty(max(nb.value, 8), dtype=torch.uint8, device=dev)
    x = torch.empty(n, dtype=CPX, device=dev)
    ev = torch.empty(n, dtype=F64, device=dev)
    info = torch.zeros(1, dtype=I32, device=dev)
    call("jvmc_minsr_solve", n, int(isC), ptr(Tt), ptr(e), float(rtol), ptr(x), ptr(ev), ptr(info), ptr(work), nb.value)
    bad = int(info.item())
    if bad != 0:
        raise RuntimeError("cuSOLVER eigen-decomposition failed inside jvmc_minsr_solve (n = %d, devInfo = %d)" % (n, bad))
    return x, ev


def _deflate_host(d, z, rho):
    """Deflation scan of the rank-one update diag(d) + rho z z^T (d ascending), after LAPACK dlaed2: components with a
    negligible z are eigenpairs already; two close poles are rotated so that one of them drops out.  O(n), sequential,
    on the host (NumPy).  Returns (d, z after rotations, deflated mask, rotations [(i, j, c, s)])."""
    n = d.size
    eps = np.finfo(np.float64).eps
    tol = 8.0 * eps * max(float(np.max(np.abs(d))), float(np.max(np.abs(z))))
    d = d.copy()
    z = z.copy()
    defl = np.zeros(n, bool)
    rots = []
    small = rho * np.abs(z) <= tol
    prev = -1
    for j in range(n):
        if small[j]:
            defl[j] = True
            continue
        if prev >= 0:
            s_, c_ = z[prev], z[j]
            tau = float(np.hypot(c_, s_))
            t = d[j] - d[prev]
            c_, s_ = c_ / tau, -s_ / tau
            if abs(t * c_ * s_) <= tol:
                z[j], z[prev] = tau, 0.0
                rots.append((prev, j, c_, s_))
                tt = d[prev] * c_ * c_ + d[j] * s_ * s_
                d[j] = d[prev] * s_ * s_ + d[j] * c_ * c_
                d[prev] = tt
                defl[prev] = True
        prev = j
    return d, z, defl, rots


def eigh_large(St):
    """Hermitian eigen-decomposition by one level of Cuppen's divide and conquer (csrc/eigh_large.cu) for matrices beyond
    cuSOLVER's n <= 32768: St is the column-major image (lower triangle referenced) and is destroyed; returns
    (ev ascending, Vt) with row k of Vt = eigenvector k, like eigh_inplace.  Replaces jnp.linalg.eigh of the reference
    (jVMC/util/tdvp.py:153-171) at the config-2 size P_c = 40 000."""
    n = St.shape[0]
    isC = St.is_complex()
    dev = St.device
    lib = _lib.load()
    _lib.require_cuda()
    m = n // 2
    if max(m, n - m) > EIGH_MAX_N:
        raise NotImplementedError("eigh_large tears once: n <= %d" % (2 * EIGH_MAX_N))
    info = torch.zeros(1, dtype=I32, device=dev)
    verbose = os.environ.get("JVMC_EIGH_VERBOSE", "0") != "0"
    import time as _time
    marks = [("start", _time.perf_counter())]

    def mark(name):
        if verbose:
            torch.cuda.synchronize()
            marks.append((name, _time.perf_counter()))

    def chk(what):
        bad = int(info.item())
        if bad != 0:
            raise RuntimeError("cuSOLVER %s failed (n = %d, devInfo = %d)" % (what, n, bad))
    # 1. tridiagonalisation A = H T H^dagger
    nb = ctypes.c_longlong(0)
    _lib.check(lib.jvmc_hetrd_workspace(n, int(isC), ctypes.byref(nb)), "jvmc_hetrd_workspace")
    work = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    dg = torch.empty(n, dtype=F64, device=dev)
    e = torch.empty(max(n - 1, 1), dtype=F64, device=dev)
    tau = torch.empty(max(n - 1, 1), dtype=CPX if isC else F64, device=dev)
    call("jvmc_hetrd", n, int(isC), ptr(St), ptr(dg), ptr(e), ptr(tau), ptr(work), nb.value, ptr(info))
    chk("hetrd")
    del work
    mark("hetrd")
    # 2. tear: T = diag(T1', T2') + rho v v^T, v = (0..0, 1, sg, 0..0)
    beta = float(e[m - 1].item())
    rho, sg = abs(beta), (1.0 if beta >= 0 else -1.0)
    halves = []
    for lo, hi, sf, sl in ((0, m, 0.0, -rho), (m, n, -rho, 0.0)):
        k = hi - lo
        T = torch.zeros((k, k), dtype=F64, device=dev)
        call("jvmc_tridiag_dense", k, ptr(dg[lo:hi].contiguous()), ptr(e[lo:hi - 1].contiguous()) if k > 1 else None,
             sf, sl, ptr(T))
        w, Qt, _ = eigh_inplace(T)            # row r of Qt = eigenvector r  (k <= EIGH_MAX_N: cuSOLVER)
        halves.append((w, Qt))
    (w1, Q1t), (w2, Q2t) = halves
    mark("eigh of the two halves")
    if rho == 0.0:                            # decoupled already
        D = torch.cat([w1, w2])
        order = torch.argsort(D)
        Zt = torch.zeros((n, n), dtype=F64, device=dev)
        Zt[:m, :m] = Q1t
        Zt[m:, m:] = Q2t
        lam, Zt = D[order], Zt[order]
    else:
        # 3. rank-one update D + rho z z^T with z = (last row of Q1, sg * first row of Q2)
        D = torch.cat([w1, w2])
        z = torch.cat([Q1t[:, m - 1], sg * Q2t[:, 0]])
        nz = float(torch.linalg.norm(z))
        z = z / nz
        rho = rho * nz * nz
        perm = torch.argsort(D, stable=True)
        d_s, z_s, defl, rots = _deflate_host(D[perm].cpu().numpy(), z[perm].cpu().numpy(), rho)
        keep = np.where(~defl)[0]
        o = np.argsort(d_s[keep], kind="stable")    # rotations may have perturbed the order of the kept poles
        keep = keep[o]
        k = keep.size
        lam_s = d_s.copy()
        U = torch.zeros((n, n), dtype=F64, device=dev)       # as a row-major tensor: U[c, r] = (column c, row r) of U
        perm_h = perm.cpu().numpy()
        if k > 0:
            dk = torch.as_tensor(d_s[keep]).to(dev)
            zk = torch.as_tensor(z_s[keep]).to(dev)
            z2 = (zk * zk).contiguous()
            orig = torch.empty(k, dtype=I32, device=dev)
            mu = torch.empty(k, dtype=F64, device=dev)
            call("jvmc_secular_roots", k, ptr(dk), ptr(z2), float(rho), float(z2.sum().item()), ptr(orig), ptr(mu))
            lam_s[keep] = (dk[orig.long()] + mu).cpu().numpy()
        order = np.argsort(lam_s, kind="stable")              # ascending eigenvalues: column position of sorted index
        colpos = np.empty(n, np.int64)
        colpos[order] = np.arange(n)
        # deflated eigenvectors: unit vectors (rotated below)
        dfl = np.where(defl)[0]
        if dfl.size:
            U[torch.as_tensor(colpos[dfl]).to(dev), torch.as_tensor(perm_h[dfl]).to(dev)] = 1.0
        if k > 0:
            rowIdx = torch.as_tensor(perm_h[keep].astype(np.int32)).to(dev)
            colIdx = torch.as_tensor(colpos[keep].astype(np.int32)).to(dev)
            zhat = torch.empty(k, dtype=F64, device=dev)
            call("jvmc_secular_vectors", k, ptr(dk), ptr(zk), float(rho), ptr(orig), ptr(mu), ptr(rowIdx), ptr(colIdx),
                 ptr(U), n, ptr(zhat))
        if rots:
            # U = G_1 .. G_m U': the rotations acted on columns (i, j) of the basis, apply them to the rows in reverse
            ri = torch.as_tensor(np.array([perm_h[r[0]] for r in reversed(rots)], np.int32)).to(dev)
            rj = torch.as_tensor(np.array([perm_h[r[1]] for r in reversed(rots)], np.int32)).to(dev)
            cc = torch.as_tensor(np.array([r[2] for r in reversed(rots)], np.float64)).to(dev)
            ss = torch.as_tensor(np.array([r[3] for r in reversed(rots)], np.float64)).to(dev)
            call("jvmc_apply_row_rotations", n, n, ptr(U), len(rots), ptr(ri), ptr(rj), ptr(cc), ptr(ss))
        lam = torch.as_tensor(lam_s[order]).to(dev)
        mark("merge: deflation (%d of %d deflated, %d rotations), secular equation, vectors" % (n - k, n, len(rots)))
        # 4. Z = diag(Q1, Q2) U   (cuBLAS dgemm through torch.matmul; row-major images: Zt[:, :m] = Ut[:, :m] Q1t)
        Zt = torch.empty((n, n), dtype=F64, device=dev)
        torch.matmul(U[:, :m], Q1t, out=Zt[:, :m])
        torch.matmul(U[:, m:], Q2t, out=Zt[:, m:])
        del U
        mark("Z = diag(Q1, Q2) U")
    del Q1t, Q2t
    # 5. back-transformation V = H Z.  H = H_0 .. H_{n-2} acts on rows 1..n-1; the reflectors sit geqrf-style in the
    #    (n-1) x (n-1) block A[1:, :n-1].  cuSOLVER's unmtr refuses n > 32768, unmqr takes <= 16384 reflectors per call:
    #    chunks of reflectors (last chunk first), column panels of Z (real -> complex for the Hermitian case).
    Vt = torch.empty((n, n), dtype=CPX if isC else F64, device=dev)
    el = 16 if isC else 8
    panel = min(n, 4096)
    chunk = min(n - 1, 4096)
    if n > 1:
        _lib.check(lib.jvmc_unmqr_workspace(n - 1, panel, chunk, int(isC), n, n, ctypes.byref(nb)), "jvmc_unmqr_workspace")
        work = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    starts = list(range(0, n - 1, chunk))
    for c0 in range(0, n, panel):
        c1 = min(n, c0 + panel)
        if isC:
            call("jvmc_real_to_complex", (c1 - c0) * n, ptr(Zt[c0:c1]), ptr(Vt[c0:c1]))
        else:
            Vt[c0:c1] = Zt[c0:c1]
        for i0 in reversed(starts):
            i1 = min(n - 1, i0 + chunk)
            # reflectors i0..i1-1: columns i0.. of A[1:, :], acting on rows i0+1.. of the panel
            Ap = ctypes.c_void_p(St.data_ptr() + ((i0 + 1) + i0 * n) * el)
            taup = ctypes.c_void_p(tau.data_ptr() + i0 * el)
            Cp = ctypes.c_void_p(Vt.data_ptr() + (c0 * n + i0 + 1) * el)
            call("jvmc_unmqr", n - 1 - i0, c1 - c0, i1 - i0, int(isC), Ap, n, taup, Cp, n, ptr(work), nb.value, ptr(info))
    chk("unmqr")
    mark("back-transformation (unmqr, %d reflector chunks x %d column panels)" % (len(starts), (n + panel - 1) // panel))
    if verbose:
        print("eigh_large n = %d (%s):" % (n, "complex" if isC else "real"), "; ".join(
            "%s %.2f s" % (marks[i][0], marks[i][1] - marks[i - 1][1]) for i in range(1, len(marks))), flush=True)
    return lam, Vt
