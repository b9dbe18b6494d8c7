"""Oracle: (Cpx)RBM log-amplitudes, per-sample gradients, flat parameter layout.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  NumPy fp64, no device code.

Conventions: ``s`` is an int array [B, N] with entries in {0,1}; ``W`` is [N, M]
(Flax Dense kernel = [in, out]); ``b`` is [M] or None.
"""
import numpy as np


def log_cosh(x):
    """Follows jVMC/nets/activation_functions.py:19-22 (sign from signbit(Re x), principal
    branch of the complex log1p per hidden unit)."""
    x = np.asarray(x, dtype=np.complex128)
    sgn = -2.0 * np.signbit(x.real) + 1.0
    x = x * sgn
    return x + np.log1p(np.exp(-2.0 * x)) - np.log(2.0)


def theta(s, W, b=None):
    """Hidden pre-activation  theta = (2s-1) W + b   (jVMC/nets/rbm.py:32-38)."""
    sig = 2.0 * np.asarray(s).reshape(np.asarray(s).shape[0], -1) - 1.0
    th = sig @ W
    if b is not None:
        th = th + b
    return th


def cpx_rbm_logpsi(s, W, b=None):
    """CpxRBM.__call__ (jVMC/nets/rbm.py:30-38): sum_j log_cosh(theta_j)."""
    return np.sum(log_cosh(theta(s, np.asarray(W, np.complex128), b)), axis=-1)


def real_rbm_logpsi(s, W, b=None):
    """RBM.__call__ (jVMC/nets/rbm.py:82-88): sum_j log(cosh(theta_j)), real parameters."""
    return np.sum(np.log(np.cosh(theta(s, np.asarray(W, np.float64), b))), axis=-1)


def tau(s, W, b=None):
    return np.tanh(theta(s, W, b))


def gradients_holomorphic(s, W, b=None):
    """NQS.gradients for a holomorphic CpxRBM: flat_gradient_holo (jVMC/vqs.py:66-69).

    Per Flax leaf (sorted keys: 'bias' before 'kernel'; kernel ravelled C-order, index
    i*M+j) the reference emits [d_leaf, 1j*d_leaf]; d logpsi/dW_ij = sigma_i tanh(theta_j),
    d logpsi/db_j = tanh(theta_j).  Returns complex128 [B, P], P = 2*(N*M (+M))."""
    s2 = np.asarray(s).reshape(np.asarray(s).shape[0], -1)
    sig = 2.0 * s2 - 1.0
    t = np.tanh(theta(s2, np.asarray(W, np.complex128), b))
    gW = (sig[:, :, None] * t[:, None, :]).reshape(s2.shape[0], -1)
    parts = []
    if b is not None:
        parts += [t, 1j * t]
    parts += [gW, 1j * gW]
    return np.concatenate(parts, axis=1)


def gradients_real(s, W, b=None):
    """NQS.gradients for a real-parameter RBM: flat_gradient (jVMC/vqs.py:46-51),
    dRe + i dIm over the real parameters, length = #parameters."""
    s2 = np.asarray(s).reshape(np.asarray(s).shape[0], -1)
    sig = 2.0 * s2 - 1.0
    t = np.tanh(theta(s2, np.asarray(W, np.float64), b))
    gW = (sig[:, :, None] * t[:, None, :]).reshape(s2.shape[0], -1)
    parts = ([t] if b is not None else []) + [gW]
    return np.concatenate(parts, axis=1).astype(np.complex128)


def flatten_params(W, b=None):
    """NQS.get_parameters (jVMC/vqs.py:449-466): per leaf [Re ravel, Im ravel] for complex
    leaves, plain ravel for real ones; leaves in sorted-key order (bias, kernel)."""
    leaves = ([np.asarray(b)] if b is not None else []) + [np.asarray(W)]
    out = []
    for leaf in leaves:
        if np.iscomplexobj(leaf):
            out += [leaf.ravel().real, leaf.ravel().imag]
        else:
            out += [leaf.ravel()]
    return np.concatenate(out)


def unflatten_params(P, N, M, bias=False, real=False):
    """NQS._param_unflatten (jVMC/vqs.py:430-444).  ``P`` may be complex (MinSR update)."""
    P = np.asarray(P)
    start = 0
    b = None
    if bias:
        if real:
            b = P[start:start + M].copy()
            start += M
        else:
            b = P[start:start + M] + 1j * P[start + M:start + 2 * M]
            start += 2 * M
    n = N * M
    if real:
        W = P[start:start + n].reshape(N, M).copy()
    else:
        W = (P[start:start + n] + 1j * P[start + n:start + 2 * n]).reshape(N, M)
    return W, b


def init_cpx_rbm(N, M, bias=False, seed=1234):
    """Synthetic stand-in for jVMC/nets/initializers.py:17-20 (U[0,.01)+iU[0,.01), bias 0);
    jax's PRNGKey stream is unavailable, numpy default_rng is used instead."""
    rng = np.random.default_rng(seed)
    W = rng.uniform(0, 0.01, (N, M)) + 1j * rng.uniform(0, 0.01, (N, M))
    b = np.zeros(M, np.complex128) if bias else None
    return W, b


def init_o1(N, M, bias=False, seed=4321, real=False):
    """Second synthetic weight set of SURVEY 8d: Re W, Im W ~ U(-a, a), a = 1/sqrt(N)."""
    rng = np.random.default_rng(seed)
    a = 1.0 / np.sqrt(N)
    W = rng.uniform(-a, a, (N, M))
    if not real:
        W = W + 1j * rng.uniform(-a, a, (N, M))
    b = None
    if bias:
        b = rng.uniform(-a, a, M)
        if not real:
            b = b + 1j * rng.uniform(-a, a, M)
    return W, b
