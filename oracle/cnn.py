"""CPU restatement of the reference's real-parameter CNN ansatz (jVMC/nets/cnn.py:18-81).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Layers: x <- act(conv(wrap_pad(x))) with Flax nn.Conv semantics (cross-correlation, kernel [F..., Cin, Cout],
VALID padding on the wrap-padded input, strides), then sum over positions and channels divided by
sqrt(#positions * #channels) of the last layer.  Flat parameter order = Flax's sorted leaves: Conv_l/bias,
Conv_l/kernel.  Parity: the reference holds no golden value for CNN outputs; pinned by finite differences
(tests/vqs_test.py:133-164 does the same for real-parameter nets) and by translation invariance.
"""
import numpy as np


def elu(z):
    return np.where(z > 0, z, np.expm1(np.minimum(z, 0)))


def poly5(x):
    q = x ** 2
    return ((0.133333333 * q - 0.333333333) * q + 1.) * x


def poly6(x):
    q = x ** 2
    return ((0.022222222 * q - 0.083333333) * q + 0.5) * q


ACT = {"elu": elu, "relu": lambda z: np.maximum(z, 0), "tanh": np.tanh, "poly5": poly5, "poly6": poly6,
       "square": lambda z: z ** 2}


def unflatten(theta, shape, F, channels, bias=True, firstLayerBias=False):
    """flat real vector -> list of (bias or None, kernel[F..., Cin, Cout]) per layer."""
    theta = np.asarray(theta, dtype=np.float64)
    layers, off, cin = [], 0, 1
    for l, c in enumerate(channels):
        hb = firstLayerBias if l == 0 else bias
        b = None
        if hb:
            b = theta[off:off + c]
            off += c
        n = int(np.prod(F)) * cin * c
        k = theta[off:off + n].reshape(tuple(F) + (cin, c))
        off += n
        layers.append((b, k))
        cin = c
    assert off == theta.size
    return layers


def num_parameters(F, channels, bias=True, firstLayerBias=False):
    n, cin = 0, 1
    for l, c in enumerate(channels):
        n += (c if (firstLayerBias if l == 0 else bias) else 0) + int(np.prod(F)) * cin * c
        cin = c
    return n


def cnn_logpsi(s, theta, F, channels, strides=None, actFun=("elu",), bias=True, firstLayerBias=False):
    """s int[B, *shape] -> float64[B]."""
    s = np.asarray(s)
    shape = s.shape[1:]
    dim = len(shape)
    strides = tuple(strides) if strides is not None else (1,) * dim
    acts = list(actFun) + [actFun[-1]] * (len(channels) - len(actFun))
    layers = unflatten(theta, shape, F, channels, bias, firstLayerBias)
    x = (2.0 * s - 1.0)[..., None]                                   # feature axis
    for (b, k), a in zip(layers, acts):
        pad = [(0, 0)] + [(0, f - 1) for f in F] + [(0, 0)]
        xp = np.pad(x, pad, mode="wrap")
        osz = [(xp.shape[1 + d] - F[d]) // strides[d] + 1 for d in range(dim)]
        out = np.zeros((x.shape[0],) + tuple(osz) + (k.shape[-1],))
        for off in np.ndindex(*F):
            sl = tuple(slice(off[d], off[d] + strides[d] * (osz[d] - 1) + 1, strides[d]) for d in range(dim))
            out += np.tensordot(xp[(slice(None),) + sl], k[off], axes=([-1], [0]))
        if b is not None:
            out = out + b
        x = ACT[a](out)
    nrm = np.sqrt(np.prod(x.shape[1:]))
    return x.reshape(x.shape[0], -1).sum(1) / nrm
