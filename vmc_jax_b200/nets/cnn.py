"""Descriptor mirroring jVMC/nets/cnn.py:18-81 (class CNN, real parameters).  Evaluation, per-sample gradients and
Metropolis sampling run in csrc/cnn.cu; parameters live in a dict with Flax's names
{"Conv_l": {"bias": [C_l], "kernel": [F..., C_{l-1}, C_l]}} (sorted-key leaf order: Conv_0/bias, Conv_0/kernel, ...)."""
import numpy as np
import torch

from . import activation_functions as act_funs


class CNN:
    """Convolutional network with real parameters: F filter diameter per axis, channels per layer, strides per axis,
    actFun per layer (the last one is repeated), bias / firstLayerBias, periodicBoundary (reference :18-42)."""
    cpx = False

    def __init__(self, F=(8,), channels=(10,), strides=(1,), actFun=(act_funs.elu,), bias=True, firstLayerBias=False,
                 periodicBoundary=True):
        self.F = tuple(int(f) for f in F)
        self.channels = tuple(int(c) for c in channels)
        self.strides = tuple(int(s) for s in strides)
        self.actFun = tuple(act_funs.resolve(f) for f in actFun)
        self.bias = bool(bias)
        self.firstLayerBias = bool(firstLayerBias)
        if not periodicBoundary:
            raise NotImplementedError("CNN with open boundaries (zero padding) has no device kernel")
        if len(self.F) not in (1, 2) or len(self.strides) != len(self.F):
            raise NotImplementedError("CNN kernels cover 1-d and 2-d lattices with one stride per axis")
        if not 1 <= len(self.channels) <= 8:
            raise NotImplementedError("CNN kernels cover 1 to 8 layers")
        self.periodicBoundary = True

    def __repr__(self):
        return "CNN(F=%r, channels=%r, strides=%r, actFun=%r, bias=%r, firstLayerBias=%r)" % (
            self.F, self.channels, self.strides, tuple(a.__name__ for a in self.actFun), self.bias, self.firstLayerBias)

    def layer_bias(self, l):
        return self.firstLayerBias if l == 0 else self.bias

    def descriptor(self, sampleShape):
        """int list understood by the jvmc_cnn_* entry points."""
        if len(sampleShape) != len(self.F):
            raise ValueError("CNN(F=%r) acts on %d-d configurations, got shape %r" % (self.F, len(self.F), tuple(sampleShape)))
        Lx, Ly = (sampleShape[0], 1) if len(sampleShape) == 1 else sampleShape
        Fx, Fy = (self.F[0], 1) if len(self.F) == 1 else self.F
        sx, sy = (self.strides[0], 1) if len(self.strides) == 1 else self.strides
        acts = [a.kernel_id for a in self.actFun] + [self.actFun[-1].kernel_id] * (len(self.channels) - len(self.actFun))
        return [len(self.channels), int(Lx), int(Ly), Fx, Fy, sx, sy, int(self.firstLayerBias), int(self.bias)] + \
            list(self.channels) + acts[:len(self.channels)]

    def init(self, seed, sampleShape, device):
        """variance_scaling(1.0, "fan_avg", "uniform") kernels, zero biases (reference :45, Flax defaults); jax's PRNG
        stream is not reproduced: numpy default_rng(seed)."""
        rng = np.random.default_rng(int(seed))
        params, cin = {}, 1
        for l, c in enumerate(self.channels):
            rf = int(np.prod(self.F))
            lim = np.sqrt(3.0 / (0.5 * (rf * cin + rf * c)))
            leaf = {"kernel": torch.as_tensor(rng.uniform(-lim, lim, self.F + (cin, c))).to(device)}
            if self.layer_bias(l):
                leaf["bias"] = torch.zeros(c, dtype=torch.float64, device=device)
            params["Conv_%d" % l] = leaf
            cin = c
        return params
