"""Oracle: SampledObs statistics (jVMC/stats.py:136-336) in NumPy, without the device axis.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

observations: [B] or [B, D]; weights: [B] summing to one over the whole sample.
"""
import numpy as np


class SampledObs:
    def __init__(self, observations=None, weights=None):
        self._weights = None if weights is None else np.asarray(weights, np.float64)
        self._data = None
        self._mean = None
        if observations is not None and weights is not None:
            obs = np.asarray(observations)
            if obs.ndim == 1:
                obs = obs[:, None]
            # stats.py:197-205: mean = sum_n w_n O_n ; data = sqrt(w_n) (O_n - mean)
            self._mean = np.tensordot(self._weights, obs, axes=(0, 0))
            self._data = np.sqrt(self._weights)[:, None] * (obs - self._mean[None, :])

    def mean(self):
        return self._mean

    def covar(self, other=None):
        """stats.py:235-245: conj(data1)^T data2."""
        other = self if other is None else other
        return np.conj(self._data).T @ other._data

    def var(self):
        """stats.py:248-252."""
        return np.sum(np.abs(self._data) ** 2, axis=0)

    def covar_data(self, other=None):
        """stats.py:255-265: per-sample outer(conj(d1), d2)/w as a new observable."""
        other = self if other is None else other
        obs = np.einsum('na,nb->nab', np.conj(self._data), other._data) / self._weights[:, None, None]
        out = SampledObs()
        out._weights = self._weights
        out._mean = np.tensordot(self._weights, obs, axes=(0, 0))
        out._data = np.sqrt(self._weights)[:, None, None] * (obs - out._mean[None])
        return out

    def covar_var(self, other=None):
        """stats.py:268-279."""
        other = self if other is None else other
        outer = np.einsum('na,nb->nab', np.conj(self._data), other._data)
        return np.sum(np.abs(outer) ** 2 / self._weights[:, None, None], axis=0) - np.abs(self.covar(other)) ** 2

    def transform(self, nonLinearFun=lambda x: x, linearFun=None):
        """stats.py:282-292: f(data/sqrt(w)+mean), optionally followed by V @ ."""
        sw = np.sqrt(self._weights).reshape((-1,) + (1,) * (self._data.ndim - 1))
        x = nonLinearFun(self._data / sw + self._mean[None])
        if linearFun is not None:
            x = np.matmul(linearFun, x)
        out = SampledObs()
        out._weights = self._weights
        out._mean = np.tensordot(self._weights, x, axes=(0, 0))
        out._data = sw * (x - out._mean[None])
        return out

    def subset(self, start=None, end=None, step=None):
        """stats.py:310-329."""
        sl = slice(start, end, step)
        out = SampledObs()
        w = self._weights[sl]
        nrm = np.sum(w)
        d = self._data[sl] / np.sqrt(nrm)
        w = w / nrm
        sw = np.sqrt(w).reshape((-1,) + (1,) * (d.ndim - 1))
        out._weights = w
        out._mean = np.tensordot(np.sqrt(w), d, axes=(0, 0)) + self._mean
        out._data = d + sw * (self._mean - out._mean)[None]
        return out

    def tangent_kernel(self):
        """stats.py:332-336."""
        return self._data @ np.conj(self._data).T
