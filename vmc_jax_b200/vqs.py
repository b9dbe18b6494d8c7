"""Mirror of jVMC/vqs.py: the NQS wrapper (evaluation, gradients, flat parameter vector).

Only (Cpx)RBM ansaetze are on the B200 hot path; other nets raise NotImplementedError (there is no
CPU / eager fallback)."""
import numpy as np
import torch

from . import global_defs
from . import kernels as K
from .nets.rbm import _RBMBase
from .nets.sym_wrapper import SymNet
from .nets.cnn import CNN
from .nets.two_nets_wrapper import TwoNets


def _to_dev(x, dtype=None):
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.asarray(x))
    x = x.to(global_defs.myDevice)
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    return x


class NQS:
    """Variational wave function psi_theta(s) = exp(r_theta(s)) (reference jVMC/vqs.py:85-491).

    Args mirror the reference (:115-120): ``net``, ``logarithmic``, ``batchSize``, ``seed``,
    ``orbit``, ``avgFun``.  ``batchSize`` is accepted for compatibility: the kernels tile the batch
    themselves and never materialise more than the outputs."""

    def __init__(self, net, logarithmic=True, batchSize=1000, seed=1234, orbit=None, avgFun=None):
        if isinstance(net, (tuple, list)):
            net = TwoNets(net)                     # reference :166-167
        if orbit is not None:
            # reference :170-173: NQS(net, orbit=...) wraps the net into SymNet itself
            net = SymNet(orbit=orbit, net=net) if avgFun is None else SymNet(orbit=orbit, net=net, avgFun=avgFun)
        if isinstance(net, SymNet) and isinstance(net.net, TwoNets):
            raise NotImplementedError("SymNet around TwoNets has no device kernel (use NQS((rbm1, rbm2)))")
        if not isinstance(net, (_RBMBase, SymNet, CNN, TwoNets)):
            raise NotImplementedError("only jVMC.nets.CpxRBM / RBM (optionally inside SymNet) and CNN have B200 kernels; "
                                      "got %r" % (net,))
        if not logarithmic:
            raise NotImplementedError("non-logarithmic networks are not supported")
        self.net = net
        self.logarithmic = logarithmic
        self.batchSize = batchSize
        self.seed = seed
        self.initialized = False
        self.parameters = None
        self.realParams = not net.cpx
        self.realNets = False
        self.holomorphic = False
        self._isGenerator = False
        self._version = 0
        self._cache = None
        self._tables = None
        self.sampleShape = None
        self.sym = net.orbit if isinstance(net, SymNet) else None     # LatticeSymmetry of an orbit-averaged RBM
        self._symTables = None
        self.kind = "cnn" if isinstance(net, CNN) else ("tworbm" if isinstance(net, TwoNets) else
                                                        ("symrbm" if self.sym is not None else "rbm"))
        self._cnnDesc = None

    @property
    def khatri_rao(self):
        """True when per-sample gradients factorise as sigma (x) tau (bare RBM): the fused E_loc / Gram kernels apply."""
        return self.kind == "rbm"

    def sym_tables(self):
        if self._symTables is None:
            if self.sym.perm.shape[1] != self.N:
                raise ValueError("orbit acts on %d sites, the configurations have %d" % (self.sym.perm.shape[1], self.N))
            self._symTables = K.SymTables(self.sym, global_defs.myDevice)
        return self._symTables

    # ------------------------------------------------------------------ initialisation
    def init_net(self, s):
        """reference :187-218 (lazy init on first use; holomorphy known analytically here)."""
        if self.initialized:
            return
        self.sampleShape = tuple(s.shape[2:])
        self.N = int(np.prod(self.sampleShape))
        if self.kind == "cnn":
            self._cnnDesc = K.CnnDesc(self.net, self.sampleShape)
        else:
            if not self.net.cpx and len(self.sampleShape) != 1:
                raise NotImplementedError("real RBM acts on the last axis only (rbm.py:88); use 1-d sampleShape")
            self.M = self.net.nets[0].numHidden if self.kind == "tworbm" else self.net.numHidden
        self.parameters = {"params": self.net.init(self.seed, self.sampleShape, global_defs.myDevice)}
        self.holomorphic = bool(self.net.cpx)
        leaves = self._leaves()
        self.paramShapes = [(int(p.numel()), tuple(p.shape)) for p in leaves]
        self.numParameters = int(sum(p.numel() for p in leaves))
        self.initialized = True

    def _leaf_keys(self, tree=None):
        """(module, leaf) pairs in Flax's flattening order (sorted keys at both levels: Dense_0/bias, Dense_0/kernel;
        Conv_0/bias, Conv_0/kernel, Conv_1/...)."""
        t = self.parameters["params"] if tree is None else tree
        return [(m, k) for m in sorted(t.keys()) for k in sorted(t[m].keys())]

    def _leaves(self, tree=None):
        t = self.parameters["params"] if tree is None else tree
        return [t[m][k] for m, k in self._leaf_keys(t)]

    def _dense(self, which=0):
        """parameter group of the (which-th) RBM: the amplitude network for the two-network ansatz"""
        return self.parameters["params"]["nets_%d/Dense_0" % which if self.kind == "tworbm" else "Dense_0"]

    @property
    def W(self):
        return self._dense()["kernel"]

    @property
    def b(self):
        return self._dense().get("bias", None)

    def _cW(self, which=0):
        d = self._dense(which)
        b = d.get("bias", None)
        return d["kernel"].to(torch.complex128), (None if b is None else b.to(torch.complex128))

    def _bump(self):
        self._version += 1
        self._cache = None
        self._tables = None

    # ------------------------------------------------------------------ evaluation
    def _flat_configs(self, s):
        s = _to_dev(s, torch.int32)
        lead = tuple(s.shape[:2])
        return s.reshape(lead[0] * lead[1], -1).contiguous(), lead

    def __call__(self, s):
        """log psi(s) for configurations s[dev, B, *shape] -> complex128[dev, B] (reference :223-251)."""
        s = _to_dev(s, torch.int32)
        self.init_net(s)
        flat, lead = self._flat_configs(s)
        if self.kind == "cnn":
            return K.cnn_logpsi(flat, self.get_parameters(), self._cnnDesc).reshape(lead)
        if self.kind == "tworbm":
            r, _ = K.rbm_logpsi(flat, *self._cW(0))
            phi, _ = K.rbm_logpsi(flat, *self._cW(1))
            return (r.real + 1j * phi.real).reshape(lead)      # reference two_nets_wrapper.py:19-21
        W, b = self._cW()
        if self.sym is not None:
            return K.symrbm_logpsi(flat, W, b, self.sym_tables()).reshape(lead)
        logpsi, tau = K.rbm_logpsi(flat, W, b)
        self._remember(flat, tau)
        return logpsi.reshape(lead)

    def _remember(self, flat, tau):
        # the cache keeps `flat` alive, so its storage cannot be recycled for other data while cached
        self._cache = (flat, flat._version, self._version, tau)

    def _tau(self, flat):
        """tanh(theta) for the given flat configurations (cached from the last evaluation)."""
        c = self._cache
        if c is not None and c[0].data_ptr() == flat.data_ptr() and c[0].shape == flat.shape \
                and c[1] == flat._version and c[2] == self._version:
            return c[3]
        W, b = self._cW()
        _, tau = K.rbm_logpsi(flat, W, b)
        self._remember(flat, tau)
        return tau

    def flip_tables(self):
        if self._tables is None:
            W, b = self._cW()
            self._tables = K.rbm_tables(W, b)
        return self._tables

    def gradients(self, s):
        """d log psi / d theta_k for every configuration: complex128[dev, B, P] in the reference's flat
        layout (holomorphic: per leaf [g, i g], reference :66-69; real parameters: :46-51)."""
        s = _to_dev(s, torch.int32)
        self.init_net(s)
        flat, lead = self._flat_configs(s)
        if self.kind == "cnn":
            g = K.cnn_grad(flat, self.get_parameters(), self._cnnDesc)
        elif self.kind == "tworbm":
            # real parameters: d Re log psi + i d Im log psi (reference :46-51) = [d r / d theta_r, i d phi / d theta_phi]
            gs = []
            for which in (0, 1):
                W, b = self._cW(which)
                _, tau = K.rbm_logpsi(flat, W, b)
                gs.append(K.rbm_grad(flat, tau, b is not None, 1))
            g = torch.cat([gs[0], 1j * gs[1]], dim=1)
        elif self.sym is not None:
            W, b = self._cW()
            _, wts = K.symrbm_logpsi(flat, W, b, self.sym_tables(), want_weights=True)
            g = K.symrbm_grad(flat, W, b, self.sym_tables(), wts, 0 if self.holomorphic else 1)
        else:
            g = K.rbm_grad(flat, self._tau(flat), self.b is not None, 0 if self.holomorphic else 1)
        return g.reshape(lead + (g.shape[-1],))

    def _leaf_slices(self):
        out, start = [], 0
        for (m, k), (size, _) in zip(self._leaf_keys(), self.paramShapes):
            n = 2 * size if self.holomorphic else size
            out.append((m, k, start, start + n))
            start += n
        return out

    def gradients_dict(self, s):
        """reference :292-314."""
        g = self.gradients(s)
        out = {}
        for m, k, lo, hi in self._leaf_slices():
            out.setdefault(m, {})[k] = g[..., lo:hi]
        return out

    def grad_dict_to_vec_map(self):
        """reference :319-334."""
        out = {}
        for m, k, lo, hi in self._leaf_slices():
            out.setdefault(m, {})[k] = torch.arange(lo, hi)
        return out

    def get_sampler_net(self):
        """reference :337-352: (function evaluating Re log psi, current parameters)."""
        def evalReal(p, x):
            tmp = self.parameters
            self.parameters = p
            self._bump()
            try:
                return self(x[None, None, ...] if x.dim() == len(self.sampleShape) else x).real
            finally:
                self.parameters = tmp
                self._bump()
        return evalReal, self.parameters

    def sample(self, numSamples, key, parameters=None):
        return None     # (Cpx)RBM is not a generator (reference :356-375)

    # ------------------------------------------------------------------ parameters
    def update_parameters(self, deltaP):
        """reference :384-404."""
        if not self.initialized:
            self.set_parameters(deltaP)
        new = self._param_unflatten(deltaP)
        cur = self.parameters["params"]
        self.parameters = {"params": {m: {k: cur[m][k] + new[m][k] for k in cur[m]} for m in cur}}
        self._bump()

    def set_parameters(self, P):
        """reference :409-425.  P: flat vector (real, or complex for MinSR updates) or a parameter dict."""
        if not self.initialized:
            raise RuntimeError("Error in NQS.set_parameters(): Network not initialized. Evaluate net on example "
                               "input for initialization.")
        if isinstance(P, dict):
            self.params = P
        else:
            self.params = self._param_unflatten(P)

    def _param_unflatten(self, P):
        """reference :430-444."""
        P = _to_dev(P)
        out, start = {}, 0
        for (m, name), (size, shape) in zip(self._leaf_keys(), self.paramShapes):
            if not self.realParams:
                leaf = (P[start:start + size] + 1j * P[start + size:start + 2 * size]).reshape(shape).to(torch.complex128)
                start += 2 * size
            else:
                leaf = P[start:start + size].reshape(shape)
                leaf = leaf.real.to(torch.float64) if leaf.is_complex() else leaf.to(torch.float64)
                start += size
            out.setdefault(m, {})[name] = leaf
        return out

    def get_parameters(self):
        """reference :449-466: per leaf [Re ravel, Im ravel] (complex) or ravel (real)."""
        if not self.initialized:
            return None
        leaves = self._leaves()
        if not self.realParams:
            return torch.cat([torch.cat([p.reshape(-1).real, p.reshape(-1).imag]) for p in leaves])
        return torch.cat([p.reshape(-1) for p in leaves])

    @property
    def is_generator(self):
        return self._isGenerator

    @property
    def params(self):
        if self.initialized:
            return self.parameters["params"]
        return None

    @params.setter
    def params(self, val):
        if "params" in val and len(val) == 1:
            val = val["params"]
        self.parameters = {"params": {m: {k: _to_dev(v) for k, v in val[m].items()} for m in val}}
        self._bump()
