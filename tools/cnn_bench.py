"""Times the CNN sampler and local-energy kernels at BASELINE configs[3] (12x12 Heisenberg, CNN F=(3,3), channels (6,4)),
incremental (csrc/cnn_inc.cu) against generic (csrc/cnn.cu) -- development aid.
    python tools/cnn_bench.py [--L 12] [--chains 1184] [--sweeps 7] [--therm 10]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vmc_jax_b200 as jVMC  # noqa: E402
import vmc_jax_b200.operator as op  # noqa: E402
from vmc_jax_b200 import kernels as K, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=12)
ap.add_argument("--chains", type=int, default=1184)
ap.add_argument("--sweeps", type=int, default=7)
ap.add_argument("--therm", type=int, default=10)
ap.add_argument("--no-generic", action="store_true")
a = ap.parse_args()
L = a.L
N = L * L
H = op.BranchFreeOperator()
for x in range(L):
    for y in range(L):
        i = x * L + y
        for j in (x * L + (y + 1) % L, ((x + 1) % L) * L + y):
            H.add(op.scal_opstr(-0.25, (op.Sx(i), op.Sx(j))))
            H.add(op.scal_opstr(-0.25, (op.Sy(i), op.Sy(j))))
            H.add(op.scal_opstr(0.25, (op.Sz(i), op.Sz(j))))
psi = jVMC.vqs.NQS(jVMC.nets.CNN(F=(3, 3), channels=(6, 4), strides=(1, 1), bias=True, firstLayerBias=False), seed=7)
dev = jVMC.global_defs.myDevice
psi(torch.zeros((1, 1, L, L), dtype=torch.int32, device=dev))
theta = psi.get_parameters()
cd = psi._cnnDesc
neel = torch.as_tensor((np.indices((L, L)).sum(0) % 2).reshape(-1).astype(np.int32)).to(dev)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best, out = 1e30, None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return out, best


def sample():
    states = neel[None, :].repeat(a.chains, 1).contiguous()
    counters = torch.zeros(2, dtype=torch.int64, device=dev)
    cfg = K.cnn_mcmc(states, theta, cd, 99, 0, 0, "spin_flip_zeroMag", 2.0, N, a.therm * N, a.sweeps, counters)
    return cfg, counters


tab = H._ensure_compiled(0)
dt, pref = tab.device_tables(), tab.eval_prefactors()
for generic in ([0] if a.no_generic else [0, 1]):
    _lib.load().jvmc_cnn_set_generic(generic)
    (cfg, counters), t_s = timed(sample)
    steps = a.chains * (a.therm + a.sweeps) * N
    acc = float(counters[1]) / float(counters[0])
    name = {0: "incremental", 1: "generic (full forward per proposal)", 2: "incremental, wide variant"}[generic]
    print("%s: sampler %.2f ms for %d samples, %.3g proposals (acceptance %.3f): %.2f us per proposal and chain, %.1f M proposals/s"
          % (name, t_s, cfg.shape[0], steps, acc, t_s * 1e3 / ((a.therm + a.sweeps) * N), steps / t_s / 1e3))
    s = cfg.reshape(1, cfg.shape[0], L, L)
    lp = psi(s)
    E, t_e = timed(lambda: H.get_O_loc(s, psi, lp))
    print("%s: E_loc %.2f ms for %d samples x %d strings (%.2f us per sample), <E>/N = %.6f"
          % (name, t_e, cfg.shape[0], tab.numOps, t_e * 1e3 / cfg.shape[0], float(E.real.mean()) / N))
_lib.load().jvmc_cnn_set_generic(0)
