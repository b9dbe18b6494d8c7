"""Times jvmc_rbm_gram_S alone at a given shape (default: config 2, N=100, M=400) -- development aid.
    python tools/gram_bench.py [--B 8192] [--N 100] [--M 400] [--reps 3] [--check]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmc_jax_b200 import kernels as K  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8192)
ap.add_argument("--N", type=int, default=100)
ap.add_argument("--M", type=int, default=400)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--tile", type=int, default=0)
ap.add_argument("--check", action="store_true")
ap.add_argument("--dbg", type=int, default=0)
ap.add_argument("--i8", action="store_true")
ap.add_argument("--max-stages", type=int, default=0, help="int8 Gram: 32-sample stages per launch (0 = default)")
ap.add_argument("--clocks", action="store_true", help="sample nvidia-smi SM clock / power during the timed repetitions")
a = ap.parse_args()
TILE_ARG = a.tile + 1000 * (0 if a.i8 else a.dbg)
from vmc_jax_b200 import _lib  # noqa: E402
if a.i8:
    _lib.load().jvmc_i8_set_debug(a.dbg | (a.max_stages << 12))
dev = "cuda:0"
rng = np.random.default_rng(0)
s = torch.as_tensor(rng.integers(0, 2, (a.B, a.N)).astype(np.int32)).to(dev)
tau = torch.as_tensor(rng.uniform(-1, 1, (a.B, a.M)) + 1j * rng.uniform(-1, 1, (a.B, a.M))).to(dev)
mu = K.rbm_moments(s, tau, torch.full((a.B,), 1.0 / a.B, dtype=torch.complex128, device=dev), False, 0)
sigT = K.pack_sigma(s, False)
Pc = a.N * a.M
A = torch.empty((Pc, Pc), dtype=torch.complex128, device=dev)
def run():
    if a.i8:
        K.rbm_gram_S_i8(tau, sigT, mu, 1.0 / a.B, 1.0, out=A)
    else:
        K.rbm_gram_S(tau, sigT, mu, 1.0 / a.B, 1.0, out=A, tile=TILE_ARG)


run()
torch.cuda.synchronize()
if a.i8:
    A8 = A.clone()
    K.rbm_gram_S(tau, sigT, mu, 1.0 / a.B, 1.0, out=A, tile=0)
    torch.cuda.synchronize()
    print("i8 vs fp64 DMMA: max abs diff %.3e, scale %.3e, hermitian %s" % (float((A8 - A).abs().max()), float(A.abs().max()),
                                                                     bool(torch.equal(A8, A8.conj().T))))
    del A8
clk = None
if a.clocks:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import ClockSampler
    clk = ClockSampler(0)
    clk.start()
ts = []
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = min(ts)
if a.reps > 4:
    print("per-rep ms:", " ".join("%.0f" % x for x in ts))
if clk is not None:
    print("clocks under load:", clk.stop())
TS = a.tile or (80 if (a.M % 80 == 0 or a.M == 40) else 64)
nT = (a.M + TS - 1) // TS
execf = 8.0 * a.B * (a.N * (a.N + 1) / 2) * (nT * (nT + 1) / 2) * TS * TS
print("gram B=%d N=%d M=%d tile=%d: %.2f ms  executed %.2f TF/s  algorithmic(4NsPc^2) %.2f TF/s  ideal(2NsPc^2) %.2f TF/s"
      % (a.B, a.N, a.M, TS, ms, execf / ms / 1e9, 4.0 * a.B * Pc ** 2 / ms / 1e9, 2.0 * a.B * Pc ** 2 / ms / 1e9))
if a.check:
    n = min(a.B, 512)
    sn, tn = s[:n].contiguous(), tau[:n].contiguous()
    O = ((2.0 * sn.to(torch.float64) - 1)[:, :, None] * tn[:, None, :]).reshape(n, -1)
    sig2 = K.pack_sigma(sn, False)
    mu2 = K.rbm_moments(sn, tn, torch.full((n,), 1.0 / n, dtype=torch.complex128, device=dev), False, 0)
    A2 = K.rbm_gram_S(tn, sig2, mu2, 1.0 / n, 1.0, tile=a.tile)
    D = (O - mu2.reshape(1, -1)) / np.sqrt(n)
    cols = torch.arange(0, Pc, max(1, Pc // 257), device=dev)
    ref = D.conj().T @ D[:, cols]
    print("check max abs err", float((A2[:, cols] - ref).abs().max()), "scale", float(ref.abs().max()))
