"""Per-stage clock64 trace of one CTA of the int8 Gram kernel (development aid; jvmc_i8_set_trace).
    python tools/gram_trace.py [--B 20000] [--tile 20] [--dbg 0,100,2] [--out gpurun_out/gram_trace.npz]
Prints per mode: kernel time, median stage period of the MMA issuer, and the median of each hand-off interval."""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmc_jax_b200 import kernels as K, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=20000)
ap.add_argument("--N", type=int, default=100)
ap.add_argument("--M", type=int, default=400)
ap.add_argument("--tile", type=int, default=20)
ap.add_argument("--dbg", default="0,100,2")
ap.add_argument("--out", default="gpurun_out/gram_trace.npz")
a = ap.parse_args()
lib = _lib.load()
dev = "cuda:0"
rng = np.random.default_rng(0)
s = torch.as_tensor(rng.integers(0, 2, (a.B, a.N)).astype(np.int32)).to(dev)
tau = torch.as_tensor(rng.uniform(-1, 1, (a.B, a.M)) + 1j * rng.uniform(-1, 1, (a.B, a.M))).to(dev)
mu = K.rbm_moments(s, tau, torch.full((a.B,), 1.0 / a.B, dtype=torch.complex128, device=dev), False, 0)
sigT = K.pack_sigma(s, False)
Pc = a.N * a.M
A = torch.empty((Pc, Pc), dtype=torch.complex128, device=dev)
print("tile list:", K._i8_tiles(a.M, torch.device(dev)).cpu().numpy()[:, :5].tolist())
res = {}
for dbg in [int(x) for x in a.dbg.split(",")]:
    lib.jvmc_i8_set_debug(dbg)
    lib.jvmc_i8_set_trace(None, 0)
    K.rbm_gram_S_i8(tau, sigT, mu, 1.0 / a.B, 1.0, out=A)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K.rbm_gram_S_i8(tau, sigT, mu, 1.0 / a.B, 1.0, out=A)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    buf = torch.zeros(16 * 832, dtype=torch.int64, device=dev)
    lib.jvmc_i8_set_trace(ctypes.c_void_p(buf.data_ptr()), a.tile)
    K.rbm_gram_S_i8(tau, sigT, mu, 1.0 / a.B, 1.0, out=A)
    torch.cuda.synchronize()
    lib.jvmc_i8_set_trace(None, 0)
    st = (a.B + 31) // 32
    full = buf.cpu().numpy().reshape(832, 16).astype(np.float64)
    t = full[:st]
    c = full[831]
    print("   CTA phases (clk): set-up %.0f | stage loop %.0f | last commit -> accumulators complete %.0f | epilogue %.0f | "
          "exit barriers %.0f | total %.0f" % (c[1] - c[0], c[2] - c[1], c[3] - c[2], c[4] - c[3], c[5] - c[4], c[5] - c[0]))
    res["dbg%d" % dbg] = t
    med = lambda x: float(np.median(x))
    g = np.arange(10, st - 2)
    per = np.diff(t[:, 3])
    print("dbg=%d: kernel %.1f ms; MMA stage period median %.0f clk (p10 %.0f, p90 %.0f)"
          % (dbg, ms, np.median(per), np.percentile(per, 10), np.percentile(per, 90)))
    print("   MMA thread: stage start -> 6 MMAs issued %.0f | -> next stage's barriers seen %.0f | -> last MMA + commits %.0f | "
          "-> next stage start %.0f" % (med(t[g, 1] - t[g, 2]), med(t[g, 8] - t[g, 1]), med(t[g, 3] - t[g, 8]),
                                         med(t[g + 1, 2] - t[g, 3])))
    if t[:, 6].any():
        print("   sign warp: saw full -> signs prepared %.0f | -> saw TMEM buffer free %.0f | -> stored %.0f ; period %.0f"
              % (med(t[g, 5] - t[g, 4]), med(t[g, 6] - t[g, 5]), med(t[g, 7] - t[g, 6]), med(np.diff(t[:, 7]))))
        print("   signs of stage g stored -> MMA thread starts stage g: %.0f ; MMA stage g-2 committed -> sign warp sees buffer free: %.0f"
              % (med(t[g, 2] - t[g, 7]), med(t[g, 6] - t[g - 2, 3])))
    if t[:, 0].any():
        print("   TMA issued(g) -> MMA stage start(g): %.0f" % med(t[g, 2] - t[g, 0]))
lib.jvmc_i8_set_debug(0)
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
np.savez(a.out, **res)
