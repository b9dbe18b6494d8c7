import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
import vmc_jax_b200 as jVMC
from vmc_jax_b200 import kernels as K
import vmc_jax_b200.operator as op
shape=(10,10); N=100; M=400
dev=jVMC.global_defs.myDevice
psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=M, bias=False), seed=1234)
psi(torch.zeros((1, 1) + shape, dtype=torch.int32, device=dev))
W, b = bench.o1_weights(N, M, False)
psi.set_parameters(torch.as_tensor(bench.flat_params(W, b)).to(dev))
smp = jVMC.sampler.MCSampler(psi, shape, 4321, updateProposer=jVMC.sampler.propose_spin_flip, numChains=2368,
                             sweepSteps=N, thermalizationSweeps=25, numSamples=2**16)
for it in range(4):
    s, logPsi, p = smp.sample()
    flat = s.reshape(-1, N).contiguous()
    tau = psi._tau(flat)
    r = K.i8_tail_ratios(tau)
    top = torch.topk(r, 5).values.tolist()
    Zr = torch.view_as_real(tau).reshape(tau.shape[0], -1)
    rms = torch.sqrt((Zr*Zr).mean(0))
    out = {}
    for T in (8, 12, 16, 24, 32):
        flag = (Zr.abs() > T*rms).any(1)
        out[T] = int(flag.sum())
    print("iter", it, "top ratios", [round(x,1) for x in top], "pred err", K.i8_predicted_error(top[0], top[1], tau.shape[0]),
          "max|tau|", float(tau.abs().max()), "outlier rows by T", out)
