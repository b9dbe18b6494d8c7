import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import vmc_jax_b200 as jVMC, vmc_jax_b200.operator as op
from vmc_jax_b200 import kernels as K
from vmc_jax_b200.stats import SampledObs, RBMGradientObs
L=20
psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=2*L, bias=False), seed=1234)
psi(torch.zeros((1,1,L),dtype=torch.int32,device='cuda'))
H = op.BranchFreeOperator()
for l in range(L):
    H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz((l+1)%L)))); H.add(op.scal_opstr(-0.7, (op.Sx(l),)))
smp = jVMC.sampler.MCSampler(psi,(L,),4321,updateProposer=jVMC.sampler.propose_spin_flip_Z2,numChains=500,sweepSteps=L,numSamples=4096,thermalizationSweeps=25)
s,lp,p = smp.sample(); E = SampledObs(H.get_O_loc(s,psi,lp),p)
def T(f,n=5):
    f(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): r=f()
    torch.cuda.synchronize(); return r,(time.perf_counter()-t0)/n*1e3
G = RBMGradientObs(psi,s,p)
def gram():
    G._A=None; return G.gram_A()
A,t = T(gram); print("gram_A %.2f ms"%t)
_,t = T(lambda: G.kr_covar_with(E)); print("kr_covar %.2f ms"%t)
_,t = T(lambda: K.eigh_inplace(A.T.contiguous())); print("eigh complex 800: %.2f ms"%t)
td = jVMC.util.TDVP(smp, rhsPrefactor=1., pinvTol=1e-8, diagonalShift=10, makeReal='real')
_,t = T(lambda: td.solve(E,G)); print("solve total (A cached) %.2f ms"%t)
os.environ["JVMC_GRAM_BACKEND"]="dmma"
import vmc_jax_b200.stats as st; st.GRAM_BACKEND="dmma"
A,t = T(gram); print("gram_A dmma %.2f ms"%t)
