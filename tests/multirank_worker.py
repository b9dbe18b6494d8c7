"""Worker of tests/test_gpu_multirank.py: one VMC step (sample -> E_loc -> <E>, F, A -> TDVP update) on this rank's share
of the chains; every rank writes its configurations and the (global) statistics to <out>.rank<r>.npz.
Launched with torch.distributed.run (world 1 or 2); JVMC_DIST_BACKEND=gloo lets two ranks share one GPU."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vmc_jax_b200 as jVMC  # noqa: E402
import vmc_jax_b200.operator as op  # noqa: E402
from vmc_jax_b200 import mpi_wrapper as mpi  # noqa: E402
from vmc_jax_b200.stats import SampledObs, RBMGradientObs  # noqa: E402


def main(out, total_chains, total_samples, mu, wscale):
    world, rank = mpi.commSize, mpi.rank
    L, M = 12, 20
    psi = jVMC.vqs.NQS(jVMC.nets.CpxRBM(numHidden=M, bias=True), seed=1234)
    psi(torch.zeros((1, 1, L), dtype=torch.int32, device=jVMC.global_defs.myDevice))
    rng = np.random.default_rng(77)
    psi.set_parameters(torch.as_tensor(wscale * rng.standard_normal(psi.get_parameters().shape[0])))
    H = op.BranchFreeOperator()
    for l in range(L):
        H.add(op.scal_opstr(-1., (op.Sz(l), op.Sz((l + 1) % L))))
        H.add(op.scal_opstr(-0.9, (op.Sx(l),)))
    smp = jVMC.sampler.MCSampler(psi, (L,), 4321, updateProposer=jVMC.sampler.propose_spin_flip,
                                 numChains=total_chains // world, sweepSteps=L, thermalizationSweeps=5,
                                 numSamples=total_samples, mu=mu)
    s, logPsi, p = smp.sample()
    Eloc = H.get_O_loc(s, psi, logPsi, 0.0)
    E = SampledObs(Eloc, p)
    G = RBMGradientObs(psi, s, p)
    Emean, Evar = E.mean()[0], E.var()[0]
    F = G.covar(E)
    A = G.gram_A()
    td = jVMC.util.TDVP(smp, snrTol=2, pinvTol=1e-8, rhsPrefactor=1.j, makeReal='imag')
    upd, res, cut = td.solve(E, G)
    # the reference-layout solve on dense gradient data: one C-ABI call (jvmc_tdvp_solve) that reduces the SNR moments
    # over ranks through the C ABI's own NCCL communicator (Python path when the ranks share a GPU over gloo)
    td2 = jVMC.util.TDVP(smp, snrTol=2, pinvTol=1e-8, rhsPrefactor=1.j, makeReal='imag')
    upd2, res2, _ = td2.solve(E, SampledObs(psi.gradients(s), p))
    comm_err = 0.0
    comm = mpi.capi_comm()
    if comm is not None:
        import ctypes
        import torch.distributed as dist
        from vmc_jax_b200 import _lib
        x = torch.arange(1000, dtype=torch.float64, device="cuda") * (rank + 1)
        ref = x.clone()
        dist.all_reduce(ref)
        _lib.call("jvmc_comm_allreduce_sum_f64", comm, _lib.ptr(x), 1000)
        comm_err = max(comm_err, float((x - ref).abs().max()))
        send = torch.full((64,), rank + 1, dtype=torch.uint8, device="cuda")
        recv = torch.zeros(64 * world, dtype=torch.uint8, device="cuda")
        _lib.call("jvmc_comm_allgather_bytes", comm, _lib.ptr(send), _lib.ptr(recv), 64)
        want = torch.arange(1, world + 1, dtype=torch.uint8, device="cuda").repeat_interleave(64)
        comm_err = max(comm_err, float((recv.to(torch.float64) - want.to(torch.float64)).abs().max()))
        b = torch.full((16,), float(rank), dtype=torch.float64, device="cuda")
        _lib.call("jvmc_comm_bcast_bytes", comm, _lib.ptr(b), 128, world - 1)
        comm_err = max(comm_err, float((b - (world - 1)).abs().max()))
        sb = torch.arange(8 * world, dtype=torch.float64, device="cuda") + rank
        rb = torch.zeros(8, dtype=torch.float64, device="cuda")
        _lib.call("jvmc_comm_reduce_scatter_sum_f64", comm, _lib.ptr(sb), _lib.ptr(rb), 8)
        tot = sum(torch.arange(8 * world, dtype=torch.float64, device="cuda") + r for r in range(world))
        comm_err = max(comm_err, float((rb - tot[8 * rank:8 * rank + 8]).abs().max()))
    # MinSR on the same samples (reference jVMC/util/minsr.py:53-80): the tangent kernel is formed once over the ranks
    # (rank k computes the k-th range of tile pairs, SUM all-reduce), pinv replicated, -Obar^dagger x all-reduced
    Gh = RBMGradientObs(psi, s, p)
    Tk = Gh.tangent_kernel()
    minsr = jVMC.util.MinSR(smp, pinvTol=1e-8)
    upd_minsr = minsr.solve(E, Gh, holomorphic=True)
    acc = smp.acceptance_ratio()
    # the half-volume Hermitian all-reduce (default from 8 ranks on) against the plain one, on rank-dependent data
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    Pc = (L + 1) * M
    X = torch.randn((Pc, Pc), dtype=torch.float64, device="cuda", generator=g) + \
        1j * torch.randn((Pc, Pc), dtype=torch.float64, device="cuda", generator=g)
    Hm = (X + X.conj().T).contiguous()
    ref = mpi._all_reduce_sum(Hm.clone())
    half = mpi.all_reduce_hermitian_blocks(Hm.clone(), M)
    herm_err = float((half - ref).abs().max() / ref.abs().max())
    torch.cuda.synchronize()
    np.savez(out + ".rank%d.npz" % rank, configs=s.cpu().numpy(), logPsi=logPsi.cpu().numpy(), p=p.cpu().numpy(),
             Emean=np.array([complex(Emean.item())]), Evar=np.array([float(Evar.item())]), F=F.cpu().numpy(),
             A=A.cpu().numpy(), update=upd.cpu().numpy(), update2=upd2.cpu().numpy(), comm_err=np.array([comm_err]),
             capi_comm=np.array([comm is not None]), residual=np.array([float(res)]), acc=np.array([float(acc)]),
             nglob=np.array([smp.get_last_number_of_samples()]), herm_err=np.array([herm_err]),
             world=np.array([world]), T=Tk.cpu().numpy(), update_minsr=upd_minsr.cpu().numpy(),
             params=psi.get_parameters().cpu().numpy())
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    import faulthandler
    import signal
    faulthandler.register(signal.SIGALRM, all_threads=True, chain=False)
    faulthandler.dump_traceback_later(200, exit=True)        # a hung collective must not outlive the test (frees the GPU)
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), float(sys.argv[5]))
