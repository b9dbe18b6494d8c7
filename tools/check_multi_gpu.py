"""Multi-GPU consistency check (run under torchrun): the half-volume Hermitian all-reduce of A must equal the plain
all-reduce.   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi_gpu.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dist.init_process_group("nccl")
from vmc_jax_b200 import mpi_wrapper as mpi, kernels as K  # noqa: E402

for R, M in ((7, 24), (101, 40), (13, 100)):
    Pc = R * M
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    X = torch.randn((Pc, Pc), dtype=torch.float64, device="cuda", generator=g) + \
        1j * torch.randn((Pc, Pc), dtype=torch.float64, device="cuda", generator=g)
    A = (X + X.conj().T).contiguous()                      # exactly Hermitian, different on every rank
    ref = mpi._all_reduce_sum(A.clone())
    out = mpi.all_reduce_hermitian_blocks(A.clone(), M)
    err = float((out - ref).abs().max())
    herm = float((out - out.conj().T).abs().max())
    if rank == 0:
        print("R=%d M=%d world=%d: max |half-volume - plain| = %.3e, hermiticity %.3e" % (R, M, dist.get_world_size(), err, herm))
    assert err < 1e-12 * float(ref.abs().max()) and herm == 0.0
if rank == 0:
    print("multi-GPU check ok")
dist.destroy_process_group()
