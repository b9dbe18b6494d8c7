"""Mirror of jVMC/mpi_wrapper.py: sample distribution and global reductions.

The reference stages every reduction through the host (pmap psum -> np.array -> MPI.Allreduce ->
device_put, reference :114-147) and gathers with pickle (:278-292).  Here one process drives one
GPU and the reductions run in-stream on device buffers through torch.distributed (NCCL over
NVLink / NVSwitch; gloo for the CPU-side tests)."""
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import global_defs

rank = 0
commSize = 1
globNumSamples = 0
myNumSamples = 0
communicationTime = 0.


def _refresh():
    global rank, commSize
    if dist.is_available() and dist.is_initialized():
        rank, commSize = dist.get_rank(), dist.get_world_size()
    else:
        rank, commSize = 0, 1


def init_distributed(backend=None):
    """Join the torchrun rendezvous (RANK / WORLD_SIZE / MASTER_ADDR from the environment)."""
    if dist.is_available() and not dist.is_initialized() and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        if backend is None:
            # JVMC_DIST_BACKEND=gloo lets several ranks share one GPU (tests on a single-GPU box); NCCL otherwise
            backend = os.environ.get("JVMC_DIST_BACKEND") or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = global_defs.myDevice
        dist.init_process_group(backend=backend, **kw)
    _refresh()


init_distributed()


_capi_comm = None


def capi_comm():
    """NCCL communicator of the C ABI (include/jvmc_b200.h: jvmc_comm_init) spanning the same ranks as the
    torch.distributed group, created on first use -- the 128-byte NCCL id of rank 0 travels over the existing group.
    It is handed to the C entry points that reduce in-stream themselves (jvmc_tdvp_solve).  None for a single rank or
    when the ranks do not own one GPU each (gloo on a shared device)."""
    global _capi_comm
    _refresh()
    if commSize == 1 or dist.get_backend() != "nccl":
        return None
    if _capi_comm is None:
        import ctypes
        from . import _lib
        lib = _lib.load()
        idbuf = ctypes.create_string_buffer(128)
        if rank == 0:
            _lib.check(lib.jvmc_comm_unique_id(idbuf), "jvmc_comm_unique_id")
        box = [bytes(idbuf.raw) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        idbuf = ctypes.create_string_buffer(box[0], 128)
        handle = ctypes.c_void_p()
        torch.cuda.set_device(global_defs.myDevice)
        _lib.check(lib.jvmc_comm_init(idbuf, rank, commSize, ctypes.byref(handle)), "jvmc_comm_init")
        _capi_comm = handle
    return _capi_comm


def distribute_sampling(numSamples, localDevices=None, numChainsPerDevice=1):
    """reference :53-95.  Samples per chain (rounded up) and the global number actually generated."""
    global globNumSamples
    _refresh()
    samplesPerProcess = numSamples // commSize
    if rank < numSamples % commSize:
        samplesPerProcess += 1
    if localDevices is None:
        globNumSamples = numSamples
        return samplesPerProcess, globNumSamples
    numChainsPerProcess = localDevices * numChainsPerDevice

    def spc(spp):
        return int((spp + numChainsPerProcess - 1) // numChainsPerProcess)
    a = numSamples % commSize
    globNumSamples = (a * spc(1 + numSamples // commSize) + (commSize - a) * spc(numSamples // commSize)) \
        * numChainsPerProcess
    return spc(samplesPerProcess), globNumSamples


def first_sample_id():
    """reference :98-111."""
    _refresh()
    mySamples = globNumSamples // commSize
    firstSampleId = rank * mySamples
    if rank < globNumSamples % commSize:
        firstSampleId += rank
    else:
        firstSampleId += globNumSamples % commSize
    return firstSampleId


def all_reduce_hermitian_blocks(A, M):
    """SUM all-reduce of a fully populated Hermitian matrix A [Pc, Pc] with block structure M x M, moving half the
    bytes: the suffix [r0 M, Pc) of every block row (contiguous per matrix row) is packed, reduced and unpacked, the
    part below the block diagonal is rebuilt by conjugate transposition (kernels.hermitian_mirror_blocks)."""
    global communicationTime
    _refresh()
    if commSize == 1:
        return A
    from . import kernels as K
    t0 = time.perf_counter()
    packed = K.hermitian_pack_blocks(A, M)
    dist.all_reduce(torch.view_as_real(packed), op=dist.ReduceOp.SUM)
    K.hermitian_pack_blocks(A, M, packed, unpack=True)
    del packed
    K.hermitian_mirror_blocks(A, M)
    communicationTime += time.perf_counter() - t0
    return A


def _all_reduce_sum(x):
    """In-place SUM all-reduce of a device tensor (complex via its real view)."""
    global communicationTime
    _refresh()
    if commSize == 1:
        return x
    t0 = time.perf_counter()
    x = x.contiguous()
    if x.is_complex():
        dist.all_reduce(torch.view_as_real(x), op=dist.ReduceOp.SUM)
    else:
        dist.all_reduce(x, op=dist.ReduceOp.SUM)
    communicationTime += time.perf_counter() - t0
    return x


def _as_tensor(data):
    if isinstance(data, torch.Tensor):
        return data
    return torch.as_tensor(np.asarray(data)).to(global_defs.myDevice)


def global_sum(data):
    """reference :114-147: sum over (device, batch) and over all ranks; shape data.shape[2:]."""
    data = _as_tensor(data)
    return _all_reduce_sum(data.sum(dim=(0, 1)))


def global_mean(data, p):
    """reference :150-179: sum_n p_n X_n."""
    data, p = _as_tensor(data), _as_tensor(p)
    w = p.reshape(p.shape + (1,) * (data.dim() - 2)).to(data.dtype if data.is_complex() else p.dtype)
    return _all_reduce_sum((w * data).sum(dim=(0, 1)))


def global_variance(data, p):
    """reference :182-221: sum_n p_n |X_n - <X>|^2 (summed over all devices)."""
    data, p = _as_tensor(data), _as_tensor(p)
    mean = global_mean(data, p)
    d = data - mean
    w = p.reshape(p.shape + (1,) * (data.dim() - 2))
    return _all_reduce_sum((w * (d.conj() * d).real).sum(dim=(0, 1)))


def global_covariance(data, p):
    """reference :224-243 (+ _cov_helper :20-24): sum_n p_n conj(X_n) X_n^T, shape [D, D]."""
    data, p = _as_tensor(data), _as_tensor(p)
    x = data.reshape(-1, data.shape[-1])
    w = p.reshape(-1, 1).to(x.dtype)
    return _all_reduce_sum(x.conj().T @ (w * x))


def bcast_unknown_size(data, root=0):
    """reference :246-275: broadcast a 1-d float64 array of a size only the root knows."""
    _refresh()
    if rank == root:
        data = np.asarray(data)
        if data.dtype != np.float64:
            raise TypeError("Datatype has to be float64.")
    if commSize == 1:
        return np.array(data)
    box = [data if rank == root else None]
    dist.broadcast_object_list(box, src=root)
    return np.array(box[0])


def gather(data):
    """reference :278-292: concatenation over ranks of data flattened over (device, batch)."""
    data = _as_tensor(data)
    flat = data.reshape((-1,) + tuple(data.shape[2:])).contiguous()
    _refresh()
    if commSize == 1:
        return flat
    global communicationTime
    t0 = time.perf_counter()
    # gloo (ranks sharing one GPU in the tests) has no all_gather of device tensors: its ranks exchange host copies
    xdev = flat.device if dist.get_backend() == "nccl" else torch.device("cpu")
    sizes = [torch.zeros(1, dtype=torch.int64, device=xdev) for _ in range(commSize)]
    dist.all_gather(sizes, torch.tensor([flat.shape[0]], dtype=torch.int64, device=xdev))
    nmax = int(max(int(s.item()) for s in sizes))
    isC = flat.is_complex()
    buf = torch.view_as_real(flat) if isC else flat
    pad = torch.zeros((nmax,) + tuple(buf.shape[1:]), dtype=buf.dtype, device=xdev)
    pad[:buf.shape[0]] = buf
    outs = [torch.empty_like(pad) for _ in range(commSize)]
    dist.all_gather(outs, pad)
    parts = [o[:int(s.item())] for o, s in zip(outs, sizes)]
    res = torch.cat(parts, dim=0).to(flat.device)
    communicationTime += time.perf_counter() - t0
    return torch.view_as_complex(res.contiguous()) if isC else res


def gather_offset(nLocal):
    """Offset of this rank's rows inside the array returned by gather()."""
    _refresh()
    if commSize == 1:
        return 0
    dev = global_defs.myDevice if dist.get_backend() == "nccl" else torch.device("cpu")
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(commSize)]
    dist.all_gather(sizes, torch.tensor([nLocal], dtype=torch.int64, device=dev))
    return int(sum(int(s.item()) for s in sizes[:rank]))


def get_communication_time():
    global communicationTime
    t = communicationTime
    communicationTime = 0.
    return t
