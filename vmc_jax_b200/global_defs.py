"""Mirror of jVMC/global_defs.py (reference :1-54): dtypes and the device this process drives.

One process per GPU (reference deployment mode (i), documentation/source/parallelism.rst:73-83):
the leading "device" axis of every array has extent 1 and the device is cuda:LOCAL_RANK."""
import collections.abc
import os

import numpy as np
import torch

tCpx = np.complex128
tReal = np.float64
tCpxTorch = torch.complex128
tRealTorch = torch.float64


def _pick_device():
    if torch.cuda.is_available():
        idx = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()
        torch.cuda.set_device(idx)
        return torch.device("cuda", idx)
    # host-logic only (tests of the distribution arithmetic under gloo); kernels refuse to run here
    return torch.device("cpu")


myDevice = _pick_device()
myPmapDevices = [myDevice]
myDeviceCount = 1


def get_iterable(x):
    if isinstance(x, collections.abc.Iterable):
        return x
    return (x,)


def set_pmap_devices(devices):
    """reference :38-46.  Only one device per process is supported (rank-per-GPU)."""
    global myPmapDevices, myDeviceCount, myDevice
    devices = list(get_iterable(devices))
    if len(devices) != 1:
        raise NotImplementedError("vmc_jax_b200 drives one GPU per process; launch one rank per GPU (torchrun)")
    d = devices[0]
    if not isinstance(d, torch.device):
        d = torch.device(d) if isinstance(d, str) else myDevice
    myPmapDevices = [d]
    myDevice = d
    myDeviceCount = 1


def device_count():
    return len(myPmapDevices)


def devices():
    return myPmapDevices


def pmap_devices_updated(pmapDevices):
    return pmapDevices is None or list(pmapDevices) != list(myPmapDevices)
