"""Mirror of jVMC/util/output_manager.py: numerical output (observables, metadata, network checkpoints) and the
wall-clock phase timers that TDVP / MinSR drive through ``outp=``.

The reference writes HDF5 through h5py (:51-200).  h5py is optional here: when it can be imported the same HDF5
file is produced (same groups, resizable float64 datasets); otherwise the identical tree -- group and dataset
names exactly as the reference creates them, e.g. ``/<group>/observables/<key>/<name>``,
``/<group>/metadata/<key>``, ``/<group>/network_checkpoints/{times,checkpoints}`` -- is kept as a flat
name -> array mapping in a NumPy ``.npz`` archive next to the requested file name (``<dataFileName>.npz``).
``read_dataset()`` reads either.  Only rank 0 writes; checkpoints read back are broadcast to all ranks."""
import os
import time as _time

import numpy as np
import torch

from .. import mpi_wrapper as mpi

try:  # pragma: no cover - not part of this image
    import h5py
except Exception:  # noqa: BLE001
    h5py = None


def _join(*parts):
    """HDF5-style absolute path from group / dataset name fragments."""
    segs = []
    for p in parts:
        segs += [x for x in str(p).split("/") if x]
    return "/" + "/".join(segs)


class _NpzTree:
    """A tree of growing datasets held in memory and mirrored into one ``.npz`` archive (keys = HDF5 paths)."""

    def __init__(self, path, append):
        self.path = path
        self.data = {}
        self.groups = set()
        if append and os.path.exists(path):
            with np.load(path, allow_pickle=False) as z:
                for k in z.files:
                    if k == "__groups__":
                        self.groups = set(str(g) for g in z[k].tolist())
                    else:
                        self.data[k] = z[k]
        self.flush()

    def require_group(self, name):
        if name != "/":
            self.groups.add(name)

    def append_row(self, name, row):
        row = np.asarray(row, dtype=np.float64)
        old = self.data.get(name)
        if old is None:
            old = np.zeros((0,) + row.shape, dtype=np.float64)
        if old.shape[1:] != row.shape:
            raise ValueError("dataset %s holds rows of shape %s, got %s" % (name, old.shape[1:], row.shape))
        self.data[name] = np.concatenate([old, row[None]], axis=0)

    def create(self, name, value):
        if name in self.data:
            raise ValueError("dataset %s exists" % name)       # h5py refuses to re-create a dataset, too
        self.data[name] = np.array(value)

    def read(self, name):
        return self.data[name]

    def contains(self, name):
        return name in self.data or name in self.groups or any(k.startswith(name.rstrip("/") + "/") for k in self.data)

    def flush(self):
        tmp = self.path + ".tmp.npz"
        np.savez(tmp, __groups__=np.array(sorted(self.groups), dtype=str), **self.data)
        os.replace(tmp, self.path)


class _H5Tree:  # pragma: no cover - exercised only where h5py exists
    """The reference's on-disk format: resizable float64 datasets in an HDF5 file."""

    def __init__(self, path, append):
        self.path = path
        with h5py.File(path, "a" if append else "w"):
            pass

    def require_group(self, name):
        with h5py.File(self.path, "a") as f:
            f.require_group(name)

    def append_row(self, name, row):
        row = np.asarray(row, dtype=np.float64)
        with h5py.File(self.path, "a") as f:
            if name not in f:
                f.create_dataset(name, (0,) + row.shape, maxshape=(None,) + row.shape, dtype="f8", chunks=True)
            ds = f[name]
            ds.resize((len(ds) + 1,) + row.shape)
            ds[-1] = row

    def create(self, name, value):
        with h5py.File(self.path, "a") as f:
            f.create_dataset(name, data=np.array(value))

    def read(self, name):
        with h5py.File(self.path, "r") as f:
            return np.array(f[name])

    def contains(self, name):
        with h5py.File(self.path, "r") as f:
            return name in f

    def flush(self):
        pass


class OutputManager:
    """``OutputManager(dataFileName, group="/", append=False)`` as in the reference (:25-37).  ``dataFileName=None``
    gives a timer-only object (nothing is written)."""

    def __init__(self, dataFileName, group="/", append=False, backend=None):
        self.fn = dataFileName
        self.currentGroup = "/"
        self.timings = {}
        self._tree = None
        if dataFileName is not None and mpi.rank == 0:
            use_h5 = (h5py is not None) if backend is None else (backend == "h5")
            if use_h5 and h5py is None:
                raise RuntimeError("backend='h5' needs h5py")
            if use_h5:
                self.path = dataFileName
                self._tree = _H5Tree(self.path, append)
            else:
                self.path = dataFileName if dataFileName.endswith(".npz") else dataFileName + ".npz"
                self._tree = _NpzTree(self.path, append)
        self.set_group(group)

    # ------------------------------------------------------------------ numerical output
    def set_group(self, group):
        """reference :39-49"""
        if group != "/":
            self.currentGroup = _join(group)
        if self._tree is not None:
            self._tree.require_group(self.currentGroup)
            self._tree.flush()

    @staticmethod
    def to_array(x):
        """reference :244-249: scalars become shape-(1,) arrays; tensors are brought to the host as float64."""
        if isinstance(x, torch.Tensor):
            x = x.detach()
            if x.is_complex():
                x = x.real
            x = x.to("cpu", torch.float64).numpy()
        if not isinstance(x, np.ndarray):
            x = np.array([x])
        if x.ndim == 0:
            x = x.reshape(1)
        return np.asarray(np.real(x), dtype=np.float64)

    def _append_series(self, sub, time, items):
        """one more row in ``<group>/<sub>/times`` and in every (path, value) of ``items``"""
        if self._tree is None:
            return
        base = _join(self.currentGroup, sub)
        self._tree.require_group(base)
        self._tree.append_row(_join(base, "times"), np.float64(float(time)))
        for path, value in items:
            self._tree.append_row(_join(base, path), self.to_array(value))
        self._tree.flush()

    def write_observables(self, time, **kwargs):
        """``write_observables(t, energy={"mean": .., "variance": ..}, ...)`` -> datasets
        ``<group>/observables/<key>/<name>`` growing along axis 0, plus ``observables/times`` (reference :51-85)."""
        items = [(_join(key, name), value) for key, obsDict in kwargs.items() for name, value in obsDict.items()]
        self._append_series("observables", time, items)

    def write_metadata(self, time, **kwargs):
        """datasets ``<group>/metadata/<key>`` and ``metadata/times`` (reference :87-116)"""
        self._append_series("metadata", time, list(kwargs.items()))

    def write_network_checkpoint(self, time, weights):
        """flat parameter vector (NQS.get_parameters()) appended to ``<group>/network_checkpoints/checkpoints``
        (reference :118-147)"""
        self._append_series("network_checkpoints", time, [("checkpoints", weights)])

    def get_network_checkpoint(self, time=-1, idx=-1):
        """(time, weights) of the checkpoint closest to ``time`` (or number ``idx``), on every rank (reference :149-174)."""
        payload = None
        if mpi.rank == 0:
            base = _join(self.currentGroup, "network_checkpoints")
            times = np.asarray(self._tree.read(_join(base, "times")))
            if time >= 0:
                idx = int(np.argmin(np.abs(times - time)))
            payload = np.concatenate([[times[idx]], np.asarray(self._tree.read(_join(base, "checkpoints")))[idx]])
        payload = mpi.bcast_unknown_size(payload)
        payload = np.asarray(payload.cpu() if isinstance(payload, torch.Tensor) else payload, dtype=np.float64)
        return float(payload[0]), payload[1:]

    def write_error_data(self, name, data, mpiRank=0):
        """reference :176-186: one-off dataset under ``/error_data``"""
        self.write_dataset(name, data, groupname="error_data", mpiRank=mpiRank)

    def write_dataset(self, name, data, groupname="/", mpiRank=0):
        """reference :188-200: one-off dataset ``/<groupname>/<name>`` (written by rank ``mpiRank``; with the npz
        backend only rank 0 owns the archive)"""
        if mpi.rank != mpiRank or self._tree is None:
            return
        self._tree.require_group(_join(groupname))
        if isinstance(data, torch.Tensor):
            data = data.detach().cpu().numpy()
        self._tree.create(_join(groupname, name), data)
        self._tree.flush()

    def read_dataset(self, name):
        """absolute HDF5-style path -> array (rank 0)"""
        return np.asarray(self._tree.read(_join(name)))

    def has(self, name):
        return self._tree.contains(_join(name))

    # ------------------------------------------------------------------ timers (reference :202-242)
    def _record(self, name):
        return self.timings.setdefault(name, {"total": 0.0, "last_total": 0.0, "newest": 0.0, "count": 0, "init": 0.0})

    def start_timing(self, name):
        self._record(name)["init"] = _time.perf_counter()

    def add_timing(self, name, elapsed):
        rec = self._record(name)
        rec["total"] += elapsed
        rec["newest"] = elapsed
        rec["count"] += 1

    def stop_timing(self, name):
        now = _time.perf_counter()
        known = name in self.timings
        self.add_timing(name, now - self.timings[name]["init"] if known else 0.0)

    def print_timings(self, indent=""):
        """time accumulated per phase since the previous call"""
        lines = ["%sRecorded timings:" % indent]
        for name, rec in self.timings.items():
            lines.append("%s  * %s: %fs" % (indent, name, rec["total"] - rec["last_total"]))
            rec["last_total"] = rec["total"]
        self.print("\n".join(lines))

    def print(self, text):
        if mpi.rank == 0:
            print(text, flush=True)
