"""Timer part of jVMC/util/output_manager.py (:202-242).  The HDF5 observable / checkpoint writer is
out of scope (h5py is not part of the hot path); TDVP/MinSR only use the timing interface."""
import time

from .. import mpi_wrapper as mpi


class OutputManager:
    def __init__(self, dataFileName=None, group="/", append=False):
        self.fn = dataFileName
        self.timings = {}

    def start_timing(self, name):
        if name not in self.timings:
            self.timings[name] = {"total": 0.0, "last_total": 0.0, "newest": 0.0, "count": 0, "init": 0.0}
        self.timings[name]["init"] = time.perf_counter()

    def stop_timing(self, name):
        toc = time.perf_counter()
        if name not in self.timings:
            self.timings[name] = {"total": 0.0, "last_total": 0.0, "newest": 0.0, "count": 0, "init": toc}
        elapsed = toc - self.timings[name]["init"]
        self.timings[name]["total"] += elapsed
        self.timings[name]["newest"] = elapsed
        self.timings[name]["count"] += 1

    def add_timing(self, name, elapsed):
        if name not in self.timings:
            self.timings[name] = {"total": 0.0, "last_total": 0.0, "newest": 0.0, "count": 0, "init": 0.0}
        self.timings[name]["total"] += elapsed
        self.timings[name]["newest"] = elapsed
        self.timings[name]["count"] += 1

    def print_timings(self, indent=""):
        self.print("%sRecorded timings:" % indent)
        for key, item in self.timings.items():
            self.print("%s  * %s: %fs" % (indent, key, item["total"] - item["last_total"]))
        for key in self.timings:
            self.timings[key]["last_total"] = self.timings[key]["total"]

    def print(self, text):
        if mpi.rank == 0:
            print(text, flush=True)

    def write_observables(self, time, **kwargs):
        raise NotImplementedError("HDF5 output is outside the B200 hot path (SURVEY 2a-20)")
