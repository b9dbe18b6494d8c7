"""Builds variants of libjvmc_b200.so with other compile-time tunables into build/variants/ (development aid):
    python tools/build_variants.py name1:-DJVMC_EL_MINB=5,-DJVMC_EL_MAXWPC=4 name2:...
Run a tool against one with JVMC_B200_LIB=build/variants/libjvmc_<name>.so (vmc_jax_b200/_lib.py)."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import NVCC, FLAGS, CSRC  # noqa: E402

out = os.path.join(ROOT, "build", "variants")
os.makedirs(out, exist_ok=True)
srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    defs = [d for d in defs.split(",") if d]
    objdir = os.path.join(out, name)
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        base = os.path.join(ROOT, "build", "obj", os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        # only the files that read a tunable are recompiled; the others are linked from the main build
        if any(d.split("=")[0][2:] in open(s).read() for d in defs):
            procs.append(subprocess.Popen([NVCC] + FLAGS + defs + ["-I", os.path.join(ROOT, "include"), "-c", s, "-o", o]))
        else:
            objs[-1] = base
    for p in procs:
        if p.wait() != 0:
            raise SystemExit("nvcc failed for variant " + name)
    lib = os.path.join(out, "libjvmc_%s.so" % name)
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs +
                          ["-L/usr/local/cuda/lib64", "-lcusolver", "-lcublas", "-ldl", "-Xlinker", "-rpath=/usr/local/cuda/lib64"])
    print("built", lib)
