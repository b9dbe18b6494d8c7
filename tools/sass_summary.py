"""Per-kernel SASS evidence: counts of the mnemonics that prove the Blackwell-native paths (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UBLKCP/UTMALDG (TMA), DMMA (fp64 tensor), LDGSTS (cp.async), plus
registers / shared memory from the ELF.  Writes profiles/<tag>_sass_summary.txt.
    python tools/sass_summary.py [tag]"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
PAT = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "DMMA", "HMMA", "IMMA",
       "LDGSTS", "SYNCS", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "ATOM", "RED", "BAR.SYNC", "SHFL", "POPC"]
lines = ["# SASS evidence per kernel (cuobjdump -sass build/obj/*.o, sm_100a); counts of static instructions",
         "# columns: kernel | registers | " + " ".join(PAT), ""]
for obj in sorted(glob.glob(os.path.join(ROOT, "build", "obj", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    regs = {}
    cur = None
    for ln in res.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", ln)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)))
    counts = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1)
            for p in PAT:
                if op == p or op.startswith(p + ".") or (p == "BAR.SYNC" and op.startswith("BAR.SYNC")):
                    counts[cur][p] += 1
    lines.append("## " + os.path.basename(obj))
    for fn, c in counts.items():
        name = subprocess.run(["c++filt", fn], stdout=subprocess.PIPE, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = name.split("(")[0][:70]
        r = regs.get(fn, ("?", "?"))
        lines.append("%-72s regs %-4s smem %-6s %s" % (name, r[0], r[1], " ".join("%s=%d" % (p, c[p]) for p in PAT if c[p])))
    lines.append("")
out = os.path.join(ROOT, "profiles", "%s_sass_summary.txt" % tag)
with open(out, "w") as f:
    f.write("\n".join(lines) + "\n")
print("wrote", out)
