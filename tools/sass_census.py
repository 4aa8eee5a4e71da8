"""Opcode census of the shipped library: per kernel, how many tcgen05 / TMA / bulk-copy / warp-reduction instructions the SASS
holds (cuobjdump -sass; runs on the CPU container).   python tools/sass_census.py [lib] > profiles/sass_census_rNN.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "garmentnets_b200", "lib", "libgarmentnets_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "REDUX", "CREDUX", "REDG", "ATOMG",
         "MATCH", "VOTE", "SHFL", "F2FP", "FFMA2", "HMMA", "LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        kernels[cur]["_total"] += 1
        op = m.group(1)
        if op in WATCH:
            kernels[cur][op] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.basename(lib)}: {len(kernels)} kernels; columns = instruction counts in the SASS (static, not executed counts)")
print("# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, "
      "SYNCS = mbarrier ops")
for (name, c), dm in zip(kernels.items(), demangle):
    short = re.sub(r"\(.*", "", dm)[:70]
    cols = " ".join(f"{k}={c[k]}" for k in WATCH if c[k])
    print(f"{short:70s} n={c['_total']:6d}  {cols}")
