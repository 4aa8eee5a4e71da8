"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: one timed step, per kernel and
(optionally) per launch.   python tools/launch_summary.py gpurun_out/x/launches.csv [--each pattern,pattern] [--step k]"""
import collections
import csv
import sys

path = sys.argv[1]
each = []
step = -3
for i, a in enumerate(sys.argv):
    if a == "--each":
        each = sys.argv[i + 1].split(",")
    if a == "--step":
        step = int(sys.argv[i + 1])
rows = list(csv.reader(open(path, errors="ignore")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, body = rows[hi], rows[hi + 1:]
kn, mv, mu, gs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
L = []
for r in body:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    v = v / 1000 if r[mu] == "ns" else (v * 1000 if r[mu] == "ms" else v)
    L.append((r[kn], v, r[gs]))
# a pipeline pass ends with the surface decoder's query kernel; take the pass `step` from the end (default: the timed one,
# followed by the stage-marks pass and the e2e passes)
ends = [i for i, (k, v, g) in enumerate(L) if "decode_lattice_kernel" in k]
print(f"# {len(L)} launches, {sum(x[1] for x in L) / 1000:.1f} ms of kernel time, {len(ends)} pipeline passes")
a, b = ends[step - 1], ends[step]
seg = L[a + 1:b + 1]
agg = collections.OrderedDict()
for k, v, g in seg:
    k = k.split("(")[0][:70]
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in agg.values())
print(f"# one pass (lattice decode to lattice decode): {len(seg)} launches, {tot / 1000:.2f} ms")
print("kernel,launches,total_us,share")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"\"{k}\",{n},{v:.1f},{v / tot:.4f}")
if each:
    print("# per launch")
    for k, v, g in seg:
        if any(s in k for s in each):
            print(f"{v:9.1f} us grid={g:16s} {k[:80]}")
