"""Summarise an .ncu-rep (read here, no GPU): headline metrics + stall reasons + hottest source lines.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_source_lines]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fp64.sum", "smsp__cycles_active.avg", "sm__cycles_active.avg"]
for v in vals:
    print("==", v[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:75s} {v[i]:>16s} {units[i]}")
    st = [(float(v[i].replace(",", "")), h) for i, h in enumerate(hdr)
          if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v[i]]
    for x, h in sorted(st, reverse=True)[:8]:
        print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {x:8.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(src))]
h = next(r for r in rows if r and r[0] == "Address")
body = [r for r in rows if len(r) == len(h) and r[0].startswith("0x")]
ci, cs, cx = h.index("Warp Stall Sampling (All Samples)"), h.index("Source"), h.index("Instructions Executed")
tot = sum(float(r[ci] or 0) for r in body)
totx = sum(float(r[cx] or 0) for r in body)
print(f"-- SASS: {len(body)} instructions, {totx:.3g} warp-instructions executed, {tot:.0f} stall samples")
# opcode histogram by executed count
import collections
ops = collections.Counter()
for r in body:
    op = r[cs].split()
    op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
    ops[op.split(".")[0]] += float(r[cx] or 0)
print("-- executed warp-instructions by opcode")
for op, n in ops.most_common(22):
    print(f"  {op:12s} {n / totx:6.3f}")
print("-- hottest SASS by stall samples")
for r in sorted(body, key=lambda r: -float(r[ci] or 0))[:nsrc]:
    print(f"  {float(r[ci]) / max(tot, 1):6.3f}  exec {float(r[cx] or 0):10.3g}  {r[cs].strip()[:110]}")
