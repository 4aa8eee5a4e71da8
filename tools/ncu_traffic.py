"""Extract per-launch DRAM traffic of a kernel from an `ncu --set full` report into profiles/ncu_traffic_r02.json (read by
bench.py for `roofline.traffic`).   python tools/ncu_traffic.py <report.ncu-rep> <key> <batch>"""
import csv
import io
import json
import os
import subprocess
import sys

rep, key, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]


def metric(name):
    i = hdr.index(name)
    v = float(vals[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]


rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
dur_i = hdr.index("gpu__time_duration.sum")
out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic_r02.json")
data = json.load(open(out_path)) if os.path.exists(out_path) else {}
data[key] = {"kernel": vals[hdr.index("Kernel Name")], "batch": batch, "dram_bytes_read": rd, "dram_bytes_write": wr,
             "dram_bytes_per_launch": rd + wr, "duration": f"{vals[dur_i]} {units[dur_i]} (under ncu, cold caches)",
             "report": os.path.basename(rep)}
json.dump(data, open(out_path, "w"), indent=1)
print(json.dumps(data[key], indent=1))
