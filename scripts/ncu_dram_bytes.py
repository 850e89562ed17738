#!/usr/bin/env python
"""Extract dram__bytes_read.sum + dram__bytes_write.sum per launch of every kernel of an `ncu --set full` report and merge
them into profiles/ncu_dram_bytes.json (read by bench.py for `roofline.traffic`).
Usage: python scripts/ncu_dram_bytes.py gpurun_out/prof.ncu-rep f16_tc:8192:5 profiles/<summary file the numbers come from>"""
import csv
import re
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, key, source = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U = rows[0], rows[1]
ki, ri, wi = H.index("Kernel Name"), H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum")


def to_bytes(x, unit):
    return float(x.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


per = {}
for r in rows[2:]:
    m = re.search(r"(\w+_kernel)", r[ki])
    name = m.group(1) if m else r[ki]
    per.setdefault(name, []).append(to_bytes(r[ri], U[ri]) + to_bytes(r[wi], U[wi]))
out = {k: sum(v) / len(v) for k, v in per.items()}          # mean over the captured launches of each kernel
out["source"] = source
path = os.path.join(ROOT, "profiles", "ncu_dram_bytes.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d[key] = out
json.dump(d, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
