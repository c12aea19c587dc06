#!/usr/bin/env python
"""DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels captured by tools/gpu_profile.sh, in the
format bench.py reads from profiles/ncu_traffic.json.  Usage: tools/ncu_traffic.py <tag> kernel..."""
import csv, io, json, os, sys

UNITS = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tag, kernels = sys.argv[1], sys.argv[2:]
out = {"capture": f"ncu --set full --clock-control none of `python bench.py --steps 1 --warmup 1 --outer 10 --profile`, tag {tag} "
                  "(dram__bytes_read.sum + dram__bytes_write.sum of one launch)",
       "shape": {"M_cpg": 1000000, "N_samples": 256, "K_known": 6, "n_unknown": 2, "dtype": "f64", "weights_storage": "u16"}, "kernels": {}}
for k in kernels:
    raw = f"gpurun_out/{tag}_{k}_raw.csv"
    if not os.path.exists(raw):
        continue
    rows = list(csv.reader(io.StringIO(open(raw).read())))
    if len(rows) < 3:
        continue
    col = {h: (rows[1][i], rows[2][i]) for i, h in enumerate(rows[0])}
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        unit, val = col[m]
        tot += float(val.replace(",", "")) * UNITS.get(unit, 1)
    out["kernels"][k] = int(tot)
print(json.dumps(out, indent=1))
