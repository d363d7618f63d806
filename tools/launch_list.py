"""Prints the kernels of the last query step from an `ncu --metrics gpu__time_duration.sum --csv` log."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr, data = rows[0], rows[1:]
ki, vi, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Metric Unit")
seq = []
for r in data:
    if len(r) > vi and r[mi] == "gpu__time_duration.sum":
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        seq.append((r[ki].split("(")[0].replace("void ", "").replace("dpf::", ""), v))
anchor = sys.argv[2] if len(sys.argv) > 2 else "k_project_dmma"
idx = [i for i, (k, v) in enumerate(seq) if k.startswith(anchor)]
tot = 0
for k, v in seq[idx[-1]:]:
    print(f"{v:9.1f} us  {k}")
    tot += v
print(f"{tot:9.1f} us  total, {len(seq) - idx[-1]} launches")
