"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` log into (a) one line per launch and (b) the share of each
kernel in the last complete query step (from the last k_project_dmma launch of a query batch to the launch before the
next one), which is what bench.py's per-stage CUDA-event times are compared with."""
import collections
import csv
import sys


def main(path, out):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = []
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        seq.append((r[ki].split("(")[0].replace("void ", "").replace("dpf::", ""), v / 1000 if r[ui] == "ns" else v))
    lines = ["# launch,kernel,us"] + [f"{i},{n},{v:.2f}" for i, (n, v) in enumerate(seq)]
    # query steps start with the hash of the query batch: k_project_dmma followed (soon) by k_probe_count
    starts = [i for i, (n, _) in enumerate(seq) if n.startswith("k_project_dmma") and any(m.startswith("k_probe_count") for m, _ in seq[i:i + 6])]
    share = []
    if len(starts) >= 2:
        a, b = starts[-2], starts[-1]
        agg = collections.OrderedDict()
        for n, v in seq[a:b]:
            agg[n] = agg.get(n, 0.0) + v
        tot = sum(agg.values())
        share = [f"# one query step (launches {a}..{b - 1}): {tot:.1f} us in kernels"] + [f"# {n},{v:.1f} us,{100 * v / tot:.1f} %" for n, v in agg.items()]
    open(out, "w").write("\n".join(share + lines) + "\n")
    print("\n".join(share))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
