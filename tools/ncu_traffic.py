"""profiles/r2_ncu_kernels.json from an `ncu --set full` capture of the bench workload: per kernel, per launch —
dram bytes, duration, tensor-pipe / L2 / issue percentages — with the command and the commit it was taken at.
bench.py reads the file for roofline.traffic (never a typed-in constant).

  python tools/ncu_traffic.py <capture.ncu-rep> "<workload_key>" "<command>" [out.json]
"""
import csv, json, subprocess, sys

KEYS = {"gpu__time_duration.sum": "time", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
        "lts__t_sector_hit_rate.pct": "l2_hit_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "launch__registers_per_thread": "registers", "smsp__inst_executed.sum": "warp_instructions",
        "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_bytes",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def main(rep, workload_key, command, out="profiles/r2_ncu_kernels.json"):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    kernels = {}
    for r in data:
        name = r[ki].split("(")[0].replace("void ", "").replace("dpf::", "").split("<")[0].strip()
        rec = {}
        for k, short in KEYS.items():
            if k in hdr:
                i = hdr.index(k)
                v = float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
                rec[short] = v
        rec["dram_bytes"] = rec.pop("dram_read", 0.0) + rec.pop("dram_write", 0.0)
        rec["time_ms"] = rec.pop("time", 0.0) * 1e3
        kernels.setdefault(name, rec)        # first launch of each kernel
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    json.dump({"workload_key": workload_key, "command": command, "commit": commit, "capture": rep, "kernels": kernels},
              open(out, "w"), indent=1)
    print(json.dumps(kernels, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:])
