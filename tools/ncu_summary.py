#!/usr/bin/env python
"""Summarise an `ncu --set full` report: one CSV row per captured launch + the per-kernel DRAM traffic
JSON that bench.py reports as `roofline.traffic`.

    python tools/ncu_summary.py gpurun_out/r02a_full.ncu-rep profiles/r02a_ncu_full_summary.csv [profiles/r02_dram_traffic.json arxiv]

The traffic JSON is keyed by workload ({"arxiv": {kernel: bytes}, "mag": {...}}): the capture of one workload only
updates its own entry.
"""
import csv
import io
import json
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
           "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum"]
# kernel-name fragment -> key used by bench.py's byte model
KEYS = {"k_aggregate_rows": "k_aggregate_fwd", "k_scatter_cols": "k_scatter_bwd", "k_aggregate_fast": "k_aggregate_fwd", "k_aggregate<": "k_aggregate_fwd", "k_combine_bwd": "k_combine_bwd",
        "k_route_minmax": "k_route_minmax", "k_scatter_bwd": "k_scatter_bwd", "k_scatter_slab": "k_scatter_bwd",
        "k_project_tc": "k_project_tc", "k_wgrad_tc": "k_wgrad_tc"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, out_csv = sys.argv[1], sys.argv[2]
    out_json = sys.argv[3] if len(sys.argv) > 3 else None
    workload = sys.argv[4] if len(sys.argv) > 4 else "arxiv"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(m) for m in METRICS if m in hdr]
    ki = hdr.index("Kernel Name")
    traffic = {}
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + [f"{hdr[c]} [{units[c]}]" for c in cols])
        for r in rows[2:]:
            w.writerow([r[ki]] + [r[c] for c in cols])
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            b = float(r[rd]) * SCALE.get(units[rd], 1.0) + float(r[wr]) * SCALE.get(units[wr], 1.0)
            for frag, key in KEYS.items():
                if frag in r[ki]:
                    traffic.setdefault(key, []).append(b)
    if out_json:
        try:
            table = json.load(open(out_json))
        except Exception:
            table = {}
        table[workload] = {k: sum(v) / len(v) for k, v in traffic.items()}
        json.dump(table, open(out_json, "w"), indent=1)
    print(open(out_csv).read())


if __name__ == "__main__":
    main()
