#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel summary kept under profiles/
(metrics as rows, kernels as columns) and, with --traffic-json, the DRAM bytes of one forward SpMM
(sum over its kernels) that bench.py reports as roofline.traffic.

    ncu -i gpurun_out/spmm_full.ncu-rep --page raw --csv > gpurun_out/spmm_full_raw.csv
    python tools/ncu_summary.py gpurun_out/spmm_full_raw.csv profiles/rNN_spmm_ncu_full_summary.csv \
        --traffic-json profiles/spmm_traffic.json --algorithmic-bytes 27090519040
"""
import argparse
import csv
import json
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "lts__t_sectors_srcunit_tex_op_read.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
]
COMPUTE_METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "smsp__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
]
UNIT_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def read_raw(path):
    with open(path, newline="") as f:
        rows = [r for r in csv.reader(f) if r]
    # skip any "==PROF==" style preamble: the header row is the one holding "Kernel Name"
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units, data = rows[start], rows[start + 1], rows[start + 2:]
    return names, units, data


def to_float(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv")
    ap.add_argument("out_csv")
    ap.add_argument("--traffic-json", default=None)
    ap.add_argument("--algorithmic-bytes", type=int, default=None)
    ap.add_argument("--kernels", type=int, default=0, help="use only the first K kernel rows (0 = all)")
    ap.add_argument("--note", default="")
    ap.add_argument("--compute", action="store_true", help="pipe-utilisation / stall metrics (compute-bound kernels)")
    args = ap.parse_args()
    names, units, data = None, None, []
    for path in args.raw_csv.split(","):          # several captures with the same metric set: one column each
        n_, u_, d_ = read_raw(path)
        if names is None:
            names, units = n_, u_
        assert n_ == names, "captures were taken with different metric sets"
        data += d_
    if args.kernels:
        data = data[:args.kernels]
    col = {n: i for i, n in enumerate(names)}
    kn = col["Kernel Name"]
    with open(args.out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [r[kn][:44] for r in data])
        for key in ("Kernel Name", "Grid Size", "Block Size"):
            if key in col:
                w.writerow([key, ""] + [r[col[key]] for r in data])
        for m in (COMPUTE_METRICS if args.compute else METRICS):
            if m in col:
                w.writerow([m, units[col[m]]] + [r[col[m]] for r in data])
            else:
                print(f"note: metric {m} not in {args.raw_csv}", file=sys.stderr)
    if args.traffic_json:
        total = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = UNIT_BYTES[units[col[m]]]
            total += sum((to_float(r[col[m]]) or 0.0) * scale for r in data)
        dur_unit = units[col["gpu__time_duration.sum"]]
        dur_scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "s": 1e3}.get(dur_unit, 1.0)
        ms = sum((to_float(r[col["gpu__time_duration.sum"]]) or 0.0) * dur_scale for r in data)
        out = {"dram_bytes_per_launch": int(round(total)),
               "source": f"{args.out_csv}: dram__bytes_read.sum + dram__bytes_write.sum summed over the {len(data)} kernels of "
                         f"one forward SpMM, ncu --set full --clock-control none, C4 workload, default tuning. {args.note}".strip(),
               "kernels": [r[kn][:60] for r in data], "ncu_duration_ms_sum": ms}
        if args.algorithmic_bytes:
            out["algorithmic_bytes"] = args.algorithmic_bytes
        with open(args.traffic_json, "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))


if __name__ == "__main__":
    main()
