"""Per-kernel table from the CSV logs of tools/ncu_metrics.sh (`ncu --metrics ... --csv --log-file`): one row per (kernel, grid, block)
with the number of profiled launches and the MEAN of duration, DRAM bytes read / written, DRAM throughput %, tensor-pipe %, occupancy,
registers, plus the achieved DRAM GB/s against the measured copy bandwidth.  ncu replays every launch cold-cache and serialised:
quote SHARES and TRAFFIC from here, never a benchmark number.  No GPU needed.

    python tools/ncu_csv_table.py gpurun_out/ncu_metrics_*.csv > profiles/r2_ncu_kernels.json
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHORT = {"gpu__time_duration.sum": "duration_ns", "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
         "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_of_active",
         "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_of_elapsed",
         "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
         "launch__registers_per_thread": "registers_per_thread", "launch__shared_mem_per_block_dynamic": "dyn_smem_bytes",
         "lts__t_sector_hit_rate.pct": "l2_hit_pct"}
SCALE = {"nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("vcla::", "")
    name = re.sub(r"\((?:int|bool)\)", "", name)
    return re.sub(r"\(.*$", "", name).replace("void ", "").strip()


def load(path):
    launches = {}
    with open(path, newline="") as fh:
        rows = [r for r in csv.reader(fh) if len(r) >= 15]
    head = next(i for i, r in enumerate(rows) if r[0] == "ID")
    for r in rows[head + 1:]:
        key = int(r[0])
        rec = launches.setdefault(key, {"kernel": short(r[4]), "block": r[7], "grid": r[8]})
        k = SHORT.get(r[12])
        if k is None:
            continue
        try:
            v = float(r[14].replace(",", ""))
        except ValueError:
            continue
        rec[k] = v * SCALE.get(r[13], 1.0) if k in ("duration_ns", "dram_read_bytes", "dram_write_bytes", "dyn_smem_bytes") else v
    return [launches[k] for k in sorted(launches)]


def main():
    try:
        peak, src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        peak, src = 6584.8, "MEASURED_PEAKS.json (round 1 value)"
    doc = {"hbm_peak_gbps": peak, "hbm_peak_source": src,
           "note": "ncu --metrics <short list> --clock-control none over tools/profile_step.py (REAL widths, shallow stack: 2 ViT / 1 Resampler / 2 LLaMA "
                   "layers, eager launches).  Every launch is replayed cold-cache and serialised: durations are for SHARES, byte counters for TRAFFIC; "
                   "never benchmark numbers.  Rows = mean over the profiled launches of one (kernel, grid, block).",
           "reports": {}}
    for path in sys.argv[1:]:
        groups, order = {}, []
        for rec in load(path):
            k = (rec["kernel"], rec["grid"], rec["block"])
            if k not in groups:
                order.append(k)
            groups.setdefault(k, []).append(rec)
        rows = []
        for k in order:
            recs = groups[k]
            row = {"kernel": k[0], "grid": k[1], "block": k[2], "launches": len(recs)}
            for key in SHORT.values():
                vals = [r[key] for r in recs if key in r]
                if vals:
                    row[key] = round(sum(vals) / len(vals), 3)
            if "duration_ns" in row and "dram_read_bytes" in row:
                tr = row["dram_read_bytes"] + row.get("dram_write_bytes", 0.0)
                row["duration_us"] = round(row["duration_ns"] / 1e3, 2)
                row["dram_traffic_mb"] = round(tr / 1e6, 3)
                row["dram_gbps"] = round(tr / row["duration_ns"], 1)
                row["dram_frac_of_measured_copy_bw"] = round(tr / row["duration_ns"] / peak, 4)
            rows.append(row)
        tot = sum(r.get("duration_ns", 0) * r["launches"] for r in rows) or 1.0
        for r in rows:
            r["share_of_profiled_time"] = round(r.get("duration_ns", 0) * r["launches"] / tot, 4)
        rows.sort(key=lambda r: -r["share_of_profiled_time"])
        doc["reports"][os.path.basename(path)] = {"total_profiled_us": round(tot / 1e3, 1), "kernels": rows}
    json.dump(doc, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
