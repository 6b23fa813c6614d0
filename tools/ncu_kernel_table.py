"""Per-kernel table from `ncu --set full` reports (tools/ncu_all.sh): one row per (kernel, grid, block) with the number of
profiled launches and the MEAN of duration, DRAM bytes read / written, DRAM throughput %, tensor-pipe %, occupancy, plus the
derived achieved DRAM GB/s against the measured copy bandwidth.  No GPU needed (calls `ncu -i ... --page raw --csv`).

    python tools/ncu_kernel_table.py gpurun_out/ncu_all_*.ncu-rep > profiles/r2_ncu_kernels.json
"""
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_summarize import measured_hbm_gbps, summarise  # noqa: E402

KEYS = ["duration_ns", "dram_read_bytes", "dram_write_bytes", "dram_throughput_pct", "tensor_pipe_pct_of_active", "tensor_pipe_pct_of_elapsed",
        "sm_throughput_pct", "warps_active_pct", "registers_per_thread", "dyn_smem_bytes", "l2_hit_pct"]


def short(name):
    name = re.sub(r"\(.*$", "", name.replace("(anonymous namespace)::", ""))
    return name.replace("vcla::", "").strip()


def main():
    peak, src = measured_hbm_gbps()
    doc = {"hbm_peak_gbps": peak, "hbm_peak_source": src,
           "note": "ncu --set full --clock-control none; each launch replayed cold-cache and serialised: use durations for SHARES and the byte "
                   "counters for traffic, never as benchmark numbers.  Rows: mean over the profiled launches of one (kernel, grid, block).",
           "reports": {}}
    for path in sys.argv[1:]:
        groups = {}
        for rec in summarise(path, peak):
            if "error" in rec:
                doc["reports"][os.path.basename(path)] = rec
                continue
            k = (short(rec["kernel"]), rec.get("grid"), rec.get("block"))
            groups.setdefault(k, []).append(rec)
        rows = []
        for (name, grid, block), recs in groups.items():
            row = {"kernel": name, "grid": grid, "block": block, "launches": len(recs)}
            for key in KEYS:
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
        rows.sort(key=lambda r: -r.get("duration_ns", 0) * r["launches"])
        tot = sum(r.get("duration_ns", 0) * r["launches"] for r in rows) or 1.0
        for r in rows:
            r["share_of_profiled_time"] = round(r.get("duration_ns", 0) * r["launches"] / tot, 4)
        doc["reports"].setdefault(os.path.basename(path), {})["kernels"] = rows
    json.dump(doc, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
