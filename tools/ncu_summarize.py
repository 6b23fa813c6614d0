"""Summarise `ncu --set full` reports into one JSON document for profiles/ (runs where there is no GPU: it only calls
`ncu -i <report> --page raw --csv`).

    python tools/ncu_summarize.py gpurun_out/ncu_*.ncu-rep > profiles/rN_ncu_kernels.json

Per profiled launch: kernel name, grid/block, duration, DRAM bytes read/written, achieved DRAM GB/s and its fraction of
the measured copy bandwidth (MEASURED_PEAKS.json), tensor-pipe activity, achieved occupancy, registers -- the columns
/opt/skills/guides/B200_PROFILING.md names.  Per-launch times under ncu are cold-cache and serialised: quote shares and
traffic from here, never a benchmark number.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# output key -> regex over ncu's metric column names (first match wins)
WANTED = {
    "duration_ns": r"^gpu__time_duration\.sum$",
    "dram_read_bytes": r"^dram__bytes_read\.sum$",
    "dram_write_bytes": r"^dram__bytes_write\.sum$",
    "dram_throughput_pct": r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
    "tensor_pipe_pct_of_active": r"^sm__pipe_tensor.*cycles_active.*pct_of_peak_sustained_active$",
    "tensor_pipe_pct_of_elapsed": r"^sm__pipe_tensor.*cycles_active.*pct_of_peak_sustained_elapsed$",
    "sm_throughput_pct": r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    "warps_active_pct": r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    "registers_per_thread": r"^launch__registers_per_thread$",
    "dyn_smem_bytes": r"^launch__shared_mem_per_block_dynamic$",
    "sm_clock_hz": r"^sm__cycles_elapsed\.avg\.per_second$",
    "l2_hit_pct": r"^lts__t_sector_hit_rate\.pct$",
}
UNIT_SCALE = {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9, "nsecond": 1.0,
              "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
              "hz": 1.0, "Khz": 1e3, "Mhz": 1e6, "Ghz": 1e9, "cycle/second": 1.0, "cycle/nsecond": 1e9, "cycle/usecond": 1e6}


def measured_hbm_gbps():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6500.0, "fallback"


def number(text):
    text = (text or "").replace(",", "").strip()
    try:
        return float(text)
    except ValueError:
        return None


def parse_raw_csv(text):
    """`--page raw --csv`: a header row (contains "Kernel Name"), a units row (its "ID" cell is empty), one row per launch."""
    rows = list(csv.reader(io.StringIO(text)))
    head = next((i for i, r in enumerate(rows) if "Kernel Name" in r), None)
    if head is None:
        return []
    names = rows[head]
    units = rows[head + 1] if head + 1 < len(rows) and (not rows[head + 1] or rows[head + 1][0].strip() == "") else None
    data = rows[head + (2 if units else 1):]
    cols = {k: next((j for j, n in enumerate(names) if re.search(rx, n)), None) for k, rx in WANTED.items()}
    idx = {n: j for j, n in enumerate(names)}
    out = []
    for r in data:
        if len(r) < len(names) or not r[0].strip():
            continue
        rec = {"kernel": r[idx["Kernel Name"]], "grid": r[idx["Grid Size"]] if "Grid Size" in idx else None,
               "block": r[idx["Block Size"]] if "Block Size" in idx else None}
        for k, j in cols.items():
            if j is None:
                continue
            v = number(r[j])
            if v is None:
                continue
            u = units[j].strip() if units and j < len(units) else ""
            if k in ("duration_ns", "dram_read_bytes", "dram_write_bytes", "sm_clock_hz", "dyn_smem_bytes") and u in UNIT_SCALE:
                v *= UNIT_SCALE[u]
            rec[k] = v
            rec.setdefault("_columns", {})[k] = names[j] + (f" [{u}]" if u else "")
        out.append(rec)
    return out


def summarise(path, peak):
    ncu = os.environ.get("NCU", "ncu")
    p = subprocess.run([ncu, "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if p.returncode != 0:
        return [{"report": os.path.basename(path), "error": p.stderr.strip()[-300:]}]
    recs = parse_raw_csv(p.stdout)
    for rec in recs:
        rec["report"] = os.path.basename(path)
        d, rd, wr = rec.get("duration_ns"), rec.get("dram_read_bytes"), rec.get("dram_write_bytes")
        if d and rd is not None and wr is not None:
            rec["duration_us"] = round(d / 1e3, 2)
            rec["dram_traffic_mb"] = round((rd + wr) / 1e6, 3)
            rec["dram_gbps"] = round((rd + wr) / d, 1)                  # bytes per ns = GB/s
            rec["dram_frac_of_measured_copy_bw"] = round((rd + wr) / d / peak, 4)
    return recs or [{"report": os.path.basename(path), "error": "no kernel rows found in --page raw --csv"}]


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    peak, src = measured_hbm_gbps()
    doc = {"hbm_peak_gbps": peak, "hbm_peak_source": src,
           "note": "ncu replays each launch cold-cache and serialised: durations are for shares/traffic, not benchmark numbers",
           "launches": []}
    for path in sys.argv[1:]:
        doc["launches"] += summarise(path, peak)
    json.dump(doc, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
