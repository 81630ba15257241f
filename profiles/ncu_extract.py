"""Condense one kernel of an `ncu --set full` report into the small JSON kept
under profiles/ (and read by bench.py for `roofline.traffic`):
    python profiles/ncu_extract.py gpurun_out/merge.ncu-rep "<capture command>" "<workload>" > profiles/rN_merge_tiles_ncu_full.json
Needs the `ncu` CLI (reads the report with `--page raw --csv`); first profiled launch only."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
    "launch__block_size",
]

if __name__ == "__main__":
    rep, capture, workload = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO("".join(l for l in raw.splitlines(True) if not l.startswith("==")))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    out = {"capture": capture, "workload": workload, "kernel": col.get("Kernel Name", ("", ""))[0][:80]}
    for k in KEEP:
        if k in col:
            v, u = col[k]
            out[k] = f"{v} {u}".strip()
    stalls = {}
    for h, (v, u) in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            x = float(v.replace(",", "") or 0)
            if x >= 0.2:
                stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(x, 2)
    out["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
    print(json.dumps(out, indent=1))
