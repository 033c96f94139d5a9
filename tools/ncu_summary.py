"""Developer tool: turn gpurun_out/*.ncu-rep / launch-list CSVs into the small committed files under profiles/.

    python tools/ncu_summary.py metrics gpurun_out/X.ncu-rep profiles/Y_metrics.csv     # DRAM / launch / stall columns of the raw page
    python tools/ncu_summary.py launches gpurun_out/L.csv                               # per-kernel totals and shares (markdown)
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

KEEP = re.compile(r"^(Kernel Name|gpu__time_duration|dram__bytes|dram__throughput|gpu__dram_throughput|dram__cycles_active\.avg|"
                  r"launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|cluster_size|waves_per_multiprocessor|occupancy_limit)|"
                  r"sm__throughput\.avg|sm__warps_active\.avg|sm__cycles_elapsed\.max|sm__inst_executed_pipe_fp64\.avg|sm__issue_active\.avg|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|lts__t_sector_hit_rate|smsp__average_warps_issue_stalled_.*_per_issue_active|"
                  r"smsp__inst_executed\.sum|sm__pipe_fp64_cycles_active\.avg)")


def metrics(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    cols = [i for i, h in enumerate(hdr) if KEEP.match(h)]
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        for r in rows:
            w.writerow([r[i] if i < len(r) else "" for i in cols])
    print("wrote", out, len(rows) - 2, "launch(es)")


def launches(path):
    tot, cnt = defaultdict(float), defaultdict(int)
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("fb200::", "")
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"| `{k}` | {cnt[k]} | {tot[k] / 1e3:.2f} | {tot[k] / cnt[k]:.1f} | {100 * tot[k] / total:.2f} % |")
    print(f"| total | {sum(cnt.values())} | {total / 1e3:.2f} | | |")


if __name__ == "__main__":
    if sys.argv[1] == "metrics":
        metrics(sys.argv[2], sys.argv[3])
    else:
        launches(sys.argv[2])
