"""Developer tool: one-screen summary of bench.py JSON lines (files given on the command line)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        txt = open(path).read().replace("NaN", "null")
        d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    except Exception as exc:
        print(path, "ERR", exc)
        continue
    print("==", path.split("/")[-1], d.get("impl", "ours"), "n_gpus", d.get("n_gpus"))
    print("  ", {k: (round(d[k], 3) if isinstance(d.get(k), float) else d.get(k)) for k in
                 ("value", "unit", "ms_per_step", "iters_per_sec_in_loop", "time_to_tol_ms", "units_per_step", "backtracks_per_step", "gpu_launches", "wall_ms_per_step")})
    r = d.get("roofline")
    if r:
        print("   roofline", {k: (round(r[k], 4) if isinstance(r.get(k), float) else r.get(k)) for k in
                              ("avg_launch_ms", "achieved", "frac", "frac_algorithmic", "traffic", "gemm_fp64_equiv_tflops", "full_width_gemm_ms") if k in r})
    if d.get("e2e"):
        e = d["e2e"]
        print("   e2e", round(e["value"], 3), "ms/step", e.get("ms_per_step") and round(e["ms_per_step"], 2), "pcie", e.get("pcie_ceiling") and
              {k: round(v, 2) for k, v in e["pcie_ceiling"].items() if isinstance(v, float)})
    c = d.get("cpu_baseline")
    if c:
        p = c.get("parity_full_size")
        print("   cpu", round(c["value"], 4), c.get("kind"), c.get("cores"), "| parity ok:", p and p.get("ok"),
              p and {k: v for k, v in p.items() if k in ("iterations", "backtracks", "solution_rel_err", "objective_history_rel_err")})
        if p and "columns" in p:
            print("   columns", [(q["column"], q["ok"], q.get("horizon")) for q in p["columns"]])
    if d.get("clocks"):
        print("   clocks", d["clocks"])
    if d.get("collective"):
        print("   collective", d["collective"])
