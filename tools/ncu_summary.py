"""Summarise an ncu report (raw page) into a small JSON: per kernel duration, DRAM bytes, FP64 pipe, stalls.
Usage: python tools/ncu_summary.py report.ncu-rep out.json"""
import csv, json, subprocess, sys, io
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
def col(n): return h.index(n) if n in h else None
keys = {"duration_us": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
        "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
        "warp_inst": "smsp__inst_executed.sum", "regs": "launch__registers_per_thread", "grid": "launch__grid_size", "block": "launch__block_size",
        "dram_throughput_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed"}
stall = [(i, n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for i, n in enumerate(h)
         if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
def tobytes(v, u):
    v = float(v.replace(",", "")); return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
res = []
for r in rows[2:]:
    d = {"kernel": r[col("Kernel Name")].split("(")[0].split("::")[-1]}
    for k, n in keys.items():
        c = col(n)
        if c is None or r[c] == "": continue
        if k.startswith("dram_r") or k.startswith("dram_w"): d[k + "_bytes"] = tobytes(r[c], units[c])
        elif k == "duration_us": d[k] = float(r[c].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[c], 1)  # ncu picks the unit
        else: d[k] = float(r[c].replace(",", ""))
    # FP64 operations the kernel executed (thread-level SASS counts: add + mul + 2 x fma), from the per-cycle rates of --set full
    cyc = col("smsp__cycles_elapsed.avg")
    ops = [col("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % o) for o in ("dadd", "dmul", "dfma")]
    if cyc is not None and all(o is not None for o in ops) and r[cyc] != "":
        f = [float(r[o].replace(",", "") or 0) for o in ops]
        d["fp64_flop"] = (f[0] + f[1] + 2 * f[2]) * float(r[cyc].replace(",", ""))
    d["top_stalls"] = [[n, round(float(r[i] or 0), 2)] for i, n in sorted(stall, key=lambda x: -float(r[x[0]] or 0))[:4]]
    res.append(d)
json.dump(res, open(out, "w"), indent=1)
tot = sum(d.get("dram_read_bytes", 0) + d.get("dram_write_bytes", 0) for d in res)
print("kernels", len(res), "total dram bytes", tot, "total us", sum(d["duration_us"] for d in res))
