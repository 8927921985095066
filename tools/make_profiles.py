#!/usr/bin/env python
"""Assemble profiles/r01_* from the files tools/gpu_profiles.sh left in gpurun_out/ (run here after the GPU call)."""
import json
import os
import shutil

G, P = "gpurun_out", "profiles"
JAC = ("k_jb_stage0", "k_jb_p01", "k_jb_p1a", "k_jb_p1b", "k_jb_p2", "k_jb_p3c", "k_scan", "k_fill", "k_sortrows")


def last_json(path):
    return [l for l in open(path) if l.startswith("{")][-1]


def one_iteration(rows):
    """Kernels of the LAST complete residual + Jacobian iteration of tools/one_jac.py (ends with k_sortrows)."""
    ends = [i for i, r in enumerate(rows) if r["kernel"].startswith("k_sortrows")]
    e = ends[-1]
    s = ends[-2] + 1 if len(ends) > 1 else 0
    return rows[s : e + 1]


for c in ("d3dHsm", "d3dHsm4x", "case1", "box2d"):
    open(os.path.join(P, "r01_bench_%s.json" % c), "w").write(last_json(os.path.join(G, "bench_%s.json" % c)))
    shutil.copy(os.path.join(G, "launches_%s.csv" % c), os.path.join(P, "r01_launches_batched_%s.csv" % c))
open(os.path.join(P, "r01_bench_reference_arm_d3dHsm.json"), "w").write(last_json(os.path.join(G, "ref_d3dHsm.json")))

traffic = {"note": "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE Jacobian sequence (k_jb_* kernels, k_scan, k_fill, k_sortrows). "
                   "'dram_bytes_per_launch' is ncu --set full with its default cold-cache replay (every kernel starts with an empty L2: upper bound); "
                   "'warm_cache' is the same sequence with --cache-control none (what a normal run sees). Sources: profiles/r01_ncu_full_batched_<config>.json, "
                   "profiles/r01_ncu_warm_<config>.json (tools/gpu_profiles.sh, tools/make_profiles.py).",
           "dram_bytes_per_launch": {}, "warm_cache_dram_bytes_per_launch": {}, "dominant_kernel": {}, "fp64_pipe_pct": {}}
for c in ("d3dHsm", "d3dHsm4x", "box2d"):
    for kind, dst in (("full", "r01_ncu_full_batched_%s.json"), ("warm", "r01_ncu_warm_%s.json")):
        it = one_iteration(json.load(open(os.path.join(G, "ncu_%s_%s.json" % (kind, c)))))
        json.dump(it, open(os.path.join(P, dst % c), "w"), indent=1)
        jac = [r for r in it if r["kernel"].startswith(JAC)]
        tot = sum(r.get("dram_read_bytes", 0) + r.get("dram_write_bytes", 0) for r in jac)
        if kind == "full":
            traffic["dram_bytes_per_launch"][c] = tot
            dom = max(jac, key=lambda r: r["duration_us"])
            traffic["dominant_kernel"][c] = {"kernel": dom["kernel"], "duration_us": dom["duration_us"],
                                             "share_of_jacobian": dom["duration_us"] / sum(r["duration_us"] for r in jac),
                                             "dram_bytes": dom.get("dram_read_bytes", 0) + dom.get("dram_write_bytes", 0)}
            traffic["fp64_pipe_pct"][c] = dom.get("fp64_pipe_pct")
        else:
            traffic["warm_cache_dram_bytes_per_launch"][c] = tot
json.dump(traffic, open(os.path.join(P, "r01_traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
