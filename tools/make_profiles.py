#!/usr/bin/env python
"""Assemble profiles/<round>_* from the files tools/gpu_profiles.sh left in gpurun_out/ (run here after the GPU call).
Usage: python tools/make_profiles.py [r02]"""
import json
import os
import shutil
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = "gpurun_out", "profiles"
JAC = ("k_jb_stage0", "k_jb_p01", "k_jb_p1a", "k_jb_p1b", "k_jb_p2", "k_jb_p3c", "k_csr")


def last_json(path):
    return [l for l in open(path) if l.startswith("{")][-1]


def one_iteration(rows, last="k_csr"):
    """Kernels of the LAST complete residual + Jacobian iteration (ends with the CSR kernel)."""
    ends = [i for i, r in enumerate(rows) if r["kernel"].startswith(last)]
    e = ends[-1]
    s = ends[-2] + 1 if len(ends) > 1 else 0
    return rows[s : e + 1]


for c in ("d3dHsm", "case1", "box2d"):
    open(os.path.join(P, "%s_bench_%s.json" % (R, c)), "w").write(last_json(os.path.join(G, "%s_bench_%s.json" % (R, c))))
open(os.path.join(P, "%s_bench_reference_arm_d3dHsm.json" % R), "w").write(last_json(os.path.join(G, "%s_ref_d3dHsm.json" % R)))
for c in ("d3dHsm", "d3dHsm4x", "general_input_example"):
    shutil.copy(os.path.join(G, "%s_launches_%s.csv" % (R, c)), os.path.join(P, "%s_launches_%s.csv" % (R, c)))
for f in ("sanitizer.txt", "general_times.txt", "gputests.txt"):
    shutil.copy(os.path.join(G, "%s_%s" % (R, f)), os.path.join(P, "%s_%s" % (R, f)))

traffic = {"note": "Per Jacobian sequence (k_jb_* kernels + k_csr): DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and FP64 operations executed "
                   "(thread-level SASS counts dadd + dmul + 2*dfma).  'dram_bytes_per_launch' is ncu --set full with its default cold-cache replay (every kernel "
                   "starts with an empty L2: upper bound); 'warm_cache' is the same sequence with --cache-control none (what a normal run sees).  Sources: "
                   "profiles/%s_ncu_full_<config>.json, profiles/%s_ncu_warm_<config>.json (tools/gpu_profiles.sh, tools/make_profiles.py)." % (R, R),
           "dram_bytes_per_launch": {}, "warm_cache_dram_bytes_per_launch": {}, "dominant_kernel": {}, "fp64_pipe_pct": {}, "fp64_flops_per_jacobian": {},
           "fp64_flops_per_residual": {}, "kernel_us_warm": {}}
for c in ("d3dHsm", "d3dHsm4x"):
    for kind in ("full", "warm"):
        it = one_iteration(json.load(open(os.path.join(G, "%s_ncu_%s_%s.json" % (R, kind, c)))))
        json.dump(it, open(os.path.join(P, "%s_ncu_%s_%s.json" % (R, kind, c)), "w"), indent=1)
        jac = [r for r in it if r["kernel"].startswith(JAC)]
        res = [r for r in it if r["kernel"].startswith("k_phase")]
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
            traffic["fp64_flops_per_jacobian"][c] = sum(r.get("fp64_flop", 0) for r in jac)
            traffic["fp64_flops_per_residual"][c] = sum(r.get("fp64_flop", 0) for r in res)
            traffic["kernel_us_warm"][c] = [[r["kernel"], round(r["duration_us"], 2)] for r in it]
# the general path: one residual (k_gen_full) + one Jacobian (k_gen_cols + 4 CSR kernels) of pyexamples/input_example
rows = json.load(open(os.path.join(G, "%s_ncu_warm_general_input_example.json" % R)))
it = one_iteration(rows, last="k_gen_sortrows")
json.dump(it, open(os.path.join(P, "%s_ncu_warm_general_input_example.json" % R), "w"), indent=1)
traffic["general_input_example"] = {"kernel_us_warm": [[r["kernel"], round(r["duration_us"], 2)] for r in it],
                                    "fp64_flops_per_jacobian": sum(r.get("fp64_flop", 0) for r in it if r["kernel"].startswith("k_gen_cols")),
                                    "dram_bytes_per_jacobian_warm": sum(r.get("dram_read_bytes", 0) + r.get("dram_write_bytes", 0) for r in it if not r["kernel"].startswith("k_gen_full"))}
# ... and of the drift case (jupyter/case_setup.py) on the 16x8 and the 4x mesh
for tag in ("jupyter", "jupyter4x"):
    src = os.path.join(G, "%s_ncu_warm_general_%s.json" % (R, tag))
    if not os.path.exists(src):
        continue
    it = one_iteration(json.load(open(src)), last="k_gen_sortrows")
    json.dump(it, open(os.path.join(P, "%s_ncu_warm_general_%s.json" % (R, tag)), "w"), indent=1)
    traffic["general_" + tag] = {"kernel_us_warm": [[r["kernel"], round(r["duration_us"], 2)] for r in it],
                                 "fp64_flops_per_jacobian": sum(r.get("fp64_flop", 0) for r in it if r["kernel"].startswith("k_gen_cols")),
                                 "dram_bytes_per_jacobian_warm": sum(r.get("dram_read_bytes", 0) + r.get("dram_write_bytes", 0) for r in it if not r["kernel"].startswith("k_gen_full"))}
json.dump(traffic, open(os.path.join(P, "%s_traffic.json" % R), "w"), indent=1)
print(json.dumps(traffic, indent=1))
