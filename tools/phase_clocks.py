"""Developer tool: per-phase clock64() deltas of k_jac blocks (library built with -DUE_JAC_PROFILE).
Usage: UE_GPU_LIB=/root/repo/gpurun_prof.so python tools/phase_clocks.py [d3dHsm|d3dHsm4x]"""
import ctypes as C
import sys
import numpy as np
sys.path.insert(0, ".")
from tests.util import make_case, psetnk_inputs, bind
from uedge_b200.capi import load_gpu

name = sys.argv[1] if len(sys.argv) > 1 else "d3dHsm"
c, yl = make_case(name, perturb=1e-3)
gpu = bind(load_gpu(), c)
b = c.bbb
y, su = psetnk_inputs(c, yl)
gpu.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
for _ in range(3):
    f = gpu.pandf1(y)
    gpu.jac_calc(y, f, b.lbw, b.ubw, b.nnzmx)
nb = 2048
out = np.zeros(nb * 8, dtype=np.int64)
assert gpu.lib.ue_gpu_debug_phase_clocks(out.ctypes.data_as(C.c_void_p), nb) == 0
t = out.reshape(nb, 8)
t = t[t[:, 0] != 0]
d = np.diff(t, axis=1)
names = ["setup+stage", "phase0", "phase1a", "phase1b", "phase2", "phase3", "compact"]
print("blocks:", len(t))
for i in range(7):
    print("%-11s median %8d  p90 %8d  max %8d cycles" % (names[i], np.median(d[:, i]), np.percentile(d[:, i], 90), d[:, i].max()))
print("total    median %8d  max %8d" % (np.median(t[:, 7] - t[:, 0]), (t[:, 7] - t[:, 0]).max()))

w = np.zeros(nb * 16, dtype=np.int64)
assert gpu.lib.ue_gpu_debug_warp_clocks(w.ctypes.data_as(C.c_void_p), nb) == 0
w = w.reshape(nb, 2, 8)[: len(t)]
print("phase 1b per warp (roles fx, fy, exe, exi, ey, -, -, -), median cycles from phase start:")
print("   ", [int(np.median(w[:, 0, k] - t[:, 3])) for k in range(8)])
print("phase 2 per warp (roles n+guard, m, e, i | second half), median cycles from phase start:")
print("   ", [int(np.median(w[:, 1, k] - t[:, 4])) for k in range(8)])
