"""Developer tool: wall-clock cost of the host entry points vs their device-pointer forms (no L2 flush)."""
import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from tests.util import make_case, psetnk_inputs, bind
from uedge_b200.capi import load_gpu
name = sys.argv[1] if len(sys.argv) > 1 else "d3dHsm"
c, yl = make_case(name, perturb=1e-3)
gpu = bind(load_gpu(), c); lib = gpu.lib; b = c.bbb; neq = b.neq
y, su = psetnk_inputs(c, yl)
gpu.step_params(np.full(neq, 1e20), y[:neq], su, np.ones(neq))
f0 = gpu.pandf1(y)
hy = [torch.from_numpy(y.copy()).pin_memory() for _ in range(2)]
hy[1][:neq] *= 1 + 1e-9
hf = torch.zeros(neq + 2, dtype=torch.float64).pin_memory()
nnzmx = int(b.nnzmx)
hjac = torch.zeros(nnzmx, dtype=torch.float64).pin_memory(); hja = torch.zeros(nnzmx, dtype=torch.int64).pin_memory(); hia = torch.zeros(neq + 1, dtype=torch.int64).pin_memory()
P = lambda t: C.cast(t.data_ptr(), C.c_void_p)
nnz = C.c_int64(0)
lib.ue_gpu_jac_calc.argtypes = [C.c_int64, C.c_double] + [C.c_void_p] * 2 + [C.c_int64] * 3 + [C.c_void_p] * 3 + [C.POINTER(C.c_int64)]
lib.ue_gpu_pandf1.argtypes = [C.c_int64, C.c_double, C.c_void_p, C.c_void_p]
bufs = [C.c_void_p() for _ in range(6)]
lib.ue_gpu_device_buffers(*[C.byref(x) for x in bufs])
d_yl, d_yldot, d_y00, d_jac, d_ja, d_ia = bufs
lib.ue_gpu_jac_calc_dev.argtypes = lib.ue_gpu_jac_calc.argtypes
lib.ue_gpu_pandf1_dev.argtypes = lib.ue_gpu_pandf1.argtypes
def t(fn, n=300):
    for _ in range(20): fn(0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
print(name)
print("pandf1 host (alternating yl)  %.1f us" % t(lambda i: lib.ue_gpu_pandf1(neq, 0.0, P(hy[i & 1]), P(hf))))
print("pandf1 dev                    %.1f us" % t(lambda i: lib.ue_gpu_pandf1_dev(neq, 0.0, d_yl, d_y00)))
def step_host(i):
    lib.ue_gpu_pandf1(neq, 0.0, P(hy[i & 1]), P(hf))
    lib.ue_gpu_jac_calc(neq, 0.0, P(hy[i & 1]), P(hf), int(b.lbw), int(b.ubw), nnzmx, P(hjac), P(hja), P(hia), C.byref(nnz))
def step_dev(i):
    lib.ue_gpu_pandf1_dev(neq, 0.0, d_yl, d_y00); lib.ue_gpu_assume_base_current(1)
    lib.ue_gpu_jac_calc_dev(neq, 0.0, d_yl, d_y00, int(b.lbw), int(b.ubw), nnzmx, d_jac, d_ja, d_ia, C.byref(nnz))
print("rhsnk+jac host                %.1f us" % t(step_host, 200))
print("rhsnk+jac dev                 %.1f us" % t(step_dev, 200))
jm = C.c_double(); rm = C.c_double(); lib.ue_gpu_last_kernel_ms(C.byref(jm), C.byref(rm))
print("last kernel ms: jac %.1f us  resid %.1f us" % (jm.value * 1e3, rm.value * 1e3))
print("noop ctypes call              %.2f us" % t(lambda i: lib.ue_gpu_kernel_launches(C.byref(nnz)), 2000))
