"""CPU arm of the benchmark: the oracle's jac_calc on all host cores.  TEST / BENCH INFRASTRUCTURE, not product code.

The reference's parallel CPU design is OpenMP over contiguous column ranges with a private copy of the module state per
thread and a concatenation of the CSC fragments (ppp/omp_parallel.F90:65-117, 319-332, 395-444).  `ue_ora_jac_calc_threads`
(oracle/ue_oracle.cpp) is that design in-process: std::thread workers, per-thread state copies, a C++ merge, ranges
re-weighted by the previous call's per-thread times.  Only bench.py (cpu_baseline / --impl reference) and tests/ load this.
"""
import ctypes as C
import os
import subprocess
import time

import numpy as np

from uedge_b200.capi import UeLib

HERE = os.path.dirname(os.path.abspath(__file__))
PORTABLE = os.path.join(HERE, "libue_oracle.so")
NATIVE = os.path.join(HERE, "libue_oracle_native.so")


def oracle_lib(native=False):
    """Path of the oracle library.  native=True: (re)build it with -march=native on THIS machine when a compiler is here
    (the library that travelled to the box is built for x86-64-v3); falls back to the portable build."""
    if native:
        try:
            subprocess.run(["make", "-C", HERE, "-B", "libue_oracle_native.so"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
            return NATIVE, "g++ -O3 -march=native -ffp-contract=off -fno-fast-math (built on this host)"
        except Exception:
            pass
    return PORTABLE, "g++ -O3 -march=x86-64-v3 -ffp-contract=off -fno-fast-math"


class OracleThreads:
    """The oracle bound to one case, with the threaded Jacobian."""

    def __init__(self, c, y, su, native=False):
        path, self.flags = oracle_lib(native)
        self.ora = UeLib(path, "ue_ora_")
        self.ora.load_static(c.static_inputs()); self.ora.init()
        b = c.bbb
        self.b = b
        self.neq = int(b.neq); self.nnzmx = int(b.nnzmx)
        self.ora.step_params(np.full(self.neq, 1e20), y[: self.neq], su, np.ones(self.neq))
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        self.f0 = np.zeros(self.neq + 2)
        self.jac = np.zeros(self.nnzmx); self.ja = np.zeros(self.nnzmx, dtype=np.int64); self.ia = np.zeros(self.neq + 1, dtype=np.int64)
        fn = self.ora.lib.ue_ora_jac_calc_threads
        fn.argtypes = [C.c_int64, C.c_int64] + [C.c_void_p] * 2 + [C.c_int64] * 3 + [C.c_void_p] * 3 + [C.POINTER(C.c_int64), C.c_void_p]
        fn.restype = C.c_int
        self.fn = fn

    def residual(self):
        self.ora.pandf1(self.y, out=self.f0)
        return self.f0

    def step(self, nthreads):
        """What sfsetnk/psetnk do to get a Jacobian: rhsnk(yl) then jac_calc (bbb/oderhs.m:9466-9468).  Returns (seconds of
        the Jacobian alone, per-thread ms, nnz)."""
        self.residual()
        ms = np.zeros(max(1, nthreads)); nnz = C.c_int64(0)
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        t0 = time.perf_counter()
        rc = self.fn(nthreads, self.neq, P(self.y), P(self.f0), int(self.b.lbw), int(self.b.ubw), self.nnzmx, P(self.jac), P(self.ja), P(self.ia), C.byref(nnz), P(ms))
        dt = time.perf_counter() - t0
        if rc:
            raise RuntimeError(self.ora.lib.ue_ora_last_error().decode())
        return dt, ms, nnz.value

    def csr(self, nnz):
        return self.jac[:nnz].copy(), self.ja[:nnz].copy(), self.ia.copy()


def time_cpu_arm(c, y, su, budget_s=12.0, nthreads=None, native=True):
    """Bounded sample: repeated (residual + threaded Jacobian) steps for ~budget_s, then a few serial ones."""
    nthreads = nthreads or os.cpu_count()
    o = OracleThreads(c, y, su, native=native)
    for _ in range(3):
        o.step(nthreads)  # warm-up; also settles the range weights
    t0 = time.perf_counter(); reps = 0; tj = 0.0; tstep = 0.0
    while True:
        ts = time.perf_counter()
        dt, ms, nnz = o.step(nthreads)
        tstep += time.perf_counter() - ts
        tj += dt; reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= 200:
            break
    tj /= reps; tstep /= reps
    t1 = time.perf_counter(); r1 = 0; tj1 = 0.0
    while True:
        dt1, _, _ = o.step(1); tj1 += dt1; r1 += 1
        if time.perf_counter() - t1 > budget_s / 3 or r1 >= 20:
            break
    tj1 /= r1
    t2 = time.perf_counter(); r2 = 0
    while time.perf_counter() - t2 < 1.0:
        o.residual(); r2 += 1
    tres = (time.perf_counter() - t2) / r2
    # the same Jacobian on half the threads: on hosts whose "CPUs" are SMT siblings the second half adds little for FP64 code
    half = max(1, nthreads // 2); tjh = None
    if half < nthreads:
        for _ in range(3):
            o.step(half)
        t3 = time.perf_counter(); r3 = 0; tjh = 0.0
        while True:
            dth, _, _ = o.step(half); tjh += dth; r3 += 1
            if time.perf_counter() - t3 > budget_s / 6 or r3 >= 50:
                break
        tjh /= r3
    try:
        import psutil
        phys = psutil.cpu_count(logical=False)
    except Exception:
        phys = None
    return dict(nnz=nnz, jac_s=tj, step_s=tstep, serial_jac_s=tj1, resid_s=tres, reps=reps, threads=nthreads, flags=o.flags,
                par_eff=tj1 / (tj * nthreads), thread_ms=[float(x) for x in ms], half_threads=half, half_jac_s=tjh,
                half_par_eff=(tj1 / (tjh * half)) if tjh else None, physical_cores=phys)
