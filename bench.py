#!/usr/bin/env python
"""bench.py — Jacobian assembly (nnz/s) and residual evaluations (evals/s) of the pandf1/jac_calc hot path.

Contract: python bench.py --gpus N --steps K --warmup W  prints ONE JSON line.
A "step" is what sfsetnk/psetnk do to get a Jacobian (bbb/oderhs.m:9851-9857, 9466-9468): one residual
evaluation rhsnk(yl) followed by one full finite-difference Jacobian assembly jac_calc(yl, yldot00), on the
d3dHsm configuration.  Both arms (CUDA and CPU) time the same step; residual evaluations/s are reported too.
`--impl reference` times the CPU restatement of the reference algorithm (the reference itself is
MPPL/Fortran and cannot be built in this image) on the box's host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def clocks_sampler(stop, out):
    q = "clocks.sm,clocks.max.sm,clocks_throttle_reasons.active"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-i", "0"],
                               capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            out.append((float(r[0]), float(r[1]), r[2].strip()))
        except Exception:
            pass
        stop.wait(0.2)


def reasons_of(mask_strs):
    names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
    got = set()
    for m in mask_strs:
        try:
            v = int(m, 16)
        except Exception:
            continue
        for bit, n in names.items():
            if v & bit:
                got.add(n)
    return sorted(got)


def cpu_baseline(c, name, perturb, budget_s=12.0, nproc=None):
    """CPU arm: the oracle's jac_calc (same algorithm and cost structure as the reference: 2*neq windowed
    pandf1 calls) split over ALL host cores by contiguous column ranges, one process per core holding the
    full state — the reference's MPI row-split design (ppp/mpi_parallel.F90).  Also times one core alone.
    Must run before CUDA is initialised (fork)."""
    from tests.cpu_pool import OraclePool
    b = c.bbb
    nproc = nproc or os.cpu_count()
    pool = OraclePool(name, perturb, nproc=nproc)
    pool.jacobian(b.neq)  # warm-up
    t0 = time.perf_counter(); reps = 0
    while True:
        jac, ja, ia = pool.jacobian(b.neq)
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= 30:
            break
    dt = (time.perf_counter() - t0) / reps
    pool.close()
    nnz = len(jac)
    p1 = OraclePool(name, perturb, nproc=1)
    p1.jacobian(b.neq)
    t1 = time.perf_counter(); r1 = 0
    while True:
        p1.jacobian(b.neq); r1 += 1
        if time.perf_counter() - t1 > budget_s / 2 or r1 >= 10:
            break
    dt1 = (time.perf_counter() - t1) / r1
    p1.close()
    from tests.util import bind, oracle, make_case
    ora = bind(oracle(), c)
    yl = make_case(name, perturb=perturb)[1]
    t2 = time.perf_counter(); r2 = 0
    while time.perf_counter() - t2 < 1.5:
        ora.pandf1(yl); r2 += 1
    tres = (time.perf_counter() - t2) / r2
    return dict(value=nnz / dt, unit="nnz/s", cores=nproc, kind="port",
                sample="%d full Jacobians of %s (neq=%d, nnz=%d) split over %d processes: %.2f ms each; 1 process: %.2f ms; serial residual %.3f ms"
                       % (reps, name, b.neq, nnz, nproc, dt * 1e3, dt1 * 1e3, tres * 1e3),
                resid_evals_per_s=1.0 / tres, jac_s=dt, serial_value=nnz / dt1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="d3dHsm")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--mode", default="replicas", choices=["replicas", "columns"],
                    help="N>1: 'replicas' = every GPU assembles the full Jacobian of its own state (ensemble, weak scaling); "
                         "'columns' = one Jacobian, columns split over the ranks (ppp MPI design, strong scaling)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    from tests.util import make_case, psetnk_inputs
    name = a.config
    c, yl = make_case(name, perturb=1e-3, seed=1234 + (rank if a.mode == "replicas" else 0))   # replicas: every rank has its own state; columns: one shared state
    c.name = name
    b = c.bbb

    if a.impl == "reference":
        if rank != 0:
            return
        cb = cpu_baseline(c, name, 1e-3, budget_s=20.0)
        line = dict(metric="jacobian_nnz_per_s", value=cb["value"], unit="nnz/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                    ms_per_step=cb["jac_s"] * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                    impl="reference", config=dict(workload="%s full Jacobian assembly, neq=%d" % (name, b.neq)),
                    cpu_baseline=dict(value=cb["value"], unit="nnz/s", cores=cb["cores"], kind=cb["kind"], sample=cb["sample"]),
                    e2e=dict(value=cb["value"], unit="nnz/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    resid_evals_per_s=cb["resid_evals_per_s"], serial_nnz_per_s=cb["serial_value"])
        print(json.dumps(line))
        return

    cb = None
    if not a.no_cpu and world == 1:
        cb = cpu_baseline(c, name, 1e-3)  # before CUDA init (the pool forks)

    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from uedge_b200.capi import load_gpu
    gpu = load_gpu()
    gpu.load_static(c.static_inputs()); gpu.init()
    y, su = psetnk_inputs(c, yl)
    gpu.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
    lib = gpu.lib
    neq = b.neq
    # multi-GPU: replicas-with-column-split (ppp MPI design): rank r assembles columns of its contiguous iv range
    lo = 1 + (neq * rank) // world; hi = (neq * (rank + 1)) // world
    if world > 1 and a.mode == "columns":
        gpu.set_column_range(lo, hi)
    f0 = gpu.pandf1(y)
    # pinned host buffers for the end-to-end path
    hy = torch.from_numpy(y.copy()).pin_memory(); hf = torch.zeros(neq + 2, dtype=torch.float64).pin_memory(); hf[:neq] = torch.from_numpy(f0)
    nnzmx = int(b.nnzmx)
    hjac = torch.zeros(nnzmx, dtype=torch.float64).pin_memory(); hja = torch.zeros(nnzmx, dtype=torch.int64).pin_memory()
    hia = torch.zeros(neq + 1, dtype=torch.int64).pin_memory(); hyd = torch.zeros(neq, dtype=torch.float64).pin_memory()
    P = lambda t: C.cast(t.data_ptr(), C.c_void_p)
    nnz = C.c_int64(0)
    lib.ue_gpu_jac_calc.argtypes = [C.c_int64, C.c_double] + [C.c_void_p] * 2 + [C.c_int64] * 3 + [C.c_void_p] * 3 + [C.POINTER(C.c_int64)]
    lib.ue_gpu_pandf1.argtypes = [C.c_int64, C.c_double, C.c_void_p, C.c_void_p]
    bufs = [C.c_void_p() for _ in range(6)]
    lib.ue_gpu_device_buffers(*[C.byref(x) for x in bufs])
    d_yl, d_yldot, d_y00, d_jac, d_ja, d_ia = bufs
    lib.ue_gpu_jac_calc_dev.argtypes = [C.c_int64, C.c_double] + [C.c_void_p] * 2 + [C.c_int64] * 3 + [C.c_void_p] * 3 + [C.POINTER(C.c_int64)]
    lib.ue_gpu_pandf1_dev.argtypes = [C.c_int64, C.c_double, C.c_void_p, C.c_void_p]
    jms = C.c_double(0); rms = C.c_double(0)

    # what the Fortran shim does per call: step_params (+ nufak) before the residual and the Jacobian
    sp_dt = np.full(neq, 1e20); sp_yo = y[:neq].copy(); sp_su = np.ascontiguousarray(su); sp_sf = np.ones(neq)
    NP = lambda x: x.ctypes.data_as(C.c_void_p)
    lib.ue_gpu_step_params.argtypes = [C.c_int64] + [C.c_void_p] * 4
    lib.ue_gpu_set_real.argtypes = [C.c_char_p, C.c_double]
    nufak = float(c.bbb.nufak)

    def shim_params():
        assert lib.ue_gpu_step_params(neq, NP(sp_dt), NP(sp_yo), NP(sp_su), NP(sp_sf)) == 0

    jac_call_s = []

    def step_e2e():
        shim_params()
        assert lib.ue_gpu_pandf1(neq, 0.0, P(hy), P(hf)) == 0          # yldot00 = rhsnk(yl)
        tj = time.perf_counter()
        shim_params()
        assert lib.ue_gpu_set_real(b"nufak", nufak) == 0
        assert lib.ue_gpu_jac_calc(neq, 0.0, P(hy), P(hf), int(b.lbw), int(b.ubw), nnzmx, P(hjac), P(hja), P(hia), C.byref(nnz)) == 0
        jac_call_s.append(time.perf_counter() - tj)

    lib.ue_gpu_rhs_jac_dev.argtypes = [C.c_int64, C.c_void_p, C.c_void_p] + [C.c_int64] * 3 + [C.c_void_p] * 3 + [C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    evms = C.c_double(0); ev_samples = []

    def step_dev():  # inputs resident in HBM; residual + Jacobian as one stream sequence, timed by CUDA events on the library's stream
        assert lib.ue_gpu_rhs_jac_dev(neq, d_yl, d_y00, int(b.lbw), int(b.ubw), nnzmx, d_jac, d_ja, d_ia, C.byref(nnz), C.byref(evms)) == 0
        ev_samples.append(evms.value)

    def kernels_dev():  # the two sequences separately, for the per-sequence CUDA-event times
        assert lib.ue_gpu_pandf1_dev(neq, 0.0, d_yl, d_y00) == 0
        lib.ue_gpu_assume_base_current(1)
        assert lib.ue_gpu_jac_calc_dev(neq, 0.0, d_yl, d_y00, int(b.lbw), int(b.ubw), nnzmx, d_jac, d_ja, d_ia, C.byref(nnz)) == 0
        lib.ue_gpu_last_kernel_ms(C.byref(jms), C.byref(rms))

    lib.ue_gpu_rhs_jac.argtypes = [C.c_int64, C.c_void_p, C.c_void_p] + [C.c_int64] * 3 + [C.c_void_p] * 3 + [C.POINTER(C.c_int64)]

    def step_e2e_fused():  # optional integration: the pair as one C-ABI call
        shim_params()
        assert lib.ue_gpu_set_real(b"nufak", nufak) == 0
        assert lib.ue_gpu_rhs_jac(neq, P(hy), P(hf), int(b.lbw), int(b.ubw), nnzmx, P(hjac), P(hja), P(hia), C.byref(nnz)) == 0

    def resid_e2e():
        shim_params()
        assert lib.ue_gpu_pandf1(neq, 0.0, P(hy), P(hyd)) == 0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > 126 MB L2

    # a Newton iteration never repeats a state: alternate two states so that nothing is reused across steps
    ystates = [y.copy(), y.copy()]
    ystates[1][:neq] *= 1 + 1e-6 * np.random.default_rng(7).uniform(-1, 1, neq)
    ystates = [torch.from_numpy(v) for v in ystates]

    def timed(fn, steps):
        tot = 0.0; jm = []; rm = []
        for it in range(steps):
            hy.copy_(ystates[it & 1])
            flush.fill_(1.0)
            barrier()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            tot += time.perf_counter() - t0
            jm.append(jms.value); rm.append(rms.value)
        return tot, jm, rm

    step_e2e()  # fills the library's device buffers (d_yl, d_yldot00)
    for _ in range(a.warmup):
        step_dev(); step_e2e()
    samples = []; stop = threading.Event()
    th = threading.Thread(target=clocks_sampler, args=(stop, samples)); th.start()
    l0 = C.c_int64(0); lib.ue_gpu_kernel_launches(C.byref(l0))
    del ev_samples[:]
    t_dev_host, _, _ = timed(step_dev, a.steps)
    t_dev = sum(ev_samples) * 1e-3  # CUDA events: seconds for a.steps steps
    l1 = C.c_int64(0); lib.ue_gpu_kernel_launches(C.byref(l1))
    _, jm, rm = timed(kernels_dev, a.steps)
    del jac_call_s[:]
    t_e2e, _, _ = timed(step_e2e, a.steps)
    jac_call_ms = 1e3 * sum(jac_call_s) / max(1, len(jac_call_s))
    t_res_e2e, _, _ = timed(resid_e2e, a.steps)
    t_fused, _, _ = timed(step_e2e_fused, a.steps)
    stop.set(); th.join()

    def warm(fn, steps):  # same step without the L2 flush (what a Newton loop sees); reported next to the flushed figure
        barrier(); t0 = time.perf_counter()
        for it in range(steps):
            hy.copy_(ystates[it & 1]); fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / steps * 1e3
    warm_dev, warm_e2e, warm_res = warm(step_dev, 4 * a.steps), warm(step_e2e, 4 * a.steps), warm(resid_e2e, 4 * a.steps)
    nnz_local = nnz.value
    if world > 1:
        v = torch.tensor([t_dev, t_e2e, t_dev_host], device="cuda", dtype=torch.float64); dist.all_reduce(v, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, t_dev_host = v.tolist()
        n = torch.tensor([nnz_local], device="cuda", dtype=torch.int64); dist.all_reduce(n)
        nnz_total = int(n.item())
    else:
        nnz_total = nnz_local
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_dev = t_dev / a.steps * 1e3; ms_e2e = t_e2e / a.steps * 1e3
    jac_ms = float(np.mean(jm)); res_ms = float(np.mean(rm))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6453.1))
    G = 16 + 3  # static real planes + int planes the kernels read (include/ue_params.h)
    traffic = None; fp64_pct = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        traffic = tj["dram_bytes_per_launch"].get(name); fp64_pct = tj["fp64_pipe_pct"].get(name)
    except Exception:
        pass
    ncell = (c.com.nx + 2) * (c.com.ny + 2)
    alg_bytes = 8 * (2 * (neq + 2) + G * ncell) + 16 * nnz_total + 8 * (neq + 1)
    achieved = alg_bytes / (jac_ms * 1e-3) / 1e9
    sm = sorted(s[0] for s in samples) or [0.0]
    line = dict(metric="jacobian_nnz_per_s", value=nnz_total / (ms_dev * 1e-3), unit="nnz/s", n_gpus=world, steps=a.steps, warmup=a.warmup,
                ms_per_step=ms_dev, higher_is_better=True, scaling="strong" if (world > 1 and a.mode == "columns") else "weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload="%s: rhsnk + jac_calc (1 residual + 1 full FD Jacobian, neq=%d, nnz=%d) per step" % (name, neq, nnz_total),
                            l2="flushed between steps (192 MB fill)",
                            timer="value/ms_per_step: CUDA events on the library's stream around the residual+Jacobian sequence (ue_gpu_rhs_jac_dev); e2e: host clock around the two C-ABI calls; host_clock_ms_per_step: host clock around the device-resident step", parallelism=("%d independent replicas (one state per GPU), no collective" % world) if a.mode == "replicas" or world == 1
                            else "one Jacobian, columns split over %d ranks (replicated state)" % world),
                e2e=dict(value=nnz_total / (ms_e2e * 1e-3), unit="nnz/s", h2d_bytes_per_step=8 * (2 * (neq + 2) + neq),
                         d2h_bytes_per_step=16 * nnz_total + 8 * (neq + 1) + 8 * neq, ms_per_step=ms_e2e,
                         resid_evals_per_s=a.steps / t_res_e2e, warm_ms_per_step=warm_e2e, warm_resid_evals_per_s=1e3 / warm_res, fused_call_ms_per_step=t_fused / a.steps * 1e3,
                         jac_calc_call_ms=jac_call_ms, jac_calc_call_nnz_per_s=nnz_total / world / (jac_call_ms * 1e-3) * world),
                warm_ms_per_step=warm_dev, host_clock_ms_per_step=t_dev_host / a.steps * 1e3,
                gpu_launches=int(l1.value - l0.value),
                resid_evals_per_s=1e3 / res_ms if res_ms > 0 else None, jac_kernel_ms=jac_ms, resid_kernel_ms=res_ms,
                roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                              kernel="Jacobian sequence k_jb_stage0/p1a/p1b/p2/p3c + scan/fill/sort (dominant: k_jb_p2), CUDA events on the library stream", fp64_pipe_pct_of_dominant_kernel=fp64_pct, peak_source="MEASURED_PEAKS.json hbm_gbs",
                              note="algorithmic bytes = 8*(2(neq+2)+G*Ncell)+16*nnz+8*(neq+1), G=%d; latency bound (d3dHsm) / FP64-issue bound (4x), not HBM bound: DESIGN.md 3.4; traffic is ncu's cold-cache replay figure" % G),
                clocks=dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max([s[1] for s in samples] or [0.0]), reasons=reasons_of([s[2] for s in samples])))
    if cb is not None:
        line["cpu_baseline"] = dict(value=cb["value"], unit="nnz/s", cores=cb["cores"], kind=cb["kind"], sample=cb["sample"])
        line["cpu_resid_evals_per_s"] = cb["resid_evals_per_s"]
        line["cpu_serial_nnz_per_s"] = cb["serial_value"]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
