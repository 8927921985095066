#!/usr/bin/env python
"""bench.py — Jacobian assembly (nnz/s) and residual evaluations (evals/s) of the pandf1/jac_calc hot path.

Contract: python bench.py --gpus N --steps K --warmup W  prints ONE JSON line.
A "step" is what sfsetnk/psetnk do to get a Jacobian (bbb/oderhs.m:9851-9857, 9466-9468): one residual
evaluation rhsnk(yl) followed by one full finite-difference Jacobian assembly jac_calc(yl, yldot00), on the
d3dHsm configuration (BASELINE.json configs[1]).  Both arms (CUDA and CPU) time the same step on the same state.

N > 1 (torchrun, one rank per GPU): ONE Jacobian is assembled by the N GPUs — the reference's MPI design
(ppp/mpi_parallel.F90): every rank holds the full state, assembles a contiguous range of columns, the CSC
column results go straight into every GPU's copy of the structural slot arrays (peer stores over NVLink fused into the
assembly kernel, `--transport p2p`; or two ncclAllReduce calls, `--transport nccl`) and every rank ends with the full CSR
(`scaling: "strong"`).
The same measurement is repeated on the 4x- and 8x-refined grids (`grids`), and N independent replicas
(one state per GPU, no collective) are kept as a secondary record (`replicas`).

`--impl reference` times the CPU restatement of the reference algorithm (the reference itself is MPPL/Fortran and
cannot be built in this image) with all host threads, in-process (the ppp OpenMP design).
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


def clocks_sampler(stop, out, dev):
    q = "clocks.sm,clocks.max.sm,clocks_throttle_reasons.active"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-i", str(dev)],
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


def bench_state(name, seed=1234):
    """Case + the two alternating states every arm uses (a Newton iteration never repeats a state)."""
    from uedge_b200.cases import make_case, psetnk_inputs
    c, yl = make_case(name, perturb=1e-3, seed=seed)
    c.name = name
    y, su = psetnk_inputs(c, yl)
    neq = c.bbb.neq
    y2 = y.copy()
    y2[:neq] *= 1 + 1e-6 * np.random.default_rng(7).uniform(-1, 1, neq)
    return c, [y, y2], su


def cpu_baseline(c, ystates, su, budget_s=12.0, nthreads=None):
    """CPU arm: the oracle's jac_calc (same algorithm and cost structure as the reference: 2*neq windowed pandf1
    calls) on ALL host cores, in-process threads over contiguous column ranges with private state copies and a C++
    merge — the reference's OpenMP design (ppp/omp_parallel.F90).  Built -O3 -march=native on this host."""
    from oracle.cpu_arm import time_cpu_arm
    b = c.bbb
    r = time_cpu_arm(c, ystates[0], su, budget_s=budget_s, nthreads=nthreads, native=True)
    sample = ("%d steps (residual + full Jacobian) of %s (neq=%d, nnz=%d), state 0 of the bench pair; %s; %d in-process threads "
              "(contiguous column ranges re-weighted by measured thread times, private state copies, C++ merge): Jacobian %.2f ms, 1 thread %.2f ms, "
              "parallel efficiency %.0f%%; on %d threads %s ms = %s%% (the host reports %s physical cores for %d logical CPUs); serial residual %.3f ms"
              % (r["reps"], c.name, b.neq, r["nnz"], r["flags"], r["threads"], r["jac_s"] * 1e3, r["serial_jac_s"] * 1e3, 100 * r["par_eff"], r["half_threads"],
                 ("%.2f" % (r["half_jac_s"] * 1e3)) if r["half_jac_s"] else "-", ("%.0f" % (100 * r["half_par_eff"])) if r["half_par_eff"] else "-",
                 r["physical_cores"], os.cpu_count(), r["resid_s"] * 1e3))
    return dict(value=r["nnz"] / r["step_s"], unit="nnz/s", cores=r["threads"], kind="port", sample=sample,
                resid_evals_per_s=1.0 / r["resid_s"], step_s=r["step_s"], jac_s=r["jac_s"], serial_value=r["nnz"] / (r["serial_jac_s"] + r["resid_s"]),
                par_eff=r["par_eff"], nnz=r["nnz"])


def general_records(steps=10):
    """The GENERAL path (ue_gen_*, include/ue_gen.h) on the reference decks it exists for: pyexamples/input_example (8x4
    non-orthogonal mesh, inertial atoms, potential equation, numvar 7), pyexamples/box2 as its deck runs it (inertial atoms) and
    jupyter/case_setup.py (BASELINE configs[2]: cross-field drifts + the new potential model on the 16x8 DIII-D mesh, numvar 7).
    Host clock around the C-ABI calls with pageable host buffers (this entry point has no device-buffer variant yet); the CPU
    figure is the general oracle on one host thread, same step."""
    from uedge_b200.cases import box2_case
    from uedge_b200.cases2 import Oracle2, box2_initial_state, inputex_case, jupyter_case, load_gen
    out = {}
    c1, y1, _ = inputex_case("default")
    c2 = box2_case(isupgon=1)
    c3, y3 = jupyter_case()
    for name, c, yl in (("input_example", c1, y1), ("box2_inertial_atoms", c2, box2_initial_state(c2)), ("jupyter_drifts_potential", c3, y3)):
        b = c.bbb
        g, o = load_gen().bind(c), Oracle2().bind(c)
        f = g.pandf1(yl); fo = o.pandf1(yl)
        jg = g.jac_calc(yl, f, b.lbw, b.ubw, b.nnzmx); jo = o.jac_calc(yl, fo, b.lbw, b.ubw, b.nnzmx)
        same = bool(np.array_equal(f, fo) and all(np.array_equal(p, q) for p, q in zip(jg, jo)))
        y2 = yl.copy(); y2[: b.neq] *= 1.0 + 1e-9  # alternate two states so that every call evaluates
        # the timed calls write into caller-owned arrays, as a host code does (no allocation inside the timed region)
        fbuf = np.zeros(b.neq + 2); bufs = None
        g.pandf1(yl, out=fbuf[: b.neq])
        n, bufs = g.jac_calc_raw(yl, fbuf, b.lbw, b.ubw, b.nnzmx)
        t = time.perf_counter()
        for i in range(steps):
            yy = y2 if i & 1 else yl
            g.pandf1(yy, out=fbuf[: b.neq]); n, bufs = g.jac_calc_raw(yy, fbuf, b.lbw, b.ubw, b.nnzmx, bufs)
        tg = (time.perf_counter() - t) / steps
        j = (bufs[0][:n],)
        t = time.perf_counter()
        for i in range(steps):
            g.pandf1(y2 if i & 1 else yl, out=fbuf[: b.neq])
        tr = (time.perf_counter() - t) / steps
        t = time.perf_counter(); fo = o.pandf1(y2); tro = time.perf_counter() - t
        t = time.perf_counter(); o.jac_calc(y2, fo, b.lbw, b.ubw, b.nnzmx); tjo = time.perf_counter() - t
        out[name] = dict(neq=int(b.neq), numvar=int(b.numvar), nnz=len(j[0]), bit_identical_to_oracle=same, e2e_ms_per_step=tg * 1e3, e2e_value=len(j[0]) / tg, unit="nnz/s",
                         resid_e2e_ms=tr * 1e3, cpu_1thread_ms_per_step=(tro + tjo) * 1e3, cpu_1thread_value=len(j[0]) / (tro + tjo), kernels="k_gen_full | k_gen_full_grid (residual), k_gen_cols_q (persistent, one warp per unknown, work queue), k_gen_count/scan/fill/sortrows")
        g._f("finalize")()
    return out


class Gpu:
    """The product library bound to one case on this rank's GPU (through the C ABI only)."""

    def __init__(self, c, ystates, su, world, rank, dist, torch, split=True, transport="p2p"):
        from uedge_b200.capi import load_gpu
        self.c, self.world, self.rank, self.dist, self.torch = c, world, rank, dist, torch
        self.gpu = load_gpu()
        self.gpu.load_static(c.static_inputs()); self.gpu.init()
        b = c.bbb
        self.b = b
        self.neq = neq = int(b.neq); self.nnzmx = int(b.nnzmx)
        lib = self.lib = self.gpu.lib
        self.split = split and world > 1
        if self.split:
            from uedge_b200.capi import split_init
            split_init(lib, world, rank, dist, torch, transport)
        self.su = np.ascontiguousarray(su)
        self.gpu.step_params(np.full(neq, 1e20), ystates[0][:neq], su, np.ones(neq))
        self.sp = [np.full(neq, 1e20), ystates[0][:neq].copy(), self.su, np.ones(neq)]
        self.ystates = [torch.from_numpy(v.copy()) for v in ystates]
        self.nufak = float(b.nufak)
        # host buffers: page-locked set and pageable set
        self.pin = self._bufs(True)
        self.pag = self._bufs(False)
        P = C.c_void_p
        lib.ue_gpu_jac_calc.argtypes = [C.c_int64, C.c_double] + [P] * 2 + [C.c_int64] * 3 + [P] * 3 + [C.POINTER(C.c_int64)]
        lib.ue_gpu_pandf1.argtypes = [C.c_int64, C.c_double, P, P]
        lib.ue_gpu_rhs_jac.argtypes = [C.c_int64, P, P] + [C.c_int64] * 3 + [P] * 3 + [C.POINTER(C.c_int64)]
        lib.ue_gpu_rhs_jac_dev.argtypes = [C.c_int64, P, P] + [C.c_int64] * 3 + [P] * 3 + [C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        lib.ue_gpu_jac_calc_dev.argtypes = [C.c_int64, C.c_double] + [P] * 2 + [C.c_int64] * 3 + [P] * 3 + [C.POINTER(C.c_int64)]
        lib.ue_gpu_pandf1_dev.argtypes = [C.c_int64, C.c_double, P, P]
        lib.ue_gpu_step_params.argtypes = [C.c_int64] + [P] * 4
        lib.ue_gpu_set_real.argtypes = [C.c_char_p, C.c_double]
        lib.ue_gpu_comm_info.argtypes = [C.POINTER(C.c_int64)] * 5
        bufs = [C.c_void_p() for _ in range(6)]
        lib.ue_gpu_device_buffers(*[C.byref(x) for x in bufs])
        self.d_yl, self.d_yldot, self.d_y00, self.d_jac, self.d_ja, self.d_ia = bufs
        self.lbw, self.ubw = int(b.lbw), int(b.ubw)
        self.nnz = C.c_int64(0); self.nnz_ref = C.byref(self.nnz); self.evms = C.c_double(0); self.jms = C.c_double(0); self.rms = C.c_double(0)
        self.ev_samples = []; self.jm = []; self.rm = []
        self.flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > 126 MB L2

    def err(self):
        self.lib.ue_gpu_last_error.restype = C.c_char_p
        return self.lib.ue_gpu_last_error().decode()

    def _bufs(self, pinned):
        torch = self.torch
        mk = (lambda n, dt: torch.zeros(n, dtype=dt).pin_memory()) if pinned else (lambda n, dt: torch.zeros(n, dtype=dt))
        h = dict(y=mk(self.neq + 2, torch.float64), f=mk(self.neq + 2, torch.float64), jac=mk(self.nnzmx, torch.float64),
                 ja=mk(self.nnzmx, torch.int64), ia=mk(self.neq + 1, torch.int64), yd=mk(self.neq, torch.float64))
        h["p"] = {k: C.c_void_p(v.data_ptr()) for k, v in h.items()}  # argument objects built once: the timed region holds the C-ABI calls, not Python conversions
        return h

    def shim_params(self):
        """what the Fortran shim does per call: step_params (+ nufak) before the residual and the Jacobian"""
        if not hasattr(self, "_sp_ptrs"):
            self._sp_ptrs = [x.ctypes.data_as(C.c_void_p) for x in self.sp]
            self._dtreal = float(self.b.dtreal)
        lib = self.lib
        if lib.ue_gpu_step_params(self.neq, *self._sp_ptrs) or lib.ue_gpu_set_real(b"dtreal", self._dtreal):
            raise RuntimeError(self.err())

    # ---- the steps ------------------------------------------------------------------------------------------------
    def step_dev(self):  # inputs resident in HBM; residual + Jacobian as one stream sequence, CUDA events on the library's stream
        b = self.b
        assert self.lib.ue_gpu_rhs_jac_dev(self.neq, self.d_yl, self.d_y00, int(b.lbw), int(b.ubw), self.nnzmx, self.d_jac, self.d_ja, self.d_ia,
                                           C.byref(self.nnz), C.byref(self.evms)) == 0, self.err()
        self.ev_samples.append(self.evms.value)

    def kernels_dev(self):  # the two sequences separately, for the per-sequence CUDA-event times
        b = self.b
        assert self.lib.ue_gpu_pandf1_dev(self.neq, 0.0, self.d_yl, self.d_y00) == 0, self.err()
        self.lib.ue_gpu_assume_base_current(1)
        assert self.lib.ue_gpu_jac_calc_dev(self.neq, 0.0, self.d_yl, self.d_y00, int(b.lbw), int(b.ubw), self.nnzmx, self.d_jac, self.d_ja, self.d_ia,
                                            C.byref(self.nnz)) == 0, self.err()
        self.lib.ue_gpu_last_kernel_ms(C.byref(self.jms), C.byref(self.rms))
        self.jm.append(self.jms.value); self.rm.append(self.rms.value)

    def step_e2e(self, h):  # the two C-ABI calls of psetnk with HOST buffers h (pinned or pageable)
        p, lib = h["p"], self.lib
        self.shim_params()
        if lib.ue_gpu_pandf1(self.neq, 0.0, p["y"], p["f"]):
            raise RuntimeError(self.err())
        self.shim_params()
        if lib.ue_gpu_set_real(b"nufak", self.nufak) or lib.ue_gpu_jac_calc(self.neq, 0.0, p["y"], p["f"], self.lbw, self.ubw, self.nnzmx, p["jac"], p["ja"], p["ia"], self.nnz_ref):
            raise RuntimeError(self.err())

    def step_e2e_fused(self, h):  # optional integration: the pair as one C-ABI call
        p, lib = h["p"], self.lib
        self.shim_params()
        if lib.ue_gpu_set_real(b"nufak", self.nufak) or lib.ue_gpu_rhs_jac(self.neq, p["y"], p["f"], self.lbw, self.ubw, self.nnzmx, p["jac"], p["ja"], p["ia"], self.nnz_ref):
            raise RuntimeError(self.err())

    def resid_e2e(self, h):
        self.shim_params()
        if self.lib.ue_gpu_pandf1(self.neq, 0.0, h["p"]["y"], h["p"]["yd"]):
            raise RuntimeError(self.err())

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, fn, steps, h=None, flush=True):
        """K steps, each bracketed by barrier + synchronize; returns the summed host-clock seconds."""
        tot = 0.0
        hy = (h or self.pin)["y"]
        for it in range(steps):
            hy.copy_(self.ystates[it & 1])
            if fn in (self.step_dev, self.kernels_dev):  # device-resident: the state is put in HBM before the timed region
                self.torch.cuda.synchronize()
                self._put_state(it & 1)
            if flush:
                self.flush.fill_(1.0)
            self.barrier()
            t0 = time.perf_counter()
            fn() if h is None else fn(h)
            self.torch.cuda.synchronize()
            tot += time.perf_counter() - t0
        return tot

    def _put_state(self, k):
        t = self.torch
        if not hasattr(self, "_dyl_t"):
            # torch view of the library's device yl buffer, through the CUDA array interface
            class _W:
                pass
            w = _W()
            w.__cuda_array_interface__ = dict(shape=(self.neq + 2,), typestr="<f8", data=(self.d_yl.value, False), version=2)
            self._dyl_t = t.as_tensor(w, device="cuda")
            self._ydev = [v.cuda() for v in self.ystates]
        self._dyl_t.copy_(self._ydev[k])

    def comm_info(self):
        v = [C.c_int64(0) for _ in range(5)]
        self.lib.ue_gpu_comm_info(*[C.byref(x) for x in v])
        return [x.value for x in v]

    def close(self):
        self.lib.ue_gpu_finalize()


def maxr(g, vals):
    """max over ranks of a list of floats"""
    if g.world == 1:
        return list(vals)
    v = g.torch.tensor(list(vals), device="cuda", dtype=g.torch.float64)
    g.dist.all_reduce(v, op=g.dist.ReduceOp.MAX)
    return v.tolist()


def measure(name, steps, warmup, world, rank, dist, torch, split=True, full=True, seed=1234, transport="p2p"):
    """One configuration on this rank.  Returns a dict of max-over-ranks figures (identical on every rank)."""
    c, ystates, su = bench_state(name, seed)
    g = Gpu(c, ystates, su, world, rank, dist, torch, split=split, transport=transport)
    neq = g.neq
    g.step_e2e(g.pin)  # fills the library's device buffers
    for _ in range(max(3, warmup)):
        g._put_state(0); g.step_dev(); g.step_e2e(g.pin)
    l0 = C.c_int64(0); g.lib.ue_gpu_kernel_launches(C.byref(l0))
    del g.ev_samples[:]
    t_dev_host = g.timed(g.step_dev, steps)
    t_dev = sum(g.ev_samples) * 1e-3  # CUDA events on the library's stream: seconds for `steps` steps
    l1 = C.c_int64(0); g.lib.ue_gpu_kernel_launches(C.byref(l1))
    nnz_dev = g.nnz.value  # state (steps-1)&1
    comm = g.comm_info()
    del g.jm[:], g.rm[:]
    g.timed(g.kernels_dev, steps)
    t_e2e = g.timed(g.step_e2e, steps, g.pin)
    out = dict(name=name, neq=neq, nnz=nnz_dev, launches=int(l1.value - l0.value))
    t_dev, t_e2e, t_dev_host, jm, rm = maxr(g, [t_dev, t_e2e, t_dev_host, float(np.mean(g.jm)), float(np.mean(g.rm))])
    out.update(ms_dev=t_dev / steps * 1e3, ms_e2e=t_e2e / steps * 1e3, ms_dev_host=t_dev_host / steps * 1e3, jac_ms=jm, res_ms=rm,
               comm_bytes=comm[4], col_range=(comm[2], comm[3]))
    if full:
        t_pag = g.timed(g.step_e2e, steps, g.pag)
        t_res = g.timed(g.resid_e2e, steps, g.pin)
        t_res_pag = g.timed(g.resid_e2e, steps, g.pag)
        t_fused = g.timed(g.step_e2e_fused, steps, g.pin)
        t_fused_pag = g.timed(g.step_e2e_fused, steps, g.pag)

        def warm(fn, n, h=None):  # same step without the L2 flush (what a Newton loop sees); reported next to the flushed figure
            return g.timed(fn, n, h, flush=False) / n * 1e3
        w_dev, w_e2e, w_res = warm(g.step_dev, 2 * steps), warm(g.step_e2e, 2 * steps, g.pin), warm(g.resid_e2e, 2 * steps, g.pin)
        t_pag, t_res, t_res_pag, t_fused, t_fused_pag, w_dev, w_e2e, w_res = maxr(g, [t_pag, t_res, t_res_pag, t_fused, t_fused_pag, w_dev, w_e2e, w_res])
        out.update(ms_e2e_pageable=t_pag / steps * 1e3, ms_res_e2e=t_res / steps * 1e3, ms_res_e2e_pageable=t_res_pag / steps * 1e3,
                   ms_fused=t_fused / steps * 1e3, ms_fused_pageable=t_fused_pag / steps * 1e3, warm_ms_dev=w_dev, warm_ms_e2e=w_e2e, warm_ms_res=w_res)
    out["ncell"] = (c.com.nx + 2) * (c.com.ny + 2)
    g.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="d3dHsm")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-grids", action="store_true", help="skip the refined-grid sub-records")
    ap.add_argument("--mode", default="columns", choices=["columns", "replicas"],
                    help="N>1: 'columns' (default) = ONE Jacobian, columns split over the ranks, NCCL all-gather of the fragments (strong scaling); "
                         "'replicas' = every GPU assembles the full Jacobian of its own state (ensemble, weak scaling)")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"],
                    help="N>1, columns: 'p2p' = the assembly kernel stores each column into every GPU's slot arrays over NVLink (CUDA IPC peer "
                         "memory) + device-side flag barrier; 'nccl' = two ncclAllReduce calls over the slot arrays (the baseline)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    name = a.config

    if a.impl == "reference":
        if rank != 0:
            return
        c, ystates, su = bench_state(name)
        b = c.bbb
        cb = cpu_baseline(c, ystates, su, budget_s=20.0)
        line = dict(metric="jacobian_nnz_per_s", value=cb["value"], unit="nnz/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                    ms_per_step=cb["step_s"] * 1e3, higher_is_better=True, scaling="strong" if a.gpus > 1 else "weak", vs_baseline=None, dtype="f64", data="synthetic",
                    impl="reference",
                    config=dict(workload="%s: rhsnk + jac_calc (1 residual + 1 full FD Jacobian, neq=%d, nnz=%d) per step" % (name, b.neq, cb["nnz"])),
                    cpu_baseline=dict(value=cb["value"], unit="nnz/s", cores=cb["cores"], kind=cb["kind"], sample=cb["sample"]),
                    e2e=dict(value=cb["value"], unit="nnz/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    resid_evals_per_s=cb["resid_evals_per_s"], serial_nnz_per_s=cb["serial_value"], parallel_efficiency=cb["par_eff"])
        print(json.dumps(line))
        return

    cb = None
    if not a.no_cpu and rank == 0 and world == 1:
        c0, ys0, su0 = bench_state(name)
        cb = cpu_baseline(c0, ys0, su0)

    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    split = a.mode == "columns"
    samples = []; stop = threading.Event()
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local)); th.start()
    seed = 1234 + (rank if not split else 0)
    m = measure(name, a.steps, a.warmup, world, rank, dist, torch, split=split, full=True, seed=seed, transport=a.transport)
    stop.set(); th.join()
    grids = {}
    if not a.no_grids:
        for gname in ("d3dHsm4x", "d3dHsm8x"):
            r = measure(gname, min(a.steps, 10), 3, world, rank, dist, torch, split=split, full=False, transport=a.transport)
            grids[gname] = dict(neq=r["neq"], nnz=r["nnz"], ms_per_step=r["ms_dev"], value=r["nnz"] / (r["ms_dev"] * 1e-3), unit="nnz/s",
                                e2e_ms_per_step=r["ms_e2e"], e2e_value=r["nnz"] / (r["ms_e2e"] * 1e-3), jac_kernel_ms=r["jac_ms"], resid_kernel_ms=r["res_ms"],
                                peer_bytes_per_step_this_rank=r["comm_bytes"], gpu_launches=r["launches"])
    general = None
    if rank == 0 and world == 1 and not a.no_grids:
        try:  # a secondary record: its failure must not take the headline line with it
            general = general_records()
        except Exception as e:  # noqa: BLE001
            general = {"error": "%s: %s" % (type(e).__name__, e)}
    replicas = None
    if world > 1 and split:
        r = measure(name, min(a.steps, 10), 3, world, rank, dist, torch, split=False, full=False, seed=1234 + rank)
        n = torch.tensor([r["nnz"]], device="cuda", dtype=torch.int64); dist.all_reduce(n)
        replicas = dict(parallelism="%d independent replicas (one state per GPU), no collective" % world, nnz_total=int(n.item()),
                        ms_per_step=r["ms_dev"], value=int(n.item()) / (r["ms_dev"] * 1e-3), unit="nnz/s", scaling="weak")
    nnz_total = m["nnz"]
    if world > 1 and not split:
        n = torch.tensor([m["nnz"]], device="cuda", dtype=torch.int64); dist.all_reduce(n)
        nnz_total = int(n.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    neq = m["neq"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6453.1))
    G = 16 + 3  # static real planes + int planes the kernels read (include/ue_params.h)
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        pass
    traffic = prof.get("dram_bytes_per_launch", {}).get(name); fp64_pct = prof.get("fp64_pipe_pct", {}).get(name)
    flops = prof.get("fp64_flops_per_jacobian", {}).get(name)
    alg_bytes = 8 * (2 * (neq + 2) + G * m["ncell"]) + 16 * m["nnz"] + 8 * (neq + 1)
    achieved = alg_bytes / (m["jac_ms"] * 1e-3) / 1e9
    sm = sorted(s[0] for s in samples) or [0.0]
    if world == 1:
        par = "1 GPU"
    elif split:
        how = ("the assembly kernel stores every column fragment into all GPUs' fragment arrays over NVLink (CUDA IPC peer memory), device-side flag barrier, no collective call"
               if a.transport == "p2p" else "column fragments exchanged by grouped in-place ncclBroadcast calls on the library stream")
        par = "one Jacobian, columns split over %d ranks (replicated state, ppp MPI design); %s; every rank returns the full CSR" % (world, how)
    else:
        par = "%d independent replicas (one state per GPU), no collective" % world
    fp64_peak = 1965e6 * 148 * 64 * 2 / 1e12  # 64 FP64 FMA lanes per SM per clock at the max SM clock: 37 TFLOP/s
    line = dict(metric="jacobian_nnz_per_s", value=nnz_total / (m["ms_dev"] * 1e-3), unit="nnz/s", n_gpus=world, steps=a.steps, warmup=a.warmup,
                ms_per_step=m["ms_dev"], higher_is_better=True, scaling="strong" if (world > 1 and split) else "weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload="%s: rhsnk + jac_calc (1 residual + 1 full FD Jacobian, neq=%d, nnz=%d) per step" % (name, neq, m["nnz"]),
                            l2="flushed between steps (192 MB fill)",
                            timer="value/ms_per_step: CUDA events on the library's stream around the residual+Jacobian sequence incl. the peer exchange "
                                  "(ue_gpu_rhs_jac_dev), max over ranks; e2e: host clock around the two C-ABI calls with host buffers, max over ranks",
                            parallelism=par),
                e2e=dict(value=nnz_total / (m["ms_e2e"] * 1e-3), unit="nnz/s", h2d_bytes_per_step=8 * (2 * (neq + 2) + neq),
                         d2h_bytes_per_step=16 * m["nnz"] + 8 * (neq + 1) + 8 * neq, ms_per_step=m["ms_e2e"], buffers="page-locked host arrays (ue_gpu_pin_host_array / cudaHostAlloc)",
                         pageable_ms_per_step=m["ms_e2e_pageable"], pageable_value=nnz_total / (m["ms_e2e_pageable"] * 1e-3),
                         resid_evals_per_s=1e3 / m["ms_res_e2e"], pageable_resid_evals_per_s=1e3 / m["ms_res_e2e_pageable"],
                         warm_ms_per_step=m["warm_ms_e2e"], warm_resid_evals_per_s=1e3 / m["warm_ms_res"],
                         fused_call_ms_per_step=m["ms_fused"], fused_call_pageable_ms_per_step=m["ms_fused_pageable"]),
                warm_ms_per_step=m["warm_ms_dev"], host_clock_ms_per_step=m["ms_dev_host"],
                gpu_launches=m["launches"],
                resid_evals_per_s=1e3 / m["res_ms"] if m["res_ms"] > 0 else None, jac_kernel_ms=m["jac_ms"], resid_kernel_ms=m["res_ms"],
                peer_bytes_per_step_this_rank=m["comm_bytes"],
                roofline=dict(bound="hbm", limiter="latency: 4-6 dependent launches of 10-25 us, each a long FP64 dependency chain per thread (ncu: issue slots <25% busy, FP64 pipe <13%, "
                                                  "DRAM <1% of peak; top stalls long_scoreboard / barrier / wait - profiles/r02_ncu_*); neither HBM nor the FP64 pipe is near its roof", achieved=achieved, peak=peak, unit="GB/s",
                              frac=achieved / peak, traffic=traffic,
                              kernel="Jacobian sequence (k_jb_p01|stage0/p1a/p1b, k_jb_p2, k_jb_p3c, CSR assembly), CUDA events on the library stream",
                              fp64_pipe_pct_of_dominant_kernel=fp64_pct, peak_source="MEASURED_PEAKS.json hbm_gbs",
                              fp64=dict(flops_per_jacobian=flops, achieved_tflops=(flops / (m["jac_ms"] * 1e-3) / 1e12) if flops else None, peak_tflops=fp64_peak,
                                        frac=(flops / (m["jac_ms"] * 1e-3) / 1e12 / fp64_peak) if flops else None,
                                        how="FP64 operations the Jacobian kernels executed, from ncu's thread-level SASS counters (dadd + dmul + 2*dfma; profiles/r02_ncu_warm_*.json, tools/ncu_summary.py)"),
                              note="algorithmic bytes = 8*(2(neq+2)+G*Ncell)+16*nnz+8*(neq+1), G=%d; traffic is ncu's cold-cache replay figure" % G),
                clocks=dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max([s[1] for s in samples] or [0.0]), reasons=reasons_of([s[2] for s in samples])))
    if grids:
        line["grids"] = grids
    if general:
        line["general_path"] = general
    if replicas:
        line["replicas"] = replicas
    if cb is not None:
        line["cpu_baseline"] = dict(value=cb["value"], unit="nnz/s", cores=cb["cores"], kind=cb["kind"], sample=cb["sample"])
        line["cpu_resid_evals_per_s"] = cb["resid_evals_per_s"]
        line["cpu_serial_nnz_per_s"] = cb["serial_value"]
        line["cpu_parallel_efficiency"] = cb["par_eff"]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
