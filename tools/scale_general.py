"""Developer tool (torchrun): ONE Jacobian of the general path split over WORLD_SIZE GPUs - kernel event times, max over ranks."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from uedge_b200.capi import split_init_gen  # noqa: E402
from uedge_b200.cases import load_grid_npz, refine_grid  # noqa: E402
from uedge_b200.cases2 import d3d_full_physics_case, load_gen  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for label, f in (("16x8", 1), ("4x (64x32)", 4)):
    c, yl = d3d_full_physics_case(refine_grid(load_grid_npz(), f, f) if f > 1 else None)
    b = c.bbb
    g = load_gen().bind(c)
    f0 = g.pandf1(yl)
    if world > 1:
        split_init_gen(g.lib, world, rank, dist, torch)
    ts = []
    for rep in range(4):
        j = g.jac_calc(yl, f0, b.lbw, b.ubw, b.nnzmx)
        km = [C.c_double(0) for _ in range(4)]
        g._f("last_kernel_ms")(C.byref(km[0]), C.byref(km[1]), C.byref(km[2])); g._f("last_comm_ms")(C.byref(km[3]))
        ts.append([km[1].value, km[3].value, km[2].value])
    t = torch.tensor(np.min(np.array(ts[1:]), axis=0), device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        cols, comm, csr = t.tolist()
        print("general path, full physics %s: neq %d nnz %d on %d GPU(s): columns %.3f ms + exchange %.3f ms + CSR %.3f ms = %.3f ms" %
              (label, b.neq, len(j[0]), world, cols, comm, csr, cols + comm + csr), flush=True)
    g._f("finalize")()
if world > 1:
    dist.destroy_process_group()
