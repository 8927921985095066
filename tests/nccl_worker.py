"""torchrun worker of test_nccl_split_jacobian: ONE Jacobian assembled by WORLD_SIZE GPUs through ue_gpu_comm_init; every
rank must return the full CSR of the single-GPU assembly, bit for bit."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.environ["UE_ROOT"])
from tests.util import bind, make_case, oracle, psetnk_inputs  # noqa: E402
from uedge_b200.capi import load_gpu, split_init  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for name in sys.argv[1:] or ["d3dHsm"]:
    c, yl = make_case(name, perturb=1e-3)
    b = c.bbb
    y, su = psetnk_inputs(c, yl)
    gpu = bind(load_gpu(), c)
    gpu.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
    f0 = gpu.pandf1(y)
    single = gpu.jac_calc(y, f0, b.lbw, b.ubw, b.nnzmx)
    lib = gpu.lib
    for transport in ("nccl", "p2p"):
        split_init(lib, world, rank, dist, torch, transport)
        info = [C.c_int64(0) for _ in range(5)]
        lib.ue_gpu_comm_info(*[C.byref(x) for x in info])
        f1 = gpu.pandf1(y)
        same = True
        for rep in range(5):  # (P2P alternates between two sets of slot arrays: several Jacobians in a row)
            split = gpu.jac_calc(y, f1, b.lbw, b.ubw, b.nnzmx)
            same = same and np.array_equal(f0, f1) and all(np.array_equal(p, q) for p, q in zip(single, split))
        for rep in range(3):
            f2, fused = gpu.rhs_jac(y, b.lbw, b.ubw, b.nnzmx)
            same = same and np.array_equal(f2, f0) and all(np.array_equal(p, q) for p, q in zip(single, fused))
        y2 = y.copy(); y2[: b.neq] *= 1.0 + 1e-4 * np.cos(np.arange(b.neq))  # a second state through the same slot arrays
        f3, fused2 = gpu.rhs_jac(y2, b.lbw, b.ubw, b.nnzmx)
        if rank == 0:  # and against the CPU oracle
            ora = bind(oracle(), c)
            ora.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
            fo = ora.pandf1(y)
            jo = ora.jac_calc(y, fo, b.lbw, b.ubw, b.nnzmx)
            same = same and all(np.array_equal(p, q) for p, q in zip(jo, split))
            fo2 = ora.pandf1(y2)
            jo2 = ora.jac_calc(y2, fo2, b.lbw, b.ubw, b.nnzmx)
            same = same and np.array_equal(fo2, f3) and all(np.array_equal(p, q) for p, q in zip(jo2, fused2))
        ok = ok and same
        print("rank %d %s %s: columns %d..%d, peer bytes %d, identical to single-GPU: %s" % (rank, name, transport, info[2].value, info[3].value, info[4].value, same), flush=True)
        dist.barrier()
        lib.ue_gpu_comm_finalize()
    lib.ue_gpu_finalize()
# the general path: input_example and the full switch set on the DIII-D mesh, columns split over the ranks
from uedge_b200.capi import split_init_gen  # noqa: E402
from uedge_b200.cases2 import Oracle2, d3d_full_physics_case, inputex_case, load_gen  # noqa: E402

for name, (c, yl) in (("input_example", inputex_case("default")[:2]), ("d3d full physics", d3d_full_physics_case())):
    b = c.bbb
    g = load_gen().bind(c)
    f0 = g.pandf1(yl)
    single = g.jac_calc(yl, f0, b.lbw, b.ubw, b.nnzmx)
    g.lib.ue_gen_last_error.restype = C.c_char_p
    split_init_gen(g.lib, world, rank, dist, torch)
    same = True
    for rep in range(3):
        split = g.jac_calc(yl, f0, b.lbw, b.ubw, b.nnzmx)
        same = same and all(np.array_equal(p, q) for p, q in zip(single, split))
    if rank == 0:
        o = Oracle2().bind(c)
        fo = o.pandf1(yl)
        jo = o.jac_calc(yl, fo, b.lbw, b.ubw, b.nnzmx)
        same = same and np.array_equal(fo, f0) and all(np.array_equal(p, q) for p, q in zip(jo, split))
    ok = ok and same
    print("rank %d general path %s: split over %d ranks identical to single-GPU: %s" % (rank, name, world, same), flush=True)
    dist.barrier()
    g._f("finalize")()
v = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(v, op=dist.ReduceOp.MIN)
if rank == 0 and int(v.item()) == 1:
    print("NCCL_SPLIT_OK")
dist.destroy_process_group()
