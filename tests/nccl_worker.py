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
from uedge_b200.capi import load_gpu  # noqa: E402

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
    idbuf = C.create_string_buffer(128)
    if rank == 0:
        assert lib.ue_gpu_comm_unique_id(idbuf) == 0
    t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).cuda()
    dist.broadcast(t, 0)
    lib.ue_gpu_comm_init.argtypes = [C.c_int64, C.c_int64, C.c_char_p]
    assert lib.ue_gpu_comm_init(world, rank, t.cpu().numpy().tobytes()) == 0, gpu.lib.ue_gpu_last_error()
    info = [C.c_int64(0) for _ in range(5)]
    lib.ue_gpu_comm_info(*[C.byref(x) for x in info])
    f1 = gpu.pandf1(y)
    for rep in range(2):
        split = gpu.jac_calc(y, f1, b.lbw, b.ubw, b.nnzmx)
        same = np.array_equal(f0, f1) and all(np.array_equal(p, q) for p, q in zip(single, split))
        ok = ok and same
    f2, fused = gpu.rhs_jac(y, b.lbw, b.ubw, b.nnzmx)
    ok = ok and np.array_equal(f2, f0) and all(np.array_equal(p, q) for p, q in zip(single, fused))
    if rank == 0:  # and against the CPU oracle
        ora = bind(oracle(), c)
        ora.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
        fo = ora.pandf1(y)
        jo = ora.jac_calc(y, fo, b.lbw, b.ubw, b.nnzmx)
        ok = ok and np.array_equal(jo[1], split[1]) and np.array_equal(jo[2], split[2]) and np.array_equal(jo[0], split[0])
    print("rank %d %s: columns %d..%d, nccl bytes %d, identical to single-GPU: %s" % (rank, name, info[2].value, info[3].value, info[4].value, same), flush=True)
    lib.ue_gpu_finalize()
v = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(v, op=dist.ReduceOp.MIN)
if rank == 0 and int(v.item()) == 1:
    print("NCCL_SPLIT_OK")
dist.destroy_process_group()
