"""CPU tests of the N>1 host logic: column split + merge, single-node multi-process (fork pool) and
torch.distributed gloo world_size=2."""
import os
import subprocess
import sys

import numpy as np

from tests.util import ROOT, bind, make_case, oracle, psetnk_inputs
from uedge_b200.split import merge_csr, split_index


def test_split_index_covers_all():
    for n, k in ((900, 1), (900, 7), (11220, 8), (13, 16)):
        r = split_index(n, k)
        assert r[0][0] == 1 and r[-1][1] == n
        assert all(r[i][1] + 1 == r[i + 1][0] for i in range(k - 1))


import pytest


@pytest.mark.parametrize("name", ["d3dHsm", "case1", "box2d"])
def test_pool_jacobian_equals_serial(built, name):
    from tests.cpu_pool import OraclePool
    c, yl = make_case(name, perturb=1e-3)
    b = c.bbb
    ora = bind(oracle(), c)
    y, su = psetnk_inputs(c, yl)
    ora.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
    f0 = ora.pandf1(y)
    ref = ora.jac_calc(y, f0, b.lbw, b.ubw, b.nnzmx)
    pool = OraclePool(name, 1e-3, nproc=3)
    got = pool.jacobian(b.neq)
    pool.close()
    assert all(np.array_equal(p, q) for p, q in zip(ref, got))


_WORKER = r'''
import os, sys, pickle
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["UE_ROOT"])
from tests.util import bind, make_case, oracle, psetnk_inputs
from uedge_b200.split import merge_csr, split_index
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
c, yl = make_case("d3dHsm", perturb=1e-3)
b = c.bbb
ora = bind(oracle(), c)
y, su = psetnk_inputs(c, yl)
ora.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
f0 = ora.pandf1(y)
lo, hi = split_index(b.neq, world)[rank]
ora.set_column_range(lo, hi)
part = ora.jac_calc(y, f0, b.lbw, b.ubw, b.nnzmx)
parts = [None] * world
dist.all_gather_object(parts, part)
if rank == 0:
    ora.set_column_range(1, b.neq)
    ora.pandf1(y)
    full = ora.jac_calc(y, f0, b.lbw, b.ubw, b.nnzmx)
    got = merge_csr(parts, b.neq)
    assert all(np.array_equal(p, q) for p, q in zip(full, got))
    print("GLOO_SPLIT_OK")
dist.destroy_process_group()
'''


def test_gloo_world2_split(built, tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    env = dict(os.environ, UE_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert "GLOO_SPLIT_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
