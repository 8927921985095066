"""CPU check of the kernels' LOGIC (no GPU): the device functions of uedge_b200/csrc/ue_device.cuh compiled for the host
(tests/hostcheck) and driven by loops that mirror the kernels of ue_gpu.cu, compared bit for bit with the oracle.  This is a
developer safety net for the build container, which has no GPU; the parity tests proper are the `-m gpu` ones."""
import os
import subprocess

import numpy as np
import pytest

from tests.util import ROOT, bind, make_case, oracle, psetnk_inputs
from uedge_b200.capi import UeLib

HK_DIR = os.path.join(ROOT, "tests", "hostcheck")


@pytest.fixture(scope="module")
def hk(built):
    subprocess.check_call(["make", "-s", "-C", HK_DIR])
    return UeLib(os.path.join(HK_DIR, "libue_hostcheck.so"), "ue_hk_")


def _check(c, yl, hk, jac=True, dt=None):
    ora = bind(oracle(), c)
    bind(hk, c)
    b = c.bbb
    n = b.neq
    fo, fh = ora.pandf1(yl), hk.pandf1(yl)
    assert np.isfinite(fo).all() and np.array_equal(fo, fh), "%d of %d residual entries differ" % ((fo != fh).sum(), n)
    if not jac:
        return
    y, su = psetnk_inputs(c, yl)
    for lib in (ora, hk):
        lib.step_params(np.full(n, 1e20) if dt is None else dt, y[:n], su, np.ones(n))
    fo, fh = ora.pandf1(y), hk.pandf1(y)
    assert np.array_equal(fo, fh)
    jo, jh = ora.jac_calc(y, fo, b.lbw, b.ubw, b.nnzmx), hk.jac_calc(y, fh, b.lbw, b.ubw, b.nnzmx)
    assert np.array_equal(jo[2], jh[2]), "ia differs: nnz %d vs %d" % (len(jo[0]), len(jh[0]))
    assert np.array_equal(jo[1], jh[1]) and np.array_equal(jo[0], jh[0])


CASES = [
    ("d3dHsm", {}),
    ("d3dHsm", {"com.istabon": 7, "bbb.isngon": 0}),
    ("case1", {}),
    ("case1", {"bbb.isbcwdt": 1}),
    ("box2d", {}),
    ("box2d", {"bbb.isupcore": 1, "bbb.iflcore": 0, "bbb.xlinc": 3, "bbb.xrinc": 2, "bbb.yinc": 3}),
    ("box2d", {"bbb.isngon": 0, "com.istabon": 7, "bbb.methn": 22, "bbb.methu": 22, "bbb.methe": 22, "bbb.methi": 22, "bbb.methg": 22}),
]


@pytest.mark.parametrize("name,ov", CASES)
def test_kernel_logic_matches_oracle(hk, name, ov):
    c, yl = make_case(name, perturb=0.03, overrides=ov, seed=17)
    rng = np.random.default_rng(2)
    _check(c, yl, hk, dt=10.0 ** rng.uniform(-6, -2, c.bbb.neq) if ov.get("bbb.isbcwdt") else None)


def test_kernel_logic_wall_boundary_conditions(hk):
    from tests.test_gpu_slab import WALL_BC_SETS
    for name in ("d3dHsm", "box2d"):
        for k, ov in enumerate(WALL_BC_SETS):
            c, yl = make_case(name, perturb=0.02, overrides=ov, seed=20 + k)
            _check(c, yl, hk)


def test_kernel_logic_set_dt(hk):
    for name, model_dt, isbcwdt in (("d3dHsm", 1, 1), ("case1", 2, 0), ("box2d", 3, 1)):
        c, yl = make_case(name, perturb=1e-2, overrides={"bbb.model_dt": model_dt, "bbb.isbcwdt": isbcwdt, "bbb.dtreal": 1e-5})
        ora = bind(oracle(), c)
        bind(hk, c)
        n = c.bbb.neq
        yo = yl[:n] * 1.01 + 1e-3
        for lib in (ora, hk):
            lib.step_params(np.full(n, 1e20), yo, np.ones(n), np.ones(n))
        y = yl.copy(); y[n] = -1.0
        (fo, do), (fh, dh) = ora.set_dt(y), hk.set_dt(y)
        assert np.array_equal(fo, fh) and np.array_equal(do, dh)
        assert np.array_equal(ora.pandf1(y), hk.pandf1(y))
