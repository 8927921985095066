"""The in-process threaded CPU arm (oracle/cpu_arm.py, ue_ora_jac_calc_threads) returns the serial oracle Jacobian bit for
bit, for any thread count, and the optimised builds of the oracle do not change a single bit of the result."""
import os
import subprocess

import numpy as np
import pytest

from tests.util import ROOT, bind, make_case, oracle, psetnk_inputs


@pytest.mark.parametrize("name", ["d3dHsm", "case1", "box2d"])
def test_threaded_jacobian_equals_serial(built, name):
    from oracle.cpu_arm import OracleThreads
    c, yl = make_case(name, perturb=1e-3)
    b = c.bbb
    y, su = psetnk_inputs(c, yl)
    ora = bind(oracle(), c)
    ora.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
    f0 = ora.pandf1(y)
    ref = ora.jac_calc(y, f0, b.lbw, b.ubw, b.nnzmx)
    o = OracleThreads(c, y, su)
    for nt in (1, 3, 8, 3):  # the second 3-thread call runs with re-weighted (uneven) ranges
        _, ms, nnz = o.step(nt)
        got = o.csr(nnz)
        assert all(np.array_equal(p, q) for p, q in zip(ref, got)), "threads=%d" % nt


def test_native_build_is_bit_identical(built):
    """-O3 -march=native (the build the CPU arm is timed with) against the portable build: same residual, same Jacobian
    values, same pattern (FP contraction and fast-math are off in both)."""
    from oracle.cpu_arm import NATIVE, oracle_lib
    from uedge_b200.capi import UeLib
    path, flags = oracle_lib(native=True)
    if path != NATIVE:
        pytest.skip("no compiler on this host")
    c, yl = make_case("d3dHsm", perturb=1e-3)
    b = c.bbb
    y, su = psetnk_inputs(c, yl)
    res = []
    for lib in (oracle(), UeLib(path, "ue_ora_")):
        o = bind(lib, c)
        o.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
        f0 = o.pandf1(y)
        res.append((f0.copy(),) + tuple(o.jac_calc(y, f0, b.lbw, b.ubw, b.nnzmx)))
    assert all(np.array_equal(p, q) for p, q in zip(*res))
