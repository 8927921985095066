"""Multi-process CPU Jacobian with the oracle: every worker holds the full state and assembles a
contiguous range of columns — the reference's MPI design (ppp/mpi_parallel.F90:2-447).  Used by
bench.py for the "all host cores" CPU arm and by the split/merge tests.  Test infrastructure only."""
import multiprocessing as mp
import os

import numpy as np

_W = {}


def _init(name, perturb):
    from tests.util import bind, make_case, oracle, psetnk_inputs
    c, yl = make_case(name, perturb=perturb)
    b = c.bbb
    ora = bind(oracle(), c)
    y, su = psetnk_inputs(c, yl)
    ora.step_params(np.full(b.neq, 1e20), y[: b.neq], su, np.ones(b.neq))
    f0 = ora.pandf1(y)
    _W.update(c=c, ora=ora, y=y, f0=f0)


def _jac_range(rng):
    c, ora = _W["c"], _W["ora"]
    b = c.bbb
    ora.pandf1(_W["y"])  # base state (OMPJacBuilder does the same before the split, omp_parallel.F90:319)
    ora.set_column_range(*rng)
    return ora.jac_calc(_W["y"], _W["f0"], b.lbw, b.ubw, b.nnzmx)


def _resid(_):
    return _W["ora"].pandf1(_W["y"])[:1]


class OraclePool:
    def __init__(self, name, perturb, nproc=None):
        self.nproc = nproc or os.cpu_count()
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.nproc, initializer=_init, initargs=(name, perturb))

    def jacobian(self, neq):
        from uedge_b200.split import merge_csr, split_index
        parts = self.pool.map(_jac_range, split_index(neq, self.nproc), chunksize=1)
        return merge_csr(parts, neq)

    def close(self):
        self.pool.close()
        self.pool.join()
