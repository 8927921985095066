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
    import time
    c, ora = _W["c"], _W["ora"]
    b = c.bbb
    t0 = time.perf_counter()
    ora.pandf1(_W["y"])  # base state (OMPJacBuilder does the same before the split, omp_parallel.F90:319)
    ora.set_column_range(*rng)
    out = ora.jac_calc(_W["y"], _W["f0"], b.lbw, b.ubw, b.nnzmx)
    return out, time.perf_counter() - t0


def _resid(_):
    return _W["ora"].pandf1(_W["y"])[:1]


class OraclePool:
    def __init__(self, name, perturb, nproc=None):
        self.nproc = nproc or os.cpu_count()
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.nproc, initializer=_init, initargs=(name, perturb))

        self.weights = None

    def jacobian(self, neq):
        """One parallel assembly.  Ranges are re-balanced with the per-worker times of the previous call,
        as the reference does (ppp/omp_parallel.F90:199-229, 395-444)."""
        from uedge_b200.split import merge_csr, split_index
        ranges = split_index(neq, self.nproc, self.weights)
        res = self.pool.map(_jac_range, ranges, chunksize=1)
        parts = [r[0] for r in res]
        times = np.array([r[1] for r in res])
        n = np.array([hi - lo + 1 for lo, hi in ranges], dtype=float)
        speed = n / np.maximum(times, 1e-9)  # unknowns per second in each range
        w = (1.0 / self.nproc) if self.weights is None else np.asarray(self.weights)
        # new range sizes ~ proportional to the measured speed, damped
        neww = speed / speed.sum()
        self.weights = 0.5 * (np.full(self.nproc, 1.0 / self.nproc) if self.weights is None else w) + 0.5 * neww
        return merge_csr(parts, neq)

    def close(self):
        self.pool.close()
        self.pool.join()
