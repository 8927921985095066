"""Column-range split of the Jacobian across workers and the merge of their results.

Host-side mirror of the reference's parallel-Jacobian driver (`ppp/`):
  split_index   ~ MPISplitIndex / OMPSplitIndex   (ppp/mpi_parallel.F90, ppp/omp_parallel.F90:395-444):
                  contiguous ranges of unknowns iv, optionally weighted by measured per-worker time
  merge_csr     ~ MPICollectBroadCastJacobian / OMPCollectJacobian (ppp/mpi_parallel.F90:262-364,
                  ppp/omp_parallel.F90:65-117): the reference concatenates per-worker CSC fragments in
                  worker (= iv) order and transposes with csrcsc; our workers return CSR restricted to
                  their columns, so the merge concatenates, row by row, the workers' segments in worker
                  order — columns stay ascending because the ranges are ordered.
"""
import numpy as np


def split_index(neq, nworkers, weights=None):
    """1-based inclusive (ivmin, ivmax) per worker; weights ~ relative speed (uniform by default)."""
    w = np.ones(nworkers) if weights is None else np.asarray(weights, dtype=float)
    edges = np.concatenate([[0.0], np.cumsum(w / w.sum())])
    cuts = np.rint(edges * neq).astype(np.int64)
    cuts[0], cuts[-1] = 0, neq
    return [(int(cuts[i]) + 1, int(cuts[i + 1])) for i in range(nworkers)]


def merge_csr(parts, neq):
    """parts: list of (jac, ja, ia) in worker order, each 1-based CSR over all neq rows but only the
    worker's columns.  Returns the full (jac, ja, ia)."""
    counts = np.zeros(neq, dtype=np.int64)
    for _, _, ia in parts:
        counts += np.diff(ia)
    ia_out = np.empty(neq + 1, dtype=np.int64)
    ia_out[0] = 1
    np.cumsum(counts, out=ia_out[1:])
    ia_out[1:] += 1
    nnz = int(ia_out[-1] - 1)
    jac = np.empty(nnz)
    ja = np.empty(nnz, dtype=np.int64)
    fill = ia_out[:-1] - 1
    for v, c, ia in parts:
        n = np.diff(ia)
        rows = np.repeat(np.arange(neq), n)
        # position inside the row = running offset of that row + index within this worker's segment
        within = np.arange(len(v)) - np.repeat(ia[:-1] - 1, n)
        pos = fill[rows] + within
        jac[pos] = v
        ja[pos] = c
        fill = fill + n
    return jac, ja, ia_out
