#!/bin/bash
# developer loop on the GPU box: parity tests, then a short bench of both grids and the host-overhead breakdown
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for p in 0; do
for c in d3dHsm d3dHsm4x; do
UE_GPU_NO_PDL=$p python bench.py --no-cpu --config $c | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('nopdl=$p', '$c', {k:round(d[k],4) for k in ('ms_per_step','warm_ms_per_step','jac_kernel_ms','resid_kernel_ms')}, {k:round(d['e2e'][k],4) for k in ('ms_per_step','warm_ms_per_step','resid_evals_per_s')})
"
done
done
python tools/host_overhead.py d3dHsm
