#!/bin/bash
# developer loop on the GPU box: parity tests, short bench (both Jacobian kernel forms)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for m in 0 1; do
for c in d3dHsm d3dHsm4x; do
UE_JAC_BLOCKED=$m python bench.py --no-cpu --config $c | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('blocked=$m', '$c', {k:round(d[k],4) for k in ('ms_per_step','jac_kernel_ms','resid_kernel_ms')}, {k:round(d['e2e'][k],4) for k in ('ms_per_step','resid_evals_per_s')}, d['gpu_launches'])
"
done
done
