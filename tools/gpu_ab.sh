#!/bin/bash
# developer loop on the GPU box: parity tests, then an A/B of an environment toggle on the d3dHsm bench
# usage: tools/gpu_ab.sh UE_GPU_NO_FUSE23
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for rep in 1 2; do
for v in 0 1; do
if [ $v = 1 ]; then export $1=1; else unset $1; fi
python bench.py --no-cpu ${2:+--config $2} 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1=$v', {k:round(d[k],4) for k in ('ms_per_step','warm_ms_per_step','jac_kernel_ms','resid_kernel_ms')}, {k:round(d['e2e'][k],4) for k in ('ms_per_step','warm_ms_per_step','resid_evals_per_s')})
"
done
done
