#!/bin/bash
# developer loop on the GPU box: A/B of two builds of the library (UE_GPU_LIB selects the .so) on the bench
# usage: tools/gpu_ab_lib.sh <other.so> [config]
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for rep in 1 2; do
for v in default "$1"; do
if [ $v = default ]; then unset UE_GPU_LIB; else export UE_GPU_LIB=$(pwd)/$v; fi
python bench.py --no-cpu ${2:+--config $2} 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v', {k:round(d[k],4) for k in ('ms_per_step','warm_ms_per_step','jac_kernel_ms','resid_kernel_ms')}, {k:round(d['e2e'][k],4) for k in ('ms_per_step','warm_ms_per_step','resid_evals_per_s')})
"
done
done
