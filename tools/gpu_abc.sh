#!/bin/bash
# A/B/C... of several builds of the library on the bench: tools/gpu_abc.sh <config> lib1.so lib2.so ...
cfg=$1; shift
for rep in 1 2; do
for v in default "$@"; do
if [ $v = default ]; then unset UE_GPU_LIB; else export UE_GPU_LIB=$(pwd)/$v; fi
python bench.py --no-cpu --config $cfg 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$cfg $v', {k:round(d[k],4) for k in ('ms_per_step','warm_ms_per_step','jac_kernel_ms','resid_kernel_ms')}, {k:round(d['e2e'][k],4) for k in ('ms_per_step','warm_ms_per_step')})
"
done
done
