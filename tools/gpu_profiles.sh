#!/bin/bash
# Round-end measurement set on the GPU box: bench lines (both arms), launch lists and ncu --set full summaries.
# Writes only small files to gpurun_out/ (the .ncu-rep files stay in /tmp on the box).
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -2
for c in d3dHsm d3dHsm4x case1 box2d; do python bench.py --config $c > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; done
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/ref_d3dHsm.json 2>/dev/null
for c in d3dHsm d3dHsm4x case1 box2d; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$c.csv python bench.py --steps 2 --warmup 1 --no-cpu --config $c > gpurun_out/bench_under_ncu_$c.log 2>&1
done
# two iterations of tools/one_jac.py are enough for one complete residual + Jacobian sequence (9 / 11 kernels each)
for c in d3dHsm d3dHsm4x box2d; do
  ncu --set full --clock-control none -c 22 -o /tmp/full_$c -f python tools/one_jac.py $c 2 > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/full_$c.ncu-rep gpurun_out/ncu_full_$c.json
  ncu --set full --clock-control none --cache-control none -c 22 -o /tmp/warm_$c -f python tools/one_jac.py $c 2 > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/warm_$c.ncu-rep gpurun_out/ncu_warm_$c.json
done
ls -la gpurun_out
