#!/bin/bash
# Round-end measurement set on ONE GPU: bench lines (both arms), launch lists and ncu --set full summaries.
# Writes only small files to gpurun_out/ (the .ncu-rep files stay in /tmp on the box).  Assembled by tools/make_profiles.py.
set -x
R=${1:-r02}
python -m pytest tests -m gpu -q 2>&1 | tail -2 > gpurun_out/${R}_gputests.txt
python bench.py > gpurun_out/${R}_bench_d3dHsm.json 2> gpurun_out/${R}_bench_d3dHsm.err
for c in case1 box2d; do python bench.py --config $c --no-grids --no-cpu > gpurun_out/${R}_bench_$c.json 2> gpurun_out/${R}_bench_$c.err; done
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${R}_ref_d3dHsm.json 2>/dev/null
for c in d3dHsm d3dHsm4x; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_$c.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-grids --config $c > gpurun_out/${R}_bench_under_ncu_$c.log 2>&1
done
# two iterations of tools/one_jac.py are enough for one complete residual + Jacobian sequence (7 / 9 kernels each)
for c in d3dHsm d3dHsm4x; do
  ncu --set full --clock-control none -c 18 -o /tmp/full_$c -f python tools/one_jac.py $c 2 > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/full_$c.ncu-rep gpurun_out/${R}_ncu_full_$c.json
  ncu --set full --clock-control none --cache-control none -c 18 -o /tmp/warm_$c -f python tools/one_jac.py $c 2 > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/warm_$c.ncu-rep gpurun_out/${R}_ncu_warm_$c.json
done
# the general path: launch list + full capture of one residual + Jacobian of pyexamples/input_example
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${R}_launches_general_input_example.csv python tools/one_gen.py input_example 2 > /dev/null 2>&1
ncu --set full --clock-control none --cache-control none -c 12 -o /tmp/gen -f python tools/one_gen.py input_example 2 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/gen.ncu-rep gpurun_out/${R}_ncu_warm_general_input_example.json
# ... and of the drift case (BASELINE configs[2]) on the 16x8 and the 4x mesh (grid-mode residual, persistent column kernel)
ncu --set full --clock-control none --cache-control none -c 12 -o /tmp/gen_j -f python tools/one_gen.py "jupyter drift case 16x8" 2 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/gen_j.ncu-rep gpurun_out/${R}_ncu_warm_general_jupyter.json
ncu --set full --clock-control none --cache-control none -c 12 -o /tmp/gen_j4 -f python tools/one_gen.py "jupyter drift case 4x" 2 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/gen_j4.ncu-rep gpurun_out/${R}_ncu_warm_general_jupyter4x.json
python tools/time_general.py --8x > gpurun_out/${R}_general_times.txt 2>&1
python tools/slowfast.py 4 >> gpurun_out/${R}_general_times.txt 2>&1
# memory checker on both paths (one residual + Jacobian each)
( compute-sanitizer --tool memcheck python tools/one_jac.py d3dHsm 1 2>&1 | tail -3; compute-sanitizer --tool memcheck python tools/one_gen.py input_example 1 2>&1 | tail -3; \
  compute-sanitizer --tool racecheck python tools/one_gen.py input_example 1 2>&1 | tail -3 ) > gpurun_out/${R}_sanitizer.txt
ls -la gpurun_out
