#!/bin/bash
# compute-sanitizer over one residual + Jacobian of each configuration family (writes gpurun_out/sanitizer.txt)
out=gpurun_out/sanitizer.txt; : > $out
for c in d3dHsm case1 box2d; do
  for t in memcheck racecheck initcheck; do
    echo "== $c --tool $t" >> $out
    timeout 600 compute-sanitizer --tool $t python tools/one_jac.py $c 1 2>&1 | grep -E "nnz|SUMMARY|ERROR|hazard|Invalid|Uninit" | head -8 >> $out
  done
done
cat $out
