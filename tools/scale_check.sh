#!/bin/bash
# N-GPU replica scaling of the bench on one box (run under gpurun --gpus 8): prints n_gpus, M nnz/s device, ms/step, M nnz/s e2e
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu 2>/dev/null | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']/1e6,1), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,1), d['config']['parallelism'])
"
done
python bench.py --no-cpu 2>/dev/null | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']/1e6,1), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,1))
"
