#!/bin/bash
# Dev: sampler-level NUTS grad-evals/s against the number of chains per GPU (C2 model, CUDA-graph rounds).
mkdir -p gpurun_out
for c in 4096 16384 65536 262144; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --nuts-chains $c 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['nuts']['sampler']; print(json.dumps({'chains':s['chains_per_gpu'],'grad_evals_per_s':s['value'],'wall_s':s['wall_s'],'rounds':s['rounds'],'r0':s['posterior_mean_r0'],'inf':s['posterior_mean_infectious_period']}))"
done
