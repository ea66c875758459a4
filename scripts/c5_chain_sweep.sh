#!/bin/bash
# Dev: config-5 NUTS run on ONE GPU at the per-GPU chain counts of the 1/2/4/8-GPU strong-scaling table.
for c in 1024 512 256 128; do
  python bench.py --workload c5 --c5-chains $c 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
n = d['nuts']
print('chains %5d  wall %.3f s  rounds %s  us/round %.1f  grad-evals/s %.3fM  leapfrogs/transition %.2f  recaptures %d' % (
    d['config']['chains_per_gpu'], n['mcmc_wall_s'], n['rounds'], 1e6 * n['mcmc_wall_s'] / n['rounds'][-1],
    d['value'] / 1e6, n['mean_leapfrogs_per_transition'], n['graph_recaptures']))"
done
