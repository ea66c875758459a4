#!/bin/bash
# Dev: kernel time vs resident CTAs per SM (dynamic-smem knob), C4 and C3.
for smem in 0 40000 50000 70000 100000 200000; do
  for wl in c4 c3; do
    echo -n "smem=$smem $wl "
    DYNODE_DEBUG_SMEM=$smem timeout 200 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu --no-nuts 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(f\"{d['ms_per_step']:.3f} ms\")"
  done
done
