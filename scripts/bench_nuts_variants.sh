#!/bin/bash
# Dev: NUTS-leg numbers (kernel C2, kernel C5, sampler) for variant builds.
for lib in dynode_b200/libdynode_b200.so dynode_b200/libvar_*.so; do
  echo -n "$lib "
  DYNODE_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1])['nuts']; print(f\"C2 kernel {d['kernel']['value']/1e6:.1f} M/s  C5 kernel {d['kernel_config5'].get('value',0)/1e6:.3f} M/s  sampler {d['sampler']['value']/1e6:.2f} M/s\")"
done
