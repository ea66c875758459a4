#!/bin/bash
# Dev: time variant builds of libdynode_b200.so (dynode_b200/libvar_*.so) with the kernel-only bench.
mkdir -p gpurun_out
for lib in dynode_b200/libdynode_b200.so dynode_b200/libvar_*.so; do
  for wl in c4 c3; do
    echo -n "$lib $wl " 
    DYNODE_B200_LIB=$PWD/$lib timeout 200 python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu --no-nuts 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(f\"{d['ms_per_step']:.3f} ms  {d['value']/1e6:.2f} Mtraj/s  fp64 frac {d['roofline_fp64']['frac']:.3f}\")"
  done
done
