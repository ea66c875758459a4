#!/bin/bash
# Round evidence for the other two kernels: C3 (seasonal SEIRS, bench --workload c3) and the fused log-likelihood (C2).
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:lane_solver_kernel -s 3 -c 1 -f -o gpurun_out/prof_c3_s3 python bench.py --workload c3 --steps 2 --warmup 3 --no-e2e --no-cpu --no-nuts > gpurun_out/ncu_full_c3_s3.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:lane_solver_kernel -s 3 -c 1 -f -o gpurun_out/prof_c2lik_s3 python scripts/loglik_kernel_bench.py > gpurun_out/ncu_full_c2lik_s3.log 2>&1
ls -la gpurun_out/prof_c3_s3.ncu-rep gpurun_out/prof_c2lik_s3.ncu-rep
