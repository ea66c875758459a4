#!/bin/bash
# Round evidence: (1) the plain bench line, (2) the ncu launch list of the same command, (3) one --set full capture
# of the dominant kernel at the bench's own ensemble size.  usage: scripts/ncu_capture.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-nuts"
$CMD > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv $CMD > gpurun_out/ncu_launch_${tag}.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:lane_solver_kernel -s 3 -c 1 -f -o gpurun_out/prof_${tag} $CMD > gpurun_out/ncu_full_${tag}.log 2>&1
ls -la gpurun_out/prof_${tag}.ncu-rep
