#!/bin/bash
# ncu --set full captures of the fused kernel for the workloads named on the command line ("workload:dtype:eps" ...)
set -u
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read -r workload dtype eps <<< "$spec"
  extra=""; [ -n "${eps:-}" ] && extra="--eps $eps"
  name="prof_${workload}_${dtype}${eps:+_eps$eps}"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fusedStep -s 4 -c 1 -f -o gpurun_out/$name \
      python bench.py --workload $workload --dtype $dtype $extra --steps 3 --warmup 3 --no-cpu-baseline --also off --no-e2e > gpurun_out/$name.log 2>&1
  echo "$name rc=$?"
done
