set -u
mkdir -p gpurun_out
one() { python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', d['config']['name'], d['config']['perturbation_eps'], round(d['value']), round(d['ms_per_step'],3))"; }
for v in "" q9x4 q9x6; do
  for eps in 2e-2 1e-5; do
  MLBM_VARIANT=$v timeout 300 python bench.py --no-cpu-baseline --also off --no-e2e --workload d2q9_elbm_shanchen_8192 --steps 50 --eps $eps 2>/dev/null | one "variant=$v"
  done
done
for ppb in 1 2 8 16; do
  MLBM_PLANES_PER_BLOCK=$ppb timeout 300 python bench.py --no-cpu-baseline --also off --no-e2e --workload d2q9_elbm_shanchen_8192 --steps 50 --eps 1e-5 2>/dev/null | one "ppb=$ppb"
  MLBM_PLANES_PER_BLOCK=$ppb timeout 300 python bench.py --no-cpu-baseline --also off --no-e2e --workload d3q27_elbm_512 --steps 20 2>/dev/null | one "ppb=$ppb"
done
MLBM_PLANES_PER_BLOCK=1 timeout 300 python bench.py --no-cpu-baseline --also off --no-e2e --workload d3q27_elbm_512 --steps 20 --eps 1e-5 2>/dev/null | one "ppb=1"
timeout 300 python bench.py --no-cpu-baseline --also off --no-e2e --workload d3q27_elbm_512 --steps 20 --eps 1e-5 2>/dev/null | one "ppb=4"
