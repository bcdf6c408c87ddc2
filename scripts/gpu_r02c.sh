#!/bin/bash
# Round-2 visit c (ONE GPU): the reworked entropic kernel (split 1024-entry table, degree-5 logarithm, column in registers,
# alphaMax in the solving thread), its column-register variants, BGK at 4 / 5 / 6 blocks per SM, the whole GPU suite, the
# rewritten bench.py, and fresh ncu captures of the entropic kernels.
set -u
mkdir -p gpurun_out
T=gpurun_out/r02c
timeout 900 python -m pytest tests -q -x -m gpu > ${T}_pytest_gpu_n1.log 2>&1; echo "pytest rc=$? $(tail -1 ${T}_pytest_gpu_n1.log)"
one() { python -c "
import json,sys
for line in sys.stdin:
    try: d=json.loads(line)
    except Exception: continue
    r=d['roofline']; e=d.get('entropic') or {}; n=e.get('newton') or {}
    print('$1', d['config']['name'], d['config']['global_length'], d['dtype'][:3], d['config']['perturbation_eps'], round(d['value']), 'MLUPS', round(d['ms_per_step'],3), 'ms kernel', round(r['kernel_ms'],3), 'frac', round(r['frac'],3), 'off', e.get('alpha_off_shortcut_fraction_at_end'), 'solved', n.get('solved_node_fraction'), 'evals', n.get('evaluations_per_solved_node'), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
B="python bench.py --no-cpu-baseline --no-e2e --also off"
out=${T}_results.txt; : > $out
for v in "" m0 m2; do
  if [ -n "$v" ]; then
    MLBM_VARIANT=$v timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden_gpu.py -q -x -m gpu -k "elbm or ELBM" > ${T}_parity_$v.log 2>&1
    echo "variant=$v parity rc=$? $(tail -1 ${T}_parity_$v.log)" >> $out
  fi
  for eps in 2e-2 1e-5; do
    MLBM_VARIANT=$v timeout 300 $B --workload d3q27_elbm_512 --steps 20 --eps $eps 2>>${T}_err.log | one "variant=${v:-default}" >> $out
  done
  [ "$v" = "" ] && for eps in 2e-2 1e-5; do
    timeout 300 $B --workload d2q9_elbm_shanchen_8192 --steps 50 --eps $eps 2>>${T}_err.log | one "variant=default" >> $out
  done
  [ "$v" = "" ] && timeout 300 $B --workload d2q9_elbm_edm_8192 --dtype f32 --steps 50 2>>${T}_err.log | one "variant=default" >> $out
done
for v in "" bgk5 bgk6; do
  for shape in 256,256,256 512,512,512; do
    MLBM_VARIANT=$v timeout 300 $B --shape $shape --steps 50 2>>${T}_err.log | one "variant=${v:-bgk4(default)}" >> $out
  done
  MLBM_VARIANT=$v timeout 300 $B --workload d3q27_bgk_512 --steps 30 2>>${T}_err.log | one "variant=${v:-bgk4(default)}" >> $out
done
cat $out
prof() {  # name, launches to skip, bench arguments...
  local name=$1 skip=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fusedStep -s $skip -c 1 -f -o ${T}_$name \
      $B "$@" --steps 3 --warmup 3 > ${T}_$name.log 2>&1
  echo "ncu $name rc=$?"
}
prof d3q27_elbm_512_f64_eps2e-2 4 --workload d3q27_elbm_512 --eps 2e-2
prof d2q9_elbm_shanchen_8192_f64_eps2e-2 4 --workload d2q9_elbm_shanchen_8192 --eps 2e-2
prof d3q19_bgk_256_f64 4 --workload d3q19_bgk_256
# the rewritten bench, as the driver runs it
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > ${T}_bench_ref_n1.json 2>${T}_bench_ref_n1.err; echo "ref rc=$?"; cut -c1-600 ${T}_bench_ref_n1.json
timeout 1200 python bench.py --steps 20 --warmup 5 > ${T}_bench_n1.json 2>${T}_bench_n1.err; echo "bench rc=$?"; python -c "
import json
d=json.loads(open('${T}_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','cpu_baseline')}); print(d['roofline']); print('\n'.join(d.get('also_summary',[])))"
