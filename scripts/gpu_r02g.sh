#!/bin/bash
# Round-2 visit g (ONE GPU): the whole GPU suite on the library with FP32 arithmetic for FP32 storage, 32-bit in-population
# offsets and the analysis stream; FP32 / FP64 BGK rates; the stored step with the analysis stream on / off.
set -u
mkdir -p gpurun_out
T=gpurun_out/r02g
timeout 900 python -m pytest tests -q -x -m gpu > ${T}_pytest_gpu_n1.log 2>&1; echo "pytest rc=$? $(tail -1 ${T}_pytest_gpu_n1.log)"; grep -E "^(FAILED|ERROR)|Error|assert" ${T}_pytest_gpu_n1.log | head -20
one() { python -c "
import json,sys
for line in sys.stdin:
    try: d=json.loads(line)
    except Exception: continue
    r=d['roofline']
    print('$1', d['config']['name'], d['config']['global_length'], d['dtype'][:3], round(d['value']), 'MLUPS', round(d['ms_per_step'],3), 'ms kernel', round(r['kernel_ms'],3), 'frac', round(r['frac'],3), 'fp64', r.get('fp64_frac'), 'stored_ms', d['config'].get('stored_step_ms'), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
B="python bench.py --no-cpu-baseline --no-e2e --also off"
out=${T}_results.txt; : > $out
timeout 300 $B --steps 50 2>>${T}_err.log | one f64 >> $out
timeout 300 $B --steps 50 --dtype f32 2>>${T}_err.log | one f32 >> $out
timeout 300 $B --workload d3q19_bgk_guo_256 --steps 50 --dtype f32 2>>${T}_err.log | one f32 >> $out
timeout 300 $B --workload d3q27_bgk_512 --steps 30 --dtype f32 2>>${T}_err.log | one f32 >> $out
timeout 300 $B --workload d3q27_bgk_512 --steps 30 2>>${T}_err.log | one f64 >> $out
for mode in 1 0; do
  MLBM_ASYNC_ANALYSIS=$mode timeout 300 $B --workload d3q19_bgk_512 --steps 100 2>>${T}_err.log | one "async=$mode" >> $out
done
timeout 300 $B --workload d3q27_elbm_512 --steps 20 --eps 2e-2 2>>${T}_err.log | one elbm >> $out
timeout 300 $B --workload d2q9_elbm_shanchen_8192 --steps 50 --eps 2e-2 2>>${T}_err.log | one elbm >> $out
timeout 300 $B --workload d2q9_elbm_edm_8192 --dtype f32 --steps 50 2>>${T}_err.log | one elbm >> $out
cat $out; tail -5 ${T}_err.log
