#!/bin/bash
# Round-2 visit d (ONE GPU): which logarithm table (split / pairs) and which place for calculateAlphaMax (solver / registers)
# per lattice; D2Q9 at 6 blocks per SM; D3Q19 BGK at 5 blocks per SM.
set -u
mkdir -p gpurun_out
T=gpurun_out/r02d
one() { python -c "
import json,sys
for line in sys.stdin:
    try: d=json.loads(line)
    except Exception: continue
    r=d['roofline']; e=d.get('entropic') or {}; n=e.get('newton') or {}
    print('$1', d['config']['name'], d['dtype'][:3], d['config']['perturbation_eps'], round(d['value']), 'MLUPS', round(d['ms_per_step'],3), 'ms frac', round(r['frac'],3), 'solved', round(n.get('solved_node_fraction') or 0,3), 'evals', round(n.get('evaluations_per_solved_node') or 0,2), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
B="python bench.py --no-cpu-baseline --no-e2e --also off"
out=${T}_results.txt; : > $out
for v in "" v1 v2 v3 v4; do
  MLBM_VARIANT=$v timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden_gpu.py -q -x -m gpu -k "elbm or ELBM or log" > ${T}_parity_${v:-default}.log 2>&1
  echo "variant=${v:-default} parity rc=$? $(tail -1 ${T}_parity_${v:-default}.log)" >> $out
  for eps in 2e-2 1e-5; do
    [ "$v" != "v4" ] && MLBM_VARIANT=$v timeout 300 $B --workload d3q27_elbm_512 --steps 20 --eps $eps 2>>${T}_err.log | one "variant=${v:-default}" >> $out
    MLBM_VARIANT=$v timeout 300 $B --workload d2q9_elbm_shanchen_8192 --steps 50 --eps $eps 2>>${T}_err.log | one "variant=${v:-default}" >> $out
  done
done
cat $out
