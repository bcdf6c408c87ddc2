#!/bin/bash
# Round-2 experiment: the entropic kernel variants that were written (and logic-checked in the emulator) without a GPU.
#   fast   = -DMLBM_ELBM_FASTPATH          optimistic register pass for blocks whose nodes all take alpha = 2 (per-block hint)
#   pf     = -DMLBM_PREFETCH_NEXT_PLANE    L2 prefetch of the block's next plane while the current one is solved
#   fastpf = both
# Build the variant libraries HERE first (they travel with the snapshot):
#   for v in fast:-DMLBM_ELBM_FASTPATH pf:-DMLBM_PREFETCH_NEXT_PLANE "fastpf:-DMLBM_ELBM_FASTPATH -DMLBM_PREFETCH_NEXT_PLANE"; do
#     MLBM_VARIANT=${v%%:*} MLBM_EXTRA_FLAGS="${v#*:}" python -m metalbm_b200.build; done
# then: gpurun --timeout 1500 -- 'bash scripts/gpu_variants_r2.sh'
set -u
mkdir -p gpurun_out
out=gpurun_out/elbm_variants_r2.txt
: > $out
one() { python -c "import json,sys; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('$1', d['config']['name'], d['dtype'][:3], d['config']['perturbation_eps'], round(d['value']), 'MLUPS', round(d['ms_per_step'],3), 'ms  frac', round(r['frac'],3))"; }
for v in "" fast pf; do
  if [ -n "$v" ] && [ ! -f metalbm_b200/libmetalbm_b200_$v.so ]; then echo "variant $v not built" >> $out; continue; fi
  # parity of the variant before its speed means anything
  MLBM_VARIANT=$v timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden_gpu.py -q -x -m gpu -k "elbm or ELBM" > gpurun_out/parity_variant_${v:-default}.log 2>&1
  echo "variant=${v:-default} parity rc=$? $(tail -1 gpurun_out/parity_variant_${v:-default}.log)" >> $out
  for w in d3q27_elbm_512:20 d2q9_elbm_shanchen_8192:50; do
    for eps in 2e-2 1e-5; do
      MLBM_VARIANT=$v timeout 300 python bench.py --no-cpu-baseline --no-e2e --also off --workload ${w%%:*} --steps ${w#*:} --eps $eps 2>/dev/null | one "variant=${v:-default}" >> $out
    done
  done
  MLBM_VARIANT=$v timeout 300 python bench.py --no-cpu-baseline --no-e2e --also off --workload d2q9_elbm_edm_8192 --dtype f32 --steps 50 2>/dev/null | one "variant=${v:-default}" >> $out
done
cat $out
