#!/bin/bash
# Round-2 profiling visit (ONE GPU): FP64 peak, copy bandwidth against footprint, ncu --set full of every shipped kernel of
# the BASELINE configs, the BGK slab experiments (stride pad, registers) and the entropic variants.
set -u
mkdir -p gpurun_out
T=gpurun_out/r02b
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > ${T}_gpu.txt 2>&1
tools/_build/fp64_peak 0 gpurun_out/FP64_PEAK.json
timeout 300 python tools/copy_bw.py gpurun_out/r02b_copy_bw.json > ${T}_copy_bw.log 2>&1; tail -8 ${T}_copy_bw.log

one() { python -c "
import json,sys
for line in sys.stdin:
    try: d=json.loads(line)
    except Exception: continue
    r=d['roofline']; print('$1', d['config']['name'], d['config']['global_length'], d['dtype'][:3], d['config']['perturbation_eps'], round(d['value']), 'MLUPS', round(d['ms_per_step'],3), 'ms kernel', round(r['kernel_ms'],3), 'frac', round(r['frac'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
B="python bench.py --no-cpu-baseline --no-e2e --also off"

# --- BGK: slab shapes, stride pad, registers
out=${T}_bgk_experiments.txt; : > $out
for shape in 256,256,256 512,512,512 128,1024,1024; do
  timeout 300 $B --shape $shape --steps 50 2>>${T}_err.log | one "default" >> $out
  
  
  MLBM_VARIANT=bgk4 timeout 300 $B --shape $shape --steps 50 2>>${T}_err.log | one "bgk4(128regs)" >> $out
done
timeout 300 $B --shape 256,256,256 --dtype f32 --steps 50 2>>${T}_err.log | one "default" >> $out
MLBM_VARIANT=bgk4 timeout 300 $B --shape 256,256,256 --dtype f32 --steps 50 2>>${T}_err.log | one "bgk4" >> $out
cat $out

# --- ncu --set full of the shipped kernels
prof() {  # name, launches to skip, bench arguments...
  local name=$1 skip=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fusedStep -s $skip -c 1 -f -o gpurun_out/r02b_$name \
      $B "$@" --steps 3 --warmup 3 > gpurun_out/r02b_$name.log 2>&1
  echo "ncu $name rc=$?"
}
prof d3q27_elbm_512_f64_eps2e-2 4 --workload d3q27_elbm_512 --eps 2e-2
prof d3q27_elbm_512_f64_eps1e-5 4 --workload d3q27_elbm_512 --eps 1e-5
prof d2q9_elbm_shanchen_8192_f64_eps2e-2 4 --workload d2q9_elbm_shanchen_8192 --eps 2e-2
prof d2q9_elbm_shanchen_8192_f64_eps1e-5 4 --workload d2q9_elbm_shanchen_8192 --eps 1e-5
prof d2q9_elbm_edm_8192_f32_eps2e-2 4 --workload d2q9_elbm_edm_8192 --dtype f32 --eps 2e-2
prof d3q19_bgk_256_f64 4 --workload d3q19_bgk_256
prof d3q19_bgk_256_f32 4 --workload d3q19_bgk_256 --dtype f32
prof d3q19_bgk_512_f64 4 --shape 512,512,512
prof d3q19_bgk_slab128x1024x1024_f64 4 --shape 128,1024,1024
ls -la gpurun_out/*.ncu-rep

# --- entropic variants (parity first)
bash scripts/gpu_variants_r2.sh
