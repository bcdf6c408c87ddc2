#!/bin/bash
# Round-2 visit f (TWO GPUs): the whole GPU suite with the final kernels (gate), the N = 2 bench as the driver launches it
# (headline = D3Q19 1024^3 strong-scaled), and the stored step with the analysis stream on / off.
set -u
mkdir -p gpurun_out
T=gpurun_out/r02f
nvidia-smi -L > ${T}_gpus.txt
timeout 1500 python -m pytest tests -q -x -m gpu -rs > ${T}_pytest_gpu_n2.log 2>&1; echo "pytest rc=$? $(tail -1 ${T}_pytest_gpu_n2.log)"
grep -c SKIPPED ${T}_pytest_gpu_n2.log
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for mode in 1 0; do
  MLBM_ASYNC_ANALYSIS=$mode timeout 600 $R bench.py --gpus 2 --workload d3q19_bgk_512 --steps 100 --warmup 5 --no-e2e --also off > ${T}_stored_async$mode.json 2>${T}_stored_async$mode.err
  python -c "
import json
d=json.loads(open('${T}_stored_async$mode.json').read().strip().splitlines()[-1])
print('async=$mode', round(d['value']), 'MLUPS ms/step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'stored_step_ms', d['config']['stored_step_ms'], d['config']['stored_mode'], d['config']['stored_steps_in_timed_region'])"
done
timeout 1500 $R bench.py --gpus 2 --steps 20 --warmup 5 --also-timeout 420 > ${T}_bench_n2.json 2>${T}_bench_n2.err; echo "bench n2 rc=$?"; tail -3 ${T}_bench_n2.err
python -c "
import json
d=json.loads(open('${T}_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','ms_per_step','scaling','e2e')}); print(d['config']); print(d['roofline']); print('\n'.join(d.get('also_summary',[])))"
