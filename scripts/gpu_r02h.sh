#!/bin/bash
# Round-2 visit h (EIGHT GPUs): the whole GPU suite on the final library -- every multi-rank case at worlds 2, 4 and 8, nothing
# skipped -- then the N = 8 bench as the driver launches it (headline: D3Q19 1024^3 strong-scaled).
set -u
mkdir -p gpurun_out
T=gpurun_out/r02h
nvidia-smi -L > ${T}_gpus.txt
nvidia-smi topo -m > ${T}_topo.txt 2>&1
timeout 1500 python -m pytest tests -q -x -m gpu -rs > ${T}_pytest_gpu_n8.log 2>&1; echo "pytest rc=$? $(tail -1 ${T}_pytest_gpu_n8.log)"
grep -c SKIPPED ${T}_pytest_gpu_n8.log
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $R bench.py --gpus 8 --steps 20 --warmup 5 --also-timeout 300 > ${T}_bench_n8.json 2>${T}_bench_n8.err; echo "bench n8 rc=$?"; tail -3 ${T}_bench_n8.err
python -c "
import json
d=json.loads(open('${T}_bench_n8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','ms_per_step','scaling','e2e')}); print(d['config']); print(d['roofline']); print('\n'.join(d.get('also_summary',[])))"
