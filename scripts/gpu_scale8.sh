#!/bin/bash
# 8-GPU visit (charged 8x: keep it short): world-8 decomposition tests, then the BASELINE multi-GPU configs.
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
free -g > gpurun_out/host_memory.txt 2>&1; nproc >> gpurun_out/host_memory.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -q -k "peer-8 or sync-8 or async-8" > gpurun_out/pytest_gpu_n8.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_n8.log
tail -4 gpurun_out/pytest_gpu_n8.log
: > gpurun_out/scale8.jsonl
run() { local n=$1; shift; echo "== n=$n $*" >> gpurun_out/scale8.err
  if [ "$n" = 1 ]; then timeout 400 python bench.py --gpus 1 --no-cpu-baseline --no-e2e --also off "$@" >> gpurun_out/scale8.jsonl 2>> gpurun_out/scale8.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --no-cpu-baseline --no-e2e --also off "$@" >> gpurun_out/scale8.jsonl 2>> gpurun_out/scale8.err; fi; }
# config 5: D3Q19 BGK 1024^3 strong-scaled, observables every 50 steps
run 8 --workload d3q19_bgk_1024 --steps 100
run 8 --workload d3q19_bgk_1024 --steps 100 --halo nccl
run 4 --workload d3q19_bgk_1024 --steps 100
# config 2 weak scaling (the driver's scaling bench): 256^3 per GPU
run 1 --workload d3q19_bgk_256 --steps 200
run 2 --workload d3q19_bgk_256 --steps 200
run 4 --workload d3q19_bgk_256 --steps 200
run 8 --workload d3q19_bgk_256 --steps 200
run 8 --workload d3q19_bgk_256 --steps 200 --halo nccl
# config 3 and 4 at 8 GPUs
run 8 --workload d3q27_elbm_512 --steps 50
run 8 --workload d2q9_elbm_shanchen_8192 --steps 200 --eps 1e-5
run 8 --workload d2q9_elbm_shanchen_8192 --steps 200 --eps 1e-5 --halo nccl
run 8 --workload d2q9_elbm_shanchen_8192 --steps 200 --dtype f32
python - <<'PY'
import json
for line in open("gpurun_out/scale8.jsonl"):
    try: d = json.loads(line)
    except Exception: continue
    r = d["roofline"]; c = d["config"]
    print(f'n={d["n_gpus"]} {c["name"]:26s} {d["dtype"][:3]} eps={c["perturbation_eps"]:<6} halo={c.get("halo","")[:6]:6s} store={c["store_every"]:<3} '
          f'{d["value"]:9.0f} MLUPS {d["ms_per_step"]:8.3f} ms/step bulk {r["kernel_ms"]:.3f} ms frac {r["frac"]:.3f} stored_step {c.get("stored_step_ms")}')
PY
tail -5 gpurun_out/scale8.err
