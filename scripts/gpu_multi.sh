#!/bin/bash
# N-GPU visit: full -m gpu suite (multi-GPU and C++ shim tests included), then weak/strong scaling benches at N ranks.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_n$N.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_n$N.log
tail -15 gpurun_out/pytest_gpu_n$N.log
: > gpurun_out/multi_n$N.jsonl
run() { local n=$1; shift; echo "== n=$n $*" >> gpurun_out/multi_n$N.err
  if [ "$n" = 1 ]; then timeout 600 python bench.py --gpus 1 --no-cpu-baseline --also off --no-e2e "$@" >> gpurun_out/multi_n$N.jsonl 2>> gpurun_out/multi_n$N.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --no-cpu-baseline --also off --no-e2e "$@" >> gpurun_out/multi_n$N.jsonl 2>> gpurun_out/multi_n$N.err; fi; }
run 1 --workload d3q19_bgk_256 --steps 100
run $N --workload d3q19_bgk_256 --steps 100
run $N --workload d3q19_bgk_256 --steps 100 --halo nccl
run $N --workload d3q19_bgk_256 --steps 100 --overlap Off
run 1 --workload d3q19_bgk_512 --steps 50
run $N --workload d3q19_bgk_512 --steps 50
run $N --workload d3q19_bgk_512 --steps 50 --halo nccl
run $N --workload d3q19_bgk_512 --steps 50 --overlap Off
run $N --workload d2q9_elbm_shanchen_8192 --steps 100 --eps 1e-5
run $N --workload d2q9_elbm_shanchen_8192 --steps 100 --eps 1e-5 --halo nccl
run $N --workload d2q9_elbm_shanchen_8192 --steps 100 --eps 1e-5 --overlap Off
python - <<PY
import json
for line in open("gpurun_out/multi_n$N.jsonl"):
    try: d = json.loads(line)
    except Exception: continue
    r = d["roofline"]
    print(f'n={d["n_gpus"]} {d["config"]["name"]:26s} overlap={d["config"]["overlap"]:3s} halo={d["config"].get("halo","")[:6]:6s} store={d["config"]["store_every"]:<3} '
          f'{d["value"]:9.0f} MLUPS  {d["ms_per_step"]:8.3f} ms/step  bulk kernel {r["kernel_ms"]:.3f} ms  frac {r["frac"]:.3f} launches {d["gpu_launches"]}')
PY
tail -5 gpurun_out/multi_n$N.err
