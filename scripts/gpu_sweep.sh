#!/bin/bash
# GPU-box visit: all parity tests (no -x), then the single-GPU throughput of every BASELINE workload.
set -u
mkdir -p gpurun_out
free -g > gpurun_out/host.txt; nproc >> gpurun_out/host.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
: > gpurun_out/sweep.jsonl
run() { echo "== $*" >> gpurun_out/sweep.err; timeout 600 python bench.py --no-cpu-baseline --also off --no-e2e "$@" >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err; }
run --workload d3q19_bgk_256 --steps 100
run --workload d3q19_bgk_256 --steps 100 --dtype f32
run --workload d3q19_bgk_guo_256 --steps 100
run --workload d3q19_bgk_512 --steps 50 --store-every 0
run --workload d3q19_bgk_512 --steps 50 --store-every 10
run --workload d3q27_elbm_512 --steps 20
run --workload d3q27_elbm_512 --steps 20 --eps 1e-5
run --workload d3q27_elbm_512 --steps 20 --dtype f32
run --workload d2q9_elbm_shanchen_8192 --steps 50
run --workload d2q9_elbm_shanchen_8192 --steps 50 --eps 1e-5
run --workload d2q9_elbm_edm_8192 --steps 50
run --workload d2q9_elbm_shanchen_8192 --steps 50 --dtype f32
run --workload d2q9_elbm_edm_8192 --steps 50 --dtype f32
tail -5 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for line in open("gpurun_out/sweep.jsonl"):
    try: d = json.loads(line)
    except Exception: continue
    r = d["roofline"]
    print(f'{d["config"]["name"]:28s} {d["dtype"][:3]} eps={d["config"]["perturbation_eps"]:<7} store={d["config"]["store_every"]:<3} '
          f'{d["value"]:9.0f} MLUPS  {d["ms_per_step"]:8.3f} ms  kernel {r["kernel_ms"]:.3f} ms  frac {r["frac"]:.3f}  clocks {d["clocks"]["sm_mhz"]} {d["clocks"]["reasons"]}')
PY
