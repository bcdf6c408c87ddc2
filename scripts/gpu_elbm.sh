#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
: > gpurun_out/elbm.jsonl
run() { echo "== $*" >> gpurun_out/elbm.err; timeout 600 python bench.py --no-cpu-baseline --also off --no-e2e "$@" >> gpurun_out/elbm.jsonl 2>> gpurun_out/elbm.err; }
run --workload d3q27_elbm_512 --steps 20
run --workload d3q27_elbm_512 --steps 20 --eps 2e-3
run --workload d3q27_elbm_512 --steps 20 --eps 1e-5
run --workload d2q9_elbm_shanchen_8192 --steps 50
run --workload d2q9_elbm_shanchen_8192 --steps 50 --eps 2e-3
run --workload d2q9_elbm_edm_8192 --steps 50
run --workload d2q9_elbm_shanchen_8192 --steps 50 --dtype f32
python - <<'PY'
import json
for line in open("gpurun_out/elbm.jsonl"):
    try: d = json.loads(line)
    except Exception: continue
    r = d["roofline"]
    print(f'{d["config"]["name"]:28s} {d["dtype"][:3]} eps={d["config"]["perturbation_eps"]:<7} '
          f'{d["value"]:9.0f} MLUPS  {d["ms_per_step"]:8.3f} ms  frac {r["frac"]:.3f}  clocks {d["clocks"]["sm_mhz"]} {d["clocks"]["reasons"]}')
PY
