#!/bin/bash
# quick ELBM throughput check (no parity suite): D3Q27 512^3 and D2Q9 8192^2, dense and sparse Newton regimes
set -u
mkdir -p gpurun_out
: > gpurun_out/elbm_quick.jsonl
run() { timeout 600 python bench.py --no-cpu-baseline --also off --no-e2e "$@" >> gpurun_out/elbm_quick.jsonl 2>> gpurun_out/elbm_quick.err; }
run --workload d3q27_elbm_512 --steps 20
run --workload d3q27_elbm_512 --steps 20 --eps 1e-5
run --workload d2q9_elbm_shanchen_8192 --steps 50
run --workload d2q9_elbm_shanchen_8192 --steps 50 --eps 1e-5
python - <<'PY'
import json
for line in open("gpurun_out/elbm_quick.jsonl"):
    try: d = json.loads(line)
    except Exception: continue
    r = d["roofline"]
    print(f'{d["config"]["name"]:28s} {d["dtype"][:3]} eps={d["config"]["perturbation_eps"]:<7} '
          f'{d["value"]:9.0f} MLUPS  {d["ms_per_step"]:8.3f} ms  frac {r["frac"]:.3f}')
PY
tail -3 gpurun_out/elbm_quick.err
