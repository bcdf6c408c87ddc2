#!/usr/bin/env python
"""Registers / spills / shared memory of every fused-kernel instantiation from the ptxas logs of the last build."""
import re
import sys
from pathlib import Path

import os
build = (Path("/tmp") / ("mlbm_build_" + os.environ["MLBM_VARIANT"])) if os.environ.get("MLBM_VARIANT") else Path(__file__).resolve().parent.parent / "metalbm_b200" / "_build"
names = {"0": "D2Q5", "1": "D2Q9", "2": "D3Q15", "3": "D3Q19", "4": "D3Q27", "5": "D2Q13", "6": "D2Q17", "7": "D2Q21", "8": "D3Q33"}
pattern = sys.argv[1] if len(sys.argv) > 1 else ""
for log in sorted(build.glob("instantiate_*.ptxas.log")):
    text = log.read_text()
    for m in re.finditer(r"Compiling entry function '_ZN4mlbm15fusedStepKernelINS_7LatticeILi(\d)EEELi(\d)ELi(\d)ELi(\d)E([df])EEv\S*' for 'sm_100a'\n"
                         r"ptxas info\s*: Function properties for \S+\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                         r"ptxas info\s*: Used (\d+) registers(.*)", text):
        lattice, collision, eq, scheme, dtype, stack, sst, sld, regs, rest = m.groups()
        label = f"{names[lattice]} { {'0': 'BGK ', '1': 'ELBM', '2': 'ELBF'}[collision] } {'Exact' if eq == '1' else 'Ma3  '} {['None', 'Guo ', 'EDM '][int(scheme)]} {'f64' if dtype == 'd' else 'f32'}"
        if pattern in label:
            print(f"{label}: {regs:>3} regs, stack {stack}, spill st/ld {sst}/{sld}{rest}")
