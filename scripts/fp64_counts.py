#!/usr/bin/env python
"""FP64 instructions of a fused kernel by loop, from its SASS (cuobjdump): the static side of the entropic FP64 roofline.

    python scripts/fp64_counts.py <object or .so> <mangled kernel name> [--list]

Prints the loop tree (backward branches), each loop with its instruction and FP64-pipe instruction count (DFMA, DMUL, DADD,
DSETP, DMNMX and the FP64 reciprocal seed MUFU.RCP64H) EXCLUSIVE of its inner loops, and the counts outside any loop but
the plane loop.  bench.py's FP64_OPS table is filled from this output (profiles/r02_sass_fp64_counts.md says how)."""
import re
import subprocess
import sys

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX", "MUFU.RCP64H")


def main():
    obj, kernel = sys.argv[1], sys.argv[2]
    text = subprocess.run(["cuobjdump", "-sass", "-fun", kernel, obj], capture_output=True, text=True, check=True).stdout
    instructions = []
    for line in text.splitlines():
        match = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if match:
            body = match.group(2).strip()
            parts = body.split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            instructions.append((int(match.group(1), 16), op, body))
    index = {address: i for i, (address, _, _) in enumerate(instructions)}
    loops = []
    for i, (address, op, body) in enumerate(instructions):
        if op.startswith("BRA"):
            target = re.search(r"0x([0-9a-f]+)", body)
            if target and int(target.group(1), 16) <= address and int(target.group(1), 16) in index:
                loops.append((index[int(target.group(1), 16)], i))
    loops.sort(key=lambda l: (l[0], -l[1]))
    owner = [None] * len(instructions)
    for number, (start, end) in enumerate(loops):          # later (inner) loops overwrite
        for i in range(start, end + 1):
            if owner[i] is None or (loops[owner[i]][1] - loops[owner[i]][0]) > (end - start):
                owner[i] = number
    def is_fp64(op):
        return any(op.startswith(name) for name in FP64)
    print(f"{len(instructions)} instructions, {sum(is_fp64(op) for _, op, _ in instructions)} FP64")
    outside = [i for i in range(len(instructions)) if owner[i] is None]
    print(f"outside every loop: {len(outside)} instructions, {sum(is_fp64(instructions[i][1]) for i in outside)} FP64")
    for number, (start, end) in enumerate(loops):
        own = [i for i in range(start, end + 1) if owner[i] == number]
        depth = sum(1 for s, e in loops if s <= start and end <= e) - 1
        print(f"{'  ' * depth}loop {number}: [{start}, {end}] {len(own)} own instructions, {sum(is_fp64(instructions[i][1]) for i in own)} FP64")
    if "--list" in sys.argv:
        for i, (address, op, body) in enumerate(instructions):
            print(i, owner[i], body)


if __name__ == "__main__":
    main()
