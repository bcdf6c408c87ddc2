#!/usr/bin/env python
"""Device-copy bandwidth (read + write bytes) against the FOOTPRINT of the copy.

MEASURED_PEAKS.json's hbm_gbs is a copy of 2 GB into 2 GB.  The lattice slabs of the BASELINE configs are 5-80 GB: this
prints what the same torch copy achieves when source and target are that large, which is the fair HBM roof for them.
    python tools/copy_bw.py [out.json]
"""
import json
import sys

import torch


def copy_bandwidth(gigabytes_per_buffer: float) -> float:
    count = int(gigabytes_per_buffer * 1e9) // 8
    a = torch.empty(count, dtype=torch.float64, device="cuda")
    b = torch.empty(count, dtype=torch.float64, device="cuda")
    a.fill_(1.0)
    b.copy_(a)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(5):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        b.copy_(a)
        stop.record()
        torch.cuda.synchronize()
        best = max(best, 2 * count * 8 / (start.elapsed_time(stop) * 1e-3) / 1e9)
    del a, b
    torch.cuda.empty_cache()
    return best


def main():
    rows = []
    for gigabytes in (1.0, 2.0, 2.55, 10.0, 20.4, 41.0, 80.0):
        try:
            rows.append({"GB_per_buffer": gigabytes, "footprint_GB": 2 * gigabytes, "copy_GBps": copy_bandwidth(gigabytes)})
        except Exception as error:  # noqa: BLE001
            rows.append({"GB_per_buffer": gigabytes, "error": str(error)[:100]})
        print(rows[-1], flush=True)
    if len(sys.argv) > 1:
        json.dump({"how": "torch b.copy_(a) of float64 buffers, CUDA events, best of 5", "rows": rows}, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
