#!/usr/bin/env python
"""Flat checkpoint container (mlbm_checkpoint_write) <-> the reference's HDF5 checkpoint, where h5py is installed.

    python tools/checkpoint_to_hdf5.py to-h5   distribution-100.mlbm distribution-100.h5
    python tools/checkpoint_to_hdf5.py from-h5 distribution-100.h5   distribution-100.mlbm <dimD> <dimQ> <Lx> <Ly> <Lz> [iteration]

The container holds exactly what DistributionWriter::writeDistribution puts into the .h5 (Writer.h:400-445): data sets
"distribution<iQ>", iQ = 0 .. dimQ - 1, each the padded global box gSD::pLength() projected on the lattice's dimD dimensions,
H5T_NATIVE_DOUBLE.  h5py is not part of the build image (this script is not exercised by the test-suite there); the container
side is (tests/test_checkpoint_gpu.py)."""
import json
import sys

import numpy as np

HEADER_BYTES = 4096


def read_container(path):
    with open(path, "rb") as handle:
        header = json.loads(handle.read(HEADER_BYTES).decode())
        box = header["padded_global_length"]
        data = np.fromfile(handle, dtype=np.float64).reshape([header["dimQ"]] + box)
    return header, data


def write_container(path, data, dim, global_length, iteration=0):
    q = data.shape[0]
    header = {"format": "metalbm_b200 checkpoint 1", "datasets": "distribution<iQ>, iQ = 0 .. dimQ - 1 (Writer.h:400-445)",
              "dtype": "float64", "dimD": dim, "dimQ": q, "global_length": list(global_length),
              "padded_global_length": list(data.shape[1:]), "header_bytes": HEADER_BYTES, "iteration": iteration, "written_by_ranks": 1}
    text = json.dumps(header).encode()
    with open(path, "wb") as handle:
        handle.write(text + b" " * (HEADER_BYTES - 1 - len(text)) + b"\n")
        np.ascontiguousarray(data, dtype=np.float64).tofile(handle)


def main():
    import h5py
    if sys.argv[1] == "to-h5":
        header, data = read_container(sys.argv[2])
        with h5py.File(sys.argv[3], "w") as out:
            for iq in range(header["dimQ"]):
                out.create_dataset(f"distribution{iq}", data=data[iq].reshape(header["padded_global_length"][:header["dimD"]]))
    else:
        dim, q, lx, ly, lz = map(int, sys.argv[4:9])
        iteration = int(sys.argv[9]) if len(sys.argv) > 9 else 0
        with h5py.File(sys.argv[2], "r") as source:
            sets = [np.asarray(source[f"distribution{iq}"], dtype=np.float64) for iq in range(q)]
        box = list(sets[0].shape) + [1] * (3 - dim)
        write_container(sys.argv[3], np.stack(sets).reshape([q] + box), dim, (lx, ly, lz), iteration)


if __name__ == "__main__":
    main()
