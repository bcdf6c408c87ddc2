"""the checkpoint of the distribution in the reference's data-set layout (DistributionWriter::writeDistribution,
Writer.h:400-445; DistributionReader::readDistribution, Reader.h:119-157): dimQ data sets "distribution<iQ>" of the PADDED
GLOBAL box (gSD::pLength(), doubles), here in a flat container (mlbm_checkpoint_write / _read).  The file's bytes are checked
against the layout built with numpy, and a restart from it continues bit-identically."""
import json

import numpy as np
import pytest

from metalbm_b200.algorithm import Algorithm
from metalbm_b200.capi import make_config
from oracle import oracle as O

pytestmark = pytest.mark.gpu

CASES = [("D3Q19", (8, 6, 5), "BGK", "F64"), ("D2Q9", (10, 7, 1), "ELBM", "F64"), ("D3Q27", (6, 4, 6), "BGK", "F32")]


@pytest.mark.parametrize("lattice,shape,collision,dtype", CASES)
def test_container_layout_and_bit_identical_restart(tmp_path, lattice, shape, collision, dtype):
    cfg = make_config(lattice=lattice, shape=shape, collision=collision, forcing_scheme="Guo", force="Kolmogorov", tau=0.6,
                      amplitude=(1e-4, 2e-4, 3e-4), wavelength=(8.0, 4.0, 16.0), dtype=dtype)
    f0 = O.synthetic_populations(cfg, eps=1e-2)
    path = tmp_path / "distribution-2.mlbm"
    with Algorithm(cfg) as first:
        domain = first.domain
        first.distribution.set_interior(f0.astype(domain.dtype))
        first.unpack()
        first.run(1, 2)
        first.write_checkpoint(path, 2)
        first.pack()
        at_checkpoint = first.distribution.array.copy()          # the local padded array = the hyperslab of rank 0 of 1
        first.run(3, 3)
        first.pack()
        continued = first.distribution.get_interior()

    raw = path.read_bytes()
    header = json.loads(raw[:4096].decode())
    dim, q = domain.dim, domain.q
    padded = list(domain.padded_length)
    assert header["format"] == "metalbm_b200 checkpoint 1" and header["dtype"] == "float64" and header["iteration"] == 2
    assert header["dimD"] == dim and header["dimQ"] == q and header["header_bytes"] == 4096
    assert header["global_length"] == list(shape) and header["padded_global_length"] == padded
    data = np.frombuffer(raw[4096:], dtype=np.float64).reshape([q] + padded)
    # data set iQ = the padded global box, x slowest, the last used dimension padded to 2 (N / 2 + 1) (Domain.h:53-57)
    assert np.array_equal(data, at_checkpoint.astype(np.float64))
    interior = data[:, :shape[0], :shape[1], :shape[2]]
    assert interior.shape == f0.shape and np.isfinite(interior).all()

    with Algorithm(cfg) as second:
        assert second.read_checkpoint(path) == 2
        second.run(3, 3)
        second.pack()
        restarted = second.distribution.get_interior()
        if collision == "BGK":
            assert np.array_equal(restarted, continued)
        else:
            # like the reference's, the checkpoint holds the distribution only: the alpha field restarts at 2 (initAlpha,
            # Initialize.h:82-88), the Newton solve starts elsewhere and stops within its 1e-8 tolerance of the same root
            assert np.abs(restarted - continued).max() <= 1e-8 * np.abs(continued).max()


def test_a_checkpoint_of_another_grid_is_refused(tmp_path):
    cfg = make_config(lattice="D2Q9", shape=(8, 6, 1), collision="BGK", tau=0.6)
    other = make_config(lattice="D2Q9", shape=(8, 8, 1), collision="BGK", tau=0.6)
    path = tmp_path / "a.mlbm"
    with Algorithm(cfg) as algorithm:
        algorithm.distribution.set_interior(O.synthetic_populations(cfg, eps=1e-3))
        algorithm.unpack()
        algorithm.write_checkpoint(path, 0)
    with Algorithm(other) as algorithm:
        with pytest.raises(RuntimeError, match="holds D2Q9 8x6x1"):
            algorithm.read_checkpoint(path)
    with Algorithm(cfg) as algorithm:
        with pytest.raises(RuntimeError, match="cannot open"):
            algorithm.read_checkpoint(tmp_path / "missing.mlbm")
