"""Collected last (file name): a GPU box must have RUN every multi-rank parity case it has the GPUs for.

The multi-rank cases (tests/test_multi_gpu.py, the `on_slabs` cases of the spectral / wide-lattice suites, the two-rank
template-layer cases, the multi-rank bench support case) skip themselves with "needs N GPUs" on a smaller box.  That is
right on a one-GPU box and wrong anywhere else: this test fails -- it does not skip -- when the box had N or more GPUs and
such a case was skipped, and when a box with at least two GPUs ran none of them."""
import pytest


@pytest.mark.gpu
def test_multi_rank_cases_ran_wherever_the_box_has_the_gpus(request):
    import os
    import torch
    from conftest import GPU_OUTCOMES
    if os.environ.get("PYTEST_XDIST_WORKER"):
        pytest.skip("the outcomes of the other xdist workers are not visible here: the guard needs a serial run (the driver's)")
    devices = torch.cuda.device_count()
    wrongly_skipped = [(node, need) for node, need in GPU_OUTCOMES["skipped_needing"] if need <= devices]
    assert not wrongly_skipped, f"{devices} GPUs present, yet skipped: {wrongly_skipped[:5]}"
    selected = [item.nodeid for item in request.session.items]
    multi_selected = [node for node in selected if "test_multi_gpu.py" in node or "on_slabs" in node]
    if devices >= 2 and multi_selected:
        ran = set(GPU_OUTCOMES["ran"])
        assert any(node in ran for node in multi_selected), "a multi-GPU box ran no multi-rank parity case"
