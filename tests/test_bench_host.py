"""Host logic of bench.py that must never lose the headline JSON line: the runner of the secondary workloads
(`run_secondary`) with fake workloads -- success, an exception, a workload that never returns (watchdog), several
"ranks" as processes over gloo -- and the static shape of the workload tables."""
from __future__ import annotations

import json
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402


def _run(script: str, timeout: int = 60, env: dict | None = None) -> subprocess.CompletedProcess:
    full = "import sys\nsys.path.insert(0, %r)\nimport bench\n" % str(ROOT) + textwrap.dedent(script)
    return subprocess.run([sys.executable, "-c", full], capture_output=True, text=True, timeout=timeout,
                          env={**os.environ, **(env or {})})


def _only_line(stdout: str) -> dict:
    lines = [l for l in stdout.splitlines() if l.strip()]
    assert len(lines) == 1, stdout
    return json.loads(lines[0])


def test_no_entries_prints_the_line_unchanged():
    proc = _run("""
        bench.run_secondary([], None, {"value": 1.5}, 0, 1, max, sum, 5)
    """)
    assert proc.returncode == 0, proc.stderr
    assert _only_line(proc.stdout) == {"value": 1.5}


def test_results_and_errors_land_under_also():
    proc = _run("""
        def measure(entry):
            if entry[0] == "bad":
                raise RuntimeError("metalbm_b200 error -5: out of memory")
            return {"name": entry[0], "value": 2.0}
        entries = [("good", "F64", None, 1, 3), ("bad", "F32", 0.02, 1, 3), ("good2", "F64", None, 1, 3)]
        bench.run_secondary(entries, measure, {"value": 1.5}, 0, 1, lambda v: v, lambda v: v, 30)
        print("after", file=sys.stderr)
    """)
    assert proc.returncode == 0, proc.stderr
    line = _only_line(proc.stdout)
    assert line["value"] == 1.5
    assert [e["name"] for e in line["also"]] == ["good", "bad", "good2"]
    assert "out of memory" in line["also"][1]["error"] and line["also"][1]["dtype"] == "f32"
    assert "also_note" not in line
    assert "after" in proc.stderr   # the caller goes on (destroy_process_group etc.)


def test_a_workload_that_never_returns_is_cut_off_and_the_line_is_printed_once():
    proc = _run("""
        import time
        def measure(entry):
            if entry[0] == "hang":
                time.sleep(3600)
            return {"name": entry[0], "value": 2.0}
        entries = [("first", "F64", None, 1, 3), ("hang", "F64", None, 1, 3), ("never", "F64", None, 1, 3)]
        bench.run_secondary(entries, measure, {"value": 1.5}, 0, 1, lambda v: v, lambda v: v, 3)
    """, timeout=30)
    assert proc.returncode == 0, proc.stderr
    line = _only_line(proc.stdout)
    assert line["value"] == 1.5 and [e["name"] for e in line["also"]] == ["first"]
    assert "watchdog" in line["also_note"]


def test_other_ranks_print_nothing():
    proc = _run("""
        bench.run_secondary([("a", "F64", None, 1, 3)], lambda e: {"name": e[0]}, None, 1, 1, lambda v: v, lambda v: v, 10)
    """)
    assert proc.returncode == 0, proc.stderr
    assert proc.stdout.strip() == ""


def test_time_budget_skips_the_rest():
    proc = _run("""
        import time
        def measure(entry):
            time.sleep(2.5)
            return {"name": entry[0]}
        entries = [("a", "F64", None, 1, 3), ("b", "F64", None, 1, 3), ("c", "F64", None, 1, 3)]
        bench.run_secondary(entries, measure, {"value": 1.0}, 0, 1, lambda v: v, lambda v: v, 3)
    """, timeout=30)
    assert proc.returncode == 0, proc.stderr
    line = _only_line(proc.stdout)
    assert line["also"][0] == {"name": "a"}
    assert all("skipped" in e for e in line["also"][1:]) and len(line["also"]) == 3


_TWO_RANKS = """
    import os, time
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group(backend="gloo")
    def reduce(op):
        def f(v):
            t = torch.tensor([float(v)], dtype=torch.float64)
            dist.all_reduce(t, op=op)
            return float(t.item())
        return f
    MODE = os.environ["MODE"]
    def measure(entry):
        if entry[0] == "symmetric-failure":
            raise RuntimeError("no memory anywhere")
        if entry[0] == "rank1-dies" and rank == 1:
            os._exit(0)
        t = torch.tensor([1.0], dtype=torch.float64)
        dist.all_reduce(t)                      # the workload's own collective
        if entry[0] == "one-rank-fails" and rank == 1:
            raise RuntimeError("only rank 1")   # after the collective: the other rank is not left waiting
        return {"name": entry[0], "ranks": float(t.item())}
    entries = [(n, "F64", None, 1, 3) for n in MODE.split(",")]
    bench.run_secondary(entries, measure, {"value": 1.5} if rank == 0 else None, rank, world,
                        reduce(dist.ReduceOp.MAX), reduce(dist.ReduceOp.SUM), int(os.environ.get("TIMEOUT", "20")))
    dist.barrier()
    dist.destroy_process_group()
"""


def _torchrun(mode: str, timeout_s: int, port: int, tmp_path: Path) -> subprocess.CompletedProcess:
    script = "import sys\nsys.path.insert(0, %r)\nimport bench\n" % str(ROOT) + textwrap.dedent(_TWO_RANKS)
    path = tmp_path / "bench_two_ranks.py"
    path.write_text(script)
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", str(port), str(path)],
                          capture_output=True, text=True, timeout=180,
                          env={**os.environ, "MODE": mode, "TIMEOUT": str(timeout_s), "OMP_NUM_THREADS": "1"})


def test_two_ranks_stay_in_step_after_symmetric_and_one_sided_failures(tmp_path):
    proc = _torchrun("ok,symmetric-failure,one-rank-fails,ok", 60, 29611, tmp_path)
    assert proc.returncode == 0, proc.stderr[-3000:]
    line = _only_line(proc.stdout)
    also = line["also"]
    assert [e.get("name") for e in also] == ["ok", "symmetric-failure", "one-rank-fails", "ok"]
    assert also[0]["ranks"] == 2.0 and also[3]["ranks"] == 2.0
    assert "no memory anywhere" in also[1]["error"] and also[1]["ranks_ok"] == 0
    assert also[2]["ranks_ok"] == 1 and also[2]["ranks"] == 2.0   # rank 0's own measurement stands, flagged
    assert "also_note" not in line


@pytest.mark.timeout(240)
def test_two_ranks_a_lost_rank_ends_in_the_watchdog_with_exit_code_zero(tmp_path):
    proc = _torchrun("ok,rank1-dies,ok", 8, 29612, tmp_path)
    # rank 0 is stuck in the workload's collective (or sees it fail): either way the line is out, once, and torchrun is happy
    assert proc.returncode == 0, proc.stderr[-3000:]
    line = _only_line(proc.stdout)
    assert line["value"] == 1.5 and line["also"][0]["name"] == "ok"


def test_secondary_tables_name_known_workloads():
    for table in (bench.ALSO_SINGLE, bench.ALSO_MULTI):
        for name, dtype, eps, mode, steps in table:
            assert name in bench.WORKLOADS and dtype in ("F64", "F32") and mode in (1, 2) and steps >= 10
            work = bench.WORKLOADS[name]
            assert (eps is None) or work["collision"] != "BGK"
    # the 1024^3 box does not fit one GPU: only in the multi-GPU table (the N > 1 HEADLINE is that box with the reductions
    # alone where the fields do not fit; the table carries the full configuration, and the weak-scaled cube of round 1)
    assert all(name != "d3q19_bgk_1024" for name, *_ in bench.ALSO_SINGLE)
    assert [e[3] for e in bench.ALSO_MULTI if e[0] == "d3q19_bgk_1024"] == [1]
    assert bench.ALSO_MULTI[0][0] == "d3q19_bgk_256"


class _FakeAlgorithm:
    """Stands in for metalbm_b200.algorithm.Algorithm: checks the calls measure_also makes, returns plausible numbers."""
    created = []

    def __init__(self, cfg, communication=None, host_distribution=True, peer_halos=None, host_fields=True):
        assert host_distribution is False and host_fields is False   # slabs that fill the GPU: no host arrays
        self.cfg, self.peer_halos, self.calls, self.closed = cfg, bool(peer_halos), [], False
        _FakeAlgorithm.created.append(self)

    def init_synthetic(self, a, b): self.calls.append(("init", a, b))
    def perturb(self, eps): self.calls.append(("perturb", eps))
    def run(self, first, count, store_every=0, sync=True, stored_mode=1): self.calls.append(("run", first, count, store_every, stored_mode))
    def kernel_time(self): return 2.0, 7
    def mark(self, slot): pass
    def synchronize(self): pass
    def elapsed_ms(self, a, b): return 50.0
    def observables(self):
        nodes = 1
        for n in self.cfg.global_length: nodes *= n
        return [1e-3, float("nan"), 0.1, float(nodes)]
    def alpha_statistics(self): return 0.5, 1.2, 2.1
    def newton_statistics(self, start=False): return (0, 0) if start else (10, 30)
    def close(self): self.closed = True


@pytest.mark.parametrize("world", [1, 2, 8])
def test_measure_also_bookkeeping_with_a_fake_context(monkeypatch, world):
    import types
    import torch
    import metalbm_b200.algorithm as A
    monkeypatch.setattr(A, "Algorithm", _FakeAlgorithm)
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda *a: (190 * 10**9, 192 * 10**9))
    args = types.SimpleNamespace(overlap="On", halo="peer", variant=0)
    identity = lambda v: v  # noqa: E731
    for entry in (bench.ALSO_SINGLE if world == 1 else bench.ALSO_MULTI):
        _FakeAlgorithm.created.clear()
        out = bench.measure_also(entry, args, 0, world, 0, lambda: None, identity, identity, 6500.0)
        name, dtype, eps, mode, steps = entry
        work = bench.WORKLOADS[name]
        shape = out["global_length"]
        nodes = shape[0] * shape[1] * shape[2]
        if "skipped" in out:   # 1024^3 does not fit ... only the memory rule may skip
            assert "GB" in out["skipped"] and name == "d3q19_bgk_1024" and world == 2 and mode == 1
            continue
        assert out["value"] == pytest.approx(nodes * steps / 0.050 / 1e6)
        assert out["mass_per_node"] == pytest.approx(1.0)
        element = 8 if dtype == "F64" else 4
        per_node = 2 * work["q"] * element + (2 * element if work["collision"] != "BGK" else 0)
        assert out["roofline_mlups_per_gpu"] == pytest.approx(6500e9 / per_node / 1e6)
        lx = shape[0] // world
        kernel_nodes = nodes // world if world == 1 else nodes // world // lx * (lx - 2)
        assert out["roofline_frac"] == pytest.approx(per_node * kernel_nodes / 2.0e-3 / 1e9 / 6500.0)
        fake, = _FakeAlgorithm.created
        assert fake.closed and fake.calls[0][0] == "init"
        assert (("perturb", eps) in fake.calls) == (eps is not None)
        timed = [c for c in fake.calls if c[0] == "run" and c[2] == steps]
        assert len(timed) == 1 and timed[0][3] == work["store_every"] and timed[0][4] == mode
        assert fake.calls[-1] == ("run", 0, 1, 1, 2)   # the liveness step reduces energy / mass / Mach only
        if world > 1:
            assert out["halo"] == "peer" and tuple(fake.cfg.global_length)[0] == shape[0]


class _FakeLib:
    def mlbm_step(self, ctx, iteration, stored):
        return 0

    def mlbm_run_async(self, ctx, first, count, store_every):
        return 0


class _FakeBenchAlgorithm(_FakeAlgorithm):
    """The part of Algorithm that bench.run_ours uses around the headline measurement."""

    def __init__(self, cfg, communication=None, host_distribution=True, peer_halos=None, host_fields=True):
        from metalbm_b200.algorithm import Distribution, Domain, FieldList
        self.cfg, self.peer_halos, self.calls, self.closed = cfg, bool(peer_halos), [], False
        self.domain = Domain(cfg)
        self.fieldList = FieldList(self.domain, allocate=host_fields)
        self.distribution = Distribution(self.domain, allocate=host_distribution)
        self._lib, self._ctx = _FakeLib(), None
        _FakeAlgorithm.created.append(self)

    def init_equilibrium(self): self.calls.append(("init_equilibrium",))
    def launch_count(self): return 7 * len(self.calls)
    def pack(self): self.calls.append(("pack",))
    def unpack(self): self.calls.append(("unpack",))


def test_run_ours_prints_one_complete_line_with_a_fake_context(monkeypatch, capsys):
    """The whole of bench.run_ours on the CPU with the device faked away: every key of the contract is there, the secondary
    workloads follow, exactly one line is printed."""
    import types
    import torch
    import metalbm_b200.algorithm as A
    import metalbm_b200.capi as capi
    monkeypatch.setattr(A, "Algorithm", _FakeBenchAlgorithm)
    monkeypatch.setattr(capi, "check", lambda status: None)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda *a: (190 * 10**9, 192 * 10**9))
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{key: v for key, v in k.items() if key not in ("pin_memory", "device")}))
    monkeypatch.setattr(bench, "cpu_baseline_leg", lambda: {"value": 50.0, "unit": "MLUPS", "cores": 4, "kind": "reference", "sample": "fake"})
    monkeypatch.setattr(bench, "EDGE", 8)
    monkeypatch.setitem(bench.WORKLOADS["d3q19_bgk_256"], "shape", (8, 8, 8))
    args = types.SimpleNamespace(gpus=1, steps=5, warmup=3, impl="ours", edge=8, variant=0, workload="d3q19_bgk_256", dtype="f64",
                                 overlap="On", halo="peer", eps=None, store_every=None, no_e2e=False, no_cpu_baseline=False,
                                 also="auto", also_timeout=120)
    _FakeAlgorithm.created.clear()
    assert bench.run_ours(args) == 0
    out = capsys.readouterr().out
    line = _only_line(out)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "also"):
        assert key in line, key
    assert line["metric"] == "MLUPS (D3Q19, FP64)" and line["unit"] == "MLUPS" and line["n_gpus"] == 1 and line["steps"] == 5
    assert line["value"] == pytest.approx(8 ** 3 * 5 / 0.050 / 1e6)
    assert line["roofline"]["bound"] == "hbm" and line["roofline"]["algorithmic_bytes_per_node"] == 304
    assert line["e2e"]["h2d_bytes_per_step"] == pytest.approx(19 * 8 ** 3 * 8 / 5)
    assert line["cpu_baseline"]["kind"] == "reference" and line["vs_baseline"] is None and line["dtype"] == "f64"
    assert [e["name"] for e in line["also"]] == [e[0] for e in bench.ALSO_SINGLE]
    assert all("value" in e or "skipped" in e for e in line["also"])
    assert all(a.closed for a in _FakeAlgorithm.created)
