"""bench.py's command line and wall-clock budget (the parts that need no GPU)."""
import importlib
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_defaults_and_flags():
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--help"], capture_output=True, text=True)
    assert p.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--config"):
        assert flag in p.stdout
    sys.path.insert(0, str(ROOT))
    bench = importlib.import_module("bench")
    # the defaults the driver relies on: one GPU, at least three warm-up steps, the C2 workload
    src = (ROOT / "bench.py").read_text()
    assert 'add_argument("--gpus", type=int, default=1)' in src
    assert 'add_argument("--warmup", type=int, default=3)' in src
    assert 'add_argument("--config", default="c2"' in src
    assert {"c2", "c4"} <= set(bench.WORKLOADS)


def test_wall_clock_budget(monkeypatch):
    sys.path.insert(0, str(ROOT))
    bench = importlib.import_module("bench")
    monkeypatch.setattr(bench, "BENCH_BUDGET_S", 100.0)
    monkeypatch.setattr(bench, "T_PROCESS_START", bench.time.perf_counter() - 30.0)
    assert 69.0 < bench.time_left() <= 70.0
    monkeypatch.setattr(bench, "T_PROCESS_START", bench.time.perf_counter() - 130.0)
    assert bench.time_left() < 0          # spent: the optional legs yield (bench.py: time_left() > ... tests)
