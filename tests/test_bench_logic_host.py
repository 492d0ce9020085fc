"""bench.py's default arm, executed end to end on the CPU with a stubbed CUDA surface (tests/tools/bench_mock_run.py):
the JSON line carries every key of the measurement contract.  Guards the benchmark's control flow against edits made
without a GPU at hand; the numbers themselves are meaningless here."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_default_bench_line_has_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "bench_mock_run.py")], capture_output=True,
                         text=True, timeout=600)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stderr[-2000:]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["unit"] == "tokens/s" and d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] > 0
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and d["roofline"]["bound"] == "tensor"
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert d["roofline_transform"]["bound"] == "hbm" and len(d["fused_linear_per_shape"]) == 4
    assert d["gpu_launches"] == 2 * 4 * 2 * d["steps"]   # (transform + GEMM) x 4 linears x 2 mock layers per step
    assert "workload" in d["config"] and "model" not in d["config"]
