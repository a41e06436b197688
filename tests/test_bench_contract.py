"""CPU: bench.py's reference arm runs without a GPU and prints the contract's JSON line."""

import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--px", "64", "--cpu-batch", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config"):
        assert key in line


def test_flop_model_matches_baseline_md():
    sys.path.insert(0, str(ROOT))
    import bench

    assert abs(bench.tower_flops(784) / 1e9 - 162.6) < 0.1      # BASELINE.md §3, cfg-2 per image
    assert abs(bench.tower_flops(196) / 1e9 - 36.4) < 0.1
    assert abs(bench.tower_flops(6272) / 1e12 - 2.570) < 0.005  # cfg-4 per sample
