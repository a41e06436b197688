"""CPU: bench.py's reference arm runs without a GPU and prints the contract's JSON line."""

import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_json_line():
    """`--impl reference` times the reference itself (baseline/_ref, installed by baseline/install_ref.sh) on the host
    cores; without the install it falls back to the oracle port and says so in `kind`."""
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--workload", "cfg5:64", "--cpu-batch", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    have_ref = (ROOT / "baseline" / "_ref" / "llm_quest").is_dir()
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config"):
        assert key in line


def test_workloads_cover_baseline_configs():
    sys.path.insert(0, str(ROOT))
    import bench

    for name, imgs in (("cfg1", 8), ("cfg2", 64), ("cfg3", 128), ("cfg4", 128), ("cfg5:224", 256), ("cfg5:1344", 8)):
        wl = bench.make_workload(name)
        assert wl.images() == imgs, (name, wl.images())
        assert wl.flops() > 0 and name.split(":")[0] in wl.describe(wl.B)
    ids = bench.vlm_input_ids(2, 4, 196, 2048, __import__("torch").Generator().manual_seed(0))
    assert ids.shape == (2, 2832) and int((ids == bench.IMG_TOKEN).sum()) == 2 * 4 * 196


def test_flop_model_matches_baseline_md():
    sys.path.insert(0, str(ROOT))
    import bench

    assert abs(bench.tower_flops(784) / 1e9 - 162.6) < 0.1      # BASELINE.md §3, cfg-2 per image
    assert abs(bench.tower_flops(196) / 1e9 - 36.4) < 0.1
    assert abs(bench.tower_flops(6272) / 1e12 - 2.570) < 0.005  # cfg-4 per sample
