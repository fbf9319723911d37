"""CPU test (no GPU): the reference arm of bench.py prints ONE JSON line with the contract's keys, and the
B200 arm's helpers build the same `config` for both arms."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, check=True, cwd=ROOT, timeout=600)
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Msps" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "2^30" in d["config"]["workload"]


def test_both_arms_share_the_config():
    sys.path.insert(0, ROOT)
    import bench

    a = bench.workload_config(30, 9)
    assert a == bench.workload_config(30, 9, 20.0, 9.5) and a["samples_per_gpu"] == 1 << 30
    assert "configs[1]" in a["workload"] and "configs[3]" in bench.workload_config(30, 33, 0.0)["workload"]
    assert abs(bench.flop_per_sample(9) - 739) < 1 and abs(bench.flop_per_sample(1) - 141) < 1
