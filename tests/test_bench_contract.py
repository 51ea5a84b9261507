"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with
the agreed keys, and non-zero ranks of a multi-rank reference launch stay silent."""

from __future__ import annotations

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*extra, env=None):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--rows", "20000", *extra]
    return subprocess.run(cmd, capture_output=True, text=True, env={**os.environ, **(env or {})}, timeout=300)


def test_reference_arm_prints_one_json_line_with_contract_keys():
    proc = run_bench("--gpus", "1")
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, proc.stdout
    out = json.loads(lines[0])
    assert out["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in out, key
    assert out["unit"] == "queries/s" and out["higher_is_better"] is True and out["vs_baseline"] is None
    assert out["steps"] == 2 and out["warmup"] == 1 and out["value"] > 0
    assert out["config"]["workload"].startswith("10M x 768") and out["config"]["rows"] == 20_000
    assert out["cpu_baseline"]["kind"] in ("reference", "port") and out["cpu_baseline"]["cores"] >= 1
    assert "FULL corpus 20000 x 768" in out["cpu_baseline"]["sample"]
    # the SAME metric string as the repo arm prints, so that the driver can divide one by the other
    sys.path.insert(0, ROOT)
    import bench

    w = dict(bench.WORKLOADS["c3"], rows=20_000)
    assert out["metric"] == bench.metric_string(w)
    assert out["config"] == {**bench.workload_config(dict(w, desc=out["config"]["workload"]), 1)}
    assert out["cpu_baseline"]["value"] == out["value"]
    assert abs(out["ms_per_step"] - 256 * 1e3 / out["value"]) < 1e-6 * out["ms_per_step"]
    assert out["e2e"] == {"value": out["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    proc = run_bench("--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert proc.returncode == 0 and proc.stdout.strip() == ""


def test_workload_table_matches_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench

    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        configs = json.load(f)["configs"]
    assert len(configs) == 5
    w = bench.WORKLOADS
    assert (w["c1"]["rows"], w["c1"]["dim"], w["c1"]["batch"], w["c1"]["k"]) == (10_000, 384, 1, 10)
    assert (w["c2"]["rows"], w["c2"]["dim"], w["c2"]["batch"], w["c2"]["k"], w["c2"]["storage"]) == (1_000_000, 768, 64, 32, "bfloat16")
    assert (w["c3"]["rows"], w["c3"]["dim"], w["c3"]["batch"], w["c3"]["k"], w["c3"]["storage"]) == (10_000_000, 768, 256, 100, "bfloat16")
    assert (w["c4"]["rows"], w["c4"]["dim"], w["c4"]["batch"], w["c4"]["k"], w["c4"]["storage"]) == (10_000_000, 1536, 1024, 100, "float16")
    assert (w["c5"]["rows"], w["c5"]["dim"], w["c5"]["batch"], w["c5"]["k"]) == (50_000, 384, 1000, 5)
    # algorithmic bytes of the bench workload: corpus once + queries + hits
    assert bench.algorithmic_bytes(10_000_000, 768, "bfloat16", 256, 100) == 15_361_093_632
