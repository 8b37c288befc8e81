"""bench.py's reference arm runs on CPU (the reference's own code from oracle/_ref): check the JSON line it
prints against the driver's contract (keys, metric, cpu_baseline, e2e) on a tiny bounded sample."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "arcs_ref")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs /root/reference once)")
def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-genome", "300000", "--cpu-pairs-per-thread", "300"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=300, env=dict(os.environ, RANK="0"))
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "k-mers/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("read k-mers/s") and d["value"] > 0 and d["vs_baseline"] is None
    for key in ("n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert p.returncode == 0 and p.stdout.strip() == ""
