"""The driver's bench contract, as far as it can be checked without a GPU: the reference arm
(`bench.py --impl reference`, the CPU oracle on a bounded sample) prints ONE JSON line carrying every key the
contract names, with the same metric/unit/config as the GPU arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-rows", "20000", "--cpu-nq", "64"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference"
    assert j["metric"] == "queries/sec at 10M x 768 k=10" and j["unit"] == "queries/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 1 and j["steps"] == 1
    assert j["value"] > 0 and j["ms_per_step"] > 0 and j["vs_baseline"] is None
    assert "BASELINE.json configs[1]" in j["config"]["workload"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "sample" in cb
    e2e = j["e2e"]
    assert e2e["value"] == j["value"] and e2e["unit"] == j["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
