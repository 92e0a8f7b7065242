"""bench.py's contract that needs no GPU: the reference arm (`--impl reference`: the oracle port in the reference's
loop structure on the host cores) prints ONE JSON line with the keys the driver reads, on the same `config`,
`metric` and `unit` as the GPU arm's workload table; under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.strip().splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run(["--impl", "reference", "--workload", "c1", "--steps", "2", "--warmup", "1"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "tracks/s"
    assert d["metric"].startswith("tracks/sec") and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("64 tracks x 64 dets") and d["config"]["tracks"] == 64
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    lines = _run(["--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0", "--gpus", "2"],
                 env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert lines == []
