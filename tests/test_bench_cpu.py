"""bench.py's reference arm on the CPU (the GPU arm needs a B200): one JSON line with the keys the driver reads.  The sample is the
bench's own (1280x720 view, every 8th block of 10 rows) at 1 spp per step instead of 16 (B200PT_BENCH_SPP, a tuning aid of bench.py)."""
import json
import os
import subprocess
import sys

import helpers


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, B200PT_BENCH_SPP="1")
    p = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--no-em"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("cornell-dielectric 1280x720")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == d["value"] and cb["cores"] >= 1 and "rows" in cb["sample"]
    # the reference's own shader source when oracle/_ref holds it (this container, and the GPU box: the library travels)
    assert (cb["kind"] == "reference") == os.path.exists(os.path.join(helpers.ROOT, "oracle", "_ref", "libshader_ref.so"))
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_other_ranks_of_the_reference_arm_exit_without_work():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""
