"""bench.py's reference arm runs without a GPU (it times the reference's CPU path: the extraction through the oracle port and
env_shade through the reference's own kernel compiled for the CPU), so its JSON line can be checked here against the driver
contract.  The arm measures one FULL step of whatever configuration it is given (no extrapolation), so the test hands it a
small one ('64' grid, one 64^2 view, n = 4)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "0", "--grid", "64", "--views", "1", "--res", "64", "--n-samples", "4"], capture_output=True,
                         text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_iters_per_sec" and d["unit"] == "iters/s"
    assert d["higher_is_better"] is True and d["scaling"] == "strong" and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["steps"] >= 1 and d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-3 * 1e3
    assert "workload" in d["config"] and "'64' grid" in d["config"]["workload"] and "model" not in d["config"]
    assert d["steps"] == 1 and set(d["parts_s"]) == {"extraction", "env_shade"} and d["parts_s"]["extraction"] > 0
    cb = d["cpu_baseline"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(cb) and cb["kind"] in ("reference", "port")
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and "extraction" in cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0 and not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
