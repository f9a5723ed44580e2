"""bench.py's reference arm runs on the host cores alone, so its JSON contract is checked here without a GPU:
one line on stdout, the keys the driver reads, and the arm's own meaning of cpu_baseline / e2e."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                        "--warmup", "0", "--ref-envs", "32"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "rodent" in d["config"]["workload"] and d["gpu_launches"] == 0


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_own_arm_line_has_the_contract_keys_in_the_committed_profile():
    """The GPU arm cannot run here; the line it printed on the B200 is committed under profiles/ and must carry every key of the
    contract, the two extra blocks (`strong`, `train`) and a roofline recomputable from its own numbers."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2*_bench.json")))
    assert files, "no round-2 bench line under profiles/"
    d = json.load(open(files[-1]))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["gpu_launches"] == d["steps"] and d["config"]["envs_per_gpu"] == 8192
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert abs(rf["achieved"] - 4776 * 8192 / (rf["avg_launch_ms"] * 1e-3) / 1e9) < 1e-6 * rf["achieved"]     # 4776 B x 8192 envs per launch
    assert d["e2e"]["h2d_bytes_per_step"] == 8192 * 38 * 4 and d["e2e"]["d2h_bytes_per_step"] == 8192 * (617 + 2) * 4
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    if "strong" in d:
        assert d["strong"]["scaling"] == "strong" and d["strong"]["global_envs"] == 8192
        assert d["train"]["scaling"] == "weak" and d["train"]["value"] > 0
