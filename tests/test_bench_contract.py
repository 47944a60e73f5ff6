"""bench.py contract pieces that run without a GPU: the reference arm's JSON line and the workload cache."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line(tmp_path):
    env = dict(os.environ, REYNA_B200_CACHE=str(tmp_path))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--reference-sample", "1024"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "assembled DoF/s" and j["unit"] == "DoF/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["steps"] == 1 and j["warmup"] == 1
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"] == {"value": j["value"], "unit": "DoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"]


def test_reference_arm_is_silent_on_other_ranks(tmp_path):
    env = dict(os.environ, REYNA_B200_CACHE=str(tmp_path), RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_workload_cache_roundtrip(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, "CACHE_DIR", str(tmp_path))
    g1, path, cached1 = bench.workload_geometry(256)
    g2, _, cached2 = bench.workload_geometry(256)
    assert not cached1 and cached2 and os.path.exists(path)
    assert g1.n_elements == g2.n_elements == 256
    import numpy as np
    assert np.array_equal(g1.interior_edges, g2.interior_edges) and np.array_equal(g1.nodes, g2.nodes)
