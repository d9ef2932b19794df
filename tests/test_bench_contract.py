"""bench.py contract checks that need no GPU: the reference arm's JSON line (the driver runs it next to the GPU arm)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_arm(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "60", "--warmup", "1", "--ref-seconds", "8"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and base["metric"].startswith(d["metric"]) and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["steps"] == 60 and d["gpu_launches"] == 0
    assert d["value"] > 0 and abs(d["e2e"]["value"] - d["value"]) < 1e-12 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["ms_per_step"] * d["steps"] < 240e3                                          # the run stays within minutes
    return d


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    """with oracle/_ref built, the arm runs the reference's own shader text compiled as C++ (kind "reference"), bounded by --ref-seconds"""
    d = _reference_arm()
    cb = d["cpu_baseline"]
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_hlsl_convert.so"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1
    if have_ref:
        assert "reference's own shader text" in cb["sample"] and d["ms_per_step"] * d["steps"] < 40e3     # 8 s asked for: a few sample frames per step


def test_reference_arm_falls_back_to_the_port():
    d = _reference_arm({"MPCVR_BENCH_FORCE_PORT": "1"})
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "band" in cb["sample"]          # 8 s over 60 steps -> a bounded band per step


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]
