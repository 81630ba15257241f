"""The CPU arm of bench.py (`--impl reference`) honours the driver's JSON contract, runs the same circuit steps as the
GPU arm and spells out a bounded sample; the GPU arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, env=e)
    lines = [ln for ln in p.stdout.strip().splitlines() if ln.startswith("{")]
    return p.returncode, (json.loads(lines[-1]) if lines else None)


def test_reference_arm_line():
    rc, d = run_bench("--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "3", "--events-per-step", "20000")
    assert rc == 0
    assert d["impl"] == "reference" and d["metric"] == "nexmark_events_per_sec" and d["unit"] == "events/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 2 and d["steps"] == 3 and d["warmup"] == 3
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "40000 events" in cb["sample"]
    assert d["config"]["query"] == "q3" and d["config"]["events_per_circuit_step_per_gpu"] == 20000
    assert "reference_sample" not in d["config"]   # small run: nothing was cut


def test_reference_arm_bounded_sample_is_declared():
    # 8 x 3 M events per circuit step, (5 + 20) bench steps: 8 circuit steps per bench step would be 4.8 G events
    sys.path.insert(0, ROOT)
    import bench

    class A:
        events_per_step, circuit_steps, steps, warmup, cpu_threads = 3_000_000, 0, 20, 5, 2

    calls = {}

    def fake_gen(query, rank, world, n_steps, E, pinned, rate=None):
        calls["n_steps"], calls["E"] = n_steps, E
        raise RuntimeError("stop before generating")

    orig = bench.gen_steps
    bench.gen_steps = fake_gen
    try:
        try:
            bench.run_reference(A, "q3", gpus=8)
        except RuntimeError:
            pass
    finally:
        bench.gen_steps = orig
    assert calls["E"] == 24_000_000
    assert calls["n_steps"] * calls["E"] <= 2_400_000_000 and calls["n_steps"] == 25 * 4   # 4 of 8 circuit steps per bench step


def test_gpu_arm_refuses_without_device():
    rc, d = run_bench("--steps", "1", env={"CUDA_VISIBLE_DEVICES": ""})
    assert rc != 0 and "no CUDA device" in d["error"]
