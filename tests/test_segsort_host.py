"""Host-side unit test of csrc/segsort.cuh (the per-row routines of the
experimental prefix-sorted consolidation path): compiled with g++ and run over
random inputs against std::stable_sort.  No GPU involved."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_segsort_host(tmp_path):
    exe = str(tmp_path / "segsort_host")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "database-stream-processor_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "segsort_host.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout
