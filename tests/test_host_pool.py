"""Host-side engine logic that needs no GPU: the worker pool's parallel-for (compiled from the engine header)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="no host compiler")
def test_worker_pool_runs_every_index_exactly_once(tmp_path):
    exe = tmp_path / "pool_stress"
    cuda_inc = "/usr/local/cuda/include"
    cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-I", cuda_inc, "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "pool_stress.cpp"), "-o", str(exe)]
    subprocess.run(cmd, check=True, capture_output=True, timeout=300)
    out = subprocess.run([str(exe), "20000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "bad=0" in out.stdout
