"""CPU: the host-side staging pool of csrc/host_stage.cu (pageable std::vector clouds of the
drop-in call copied by worker threads into a pinned bounce buffer) built against a shim of the CUDA
runtime and stress-tested under ThreadSanitizer — the GPU box only ever sees it working."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tsan", [True, False])
def test_stage_pool_copies_every_byte(tmp_path, tsan):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    src = tmp_path / "host_stage.cc"
    shutil.copy(os.path.join(ROOT, "coxgraph_b200", "csrc", "host_stage.cu"), src)
    exe = tmp_path / ("stress_tsan" if tsan else "stress")
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-I", os.path.join(ROOT, "tests", "host", "cuda_shim"),
           "-I", os.path.join(ROOT, "coxgraph_b200", "csrc"),
           os.path.join(ROOT, "tests", "host", "stage_stress.cc"), str(src), "-o", str(exe), "-lpthread"]
    if tsan:
        cmd.insert(1, "-fsanitize=thread")
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0 and tsan:
        pytest.skip("ThreadSanitizer is not available: " + build.stderr[-200:])
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe), "60" if tsan else "300"], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "stage_stress ok" in run.stdout
    assert "ThreadSanitizer" not in run.stderr, run.stderr[-2000:]
