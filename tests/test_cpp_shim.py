"""The header-only C++ shim (include/raft/distance/*.cuh) compiles against the C ABI everywhere and
passes its self-check on the GPU box."""
import os
import shutil
import subprocess

import pytest

from raft_b200 import _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "shim_test")


def compile_shim():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    _build.build()
    cmd = [nvcc, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-o", EXE,
           os.path.join(ROOT, "tests", "cpp", "shim_test.cu"), "-L", os.path.dirname(_build.SO_PATH),
           "-lraft_b200", "-Xlinker", "-rpath", "-Xlinker", os.path.dirname(_build.SO_PATH)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr


def test_shim_compiles_and_links():
    compile_shim()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_shim_runs_on_gpu():
    if not os.path.exists(EXE):
        compile_shim()
    res = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "PASS" in res.stdout, res.stdout + res.stderr


def test_block_selection_helpers_partition_the_blocks():
    """sel_to_blk / sel_count (shared by the kernels and the launch code) on the host: the exact sample
    pass, the trial screen and the main screen visit every y block exactly once, in ascending order."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = os.path.join(ROOT, "tests", "cpp", "host_logic_test")
    cmd = [nvcc, "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
           os.path.join(ROOT, "tests", "cpp", "host_logic_test.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    res = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0 and "PASS" in res.stdout, res.stdout + res.stderr
