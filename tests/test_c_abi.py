"""The C ABI from a plain C host (tests/c_abi/abi_example.c): the header compiles as C and links against the shared
library without torch / Python / C++ (CPU tier), and the program runs against the GPU (GPU tier)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "abi_example")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-std=c99", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           os.path.join(ROOT, "tests", "c_abi", "abi_example.c"), "-L", os.path.join(ROOT, "poseprobe_b200"),
           "-lposeprobe_b200", "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lm", "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


def test_header_is_c_and_links(tmp_path):
    _build(tmp_path)


@pytest.mark.gpu
def test_c_host_program_runs(tmp_path, cuda):
    exe = _build(tmp_path)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.pathsep.join([os.path.join(ROOT, "poseprobe_b200"), os.path.join(CUDA, "lib64"),
                                              env.get("LD_LIBRARY_PATH", "")])
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stdout
    assert "c abi example ok" in r.stdout
