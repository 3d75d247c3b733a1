"""Multi-GPU from C++ host code (tests/cpp/multi_gpu_host.cpp): one process, one handle and one host thread per GPU, tables built on
rank 0 and broadcast over NCCL through the C ABI (mercury_b200_broadcast_tables), contiguous frame shards, no data-path collective.
The program checks every rank's payloads itself (exit code) and prints one JSON line; here it runs on however many GPUs the box has
(`gpurun --gpus N -- ./multi_gpu_host ...` is the same binary at N = 2 / 4 / 8).  Without a GPU it must fail loudly (exit code 3)."""
import json
import os
import subprocess

import pytest

from mercury_b200 import _lib

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
SRC = os.path.join(ROOT, "tests", "cpp", "multi_gpu_host.cpp")
CUDA = "/usr/local/cuda"


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if not os.path.exists("/usr/include/nccl.h") and not os.path.exists(os.path.join(CUDA, "include", "nccl.h")):
        pytest.skip("nccl.h not installed on this box")
    _lib.lib()
    out = str(tmp_path_factory.mktemp("cpp") / "multi_gpu_host")
    libdir = os.path.join(ROOT, "mercury_b200")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
                           SRC, "-o", out, "-L", libdir, "-lmercury_b200", "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lnccl", "-pthread",
                           f"-Wl,-rpath,{libdir}"])
    return out


def test_cpp_multi_gpu_host_compiles_and_refuses_to_run_without_a_device(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu-marked test")
    r = subprocess.run([exe, _lib.LDPC_TABLES, "2", "64", "1"], capture_output=True, text=True)
    assert r.returncode == 3 and "no usable CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["c64", "i16"])
def test_cpp_multi_gpu_host_decodes_every_shard(exe, fmt):
    import torch
    n = torch.cuda.device_count()
    r = subprocess.run([exe, _lib.LDPC_TABLES, str(n), "3001", "2", "8", fmt], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == n and line["frames_total"] == 3001 * n and line["payload_mismatches"] == 0
    assert line["frames_decoded"] >= 0.99 * line["frames_total"] and line["e2e_frames_per_s"] > 0
