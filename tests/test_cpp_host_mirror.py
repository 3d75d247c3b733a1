"""The C++ host-side mirror of the reference's cl_telecom_system (include/mercury_b200.hpp) driven from a C++ program the
way reference code drives the reference object: compiled with plain g++ -std=c++14 (the reference's own dialect, Makefile:43),
no CUDA headers.  Without a GPU it must fail loudly (exit code 3), never fall back."""
import os
import subprocess

import numpy as np
import pytest

import mercury_b200 as mb
from mercury_b200 import _lib

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    _lib.lib()  # builds the shared library if needed
    out = str(tmp_path_factory.mktemp("cpp") / "host_mirror_test")
    libdir = os.path.join(ROOT, "mercury_b200")
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", out,
                           "-L", libdir, "-lmercury_b200", f"-Wl,-rpath,{libdir}"])
    return out


def run(exe, cfg, iters, x):
    path = os.path.join(os.path.dirname(exe), f"frame{cfg}.bin")
    np.ascontiguousarray(x, np.complex128).tofile(path)
    return subprocess.run([exe, _lib.LDPC_TABLES, str(cfg), str(iters), path], capture_output=True, text=True)


def test_cpp_mirror_compiles_and_refuses_to_run_without_a_device(exe, golden_dir):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu-marked test")
    g = np.load(os.path.join(golden_dir, "rx_mode08.npz"))
    r = run(exe, 8, 50, g["x"])
    assert r.returncode == 3 and "no usable CUDA device" in r.stderr  # no CPU fallback exists


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [0, 8, 16])
def test_cpp_mirror_reproduces_the_reference_frame(exe, golden_dir, cfg):
    g = np.load(os.path.join(golden_dir, f"rx_mode{cfg:02d}.npz"))
    r = run(exe, cfg, int(g["ldpc_iters"]), g["x"])
    assert r.returncode == 0, r.stderr
    lines = dict(l.split(" ", 1) for l in r.stdout.strip().splitlines())
    head = lines["decoded"].split()
    assert int(head[0]) == int(g["decoded"]) == mb.YES and int(head[2]) == int(g["iterations"]) and int(head[4]) == int(g["crc"])
    assert abs(float(head[8]) - float(g["snr"])) < 2e-3
    fb = mb.MODES[cfg]["frame_bytes"]
    assert int(head[10]) == fb and int(head[12]) == mb.MODES[cfg]["nReal"] - 16
    got = np.array(lines["bytes"].split(), int)
    assert np.array_equal(got, g["payload"].astype(int))
    bits = np.array(lines["bits"].split(), int)
    ref_bytes = g["bytes"].astype(int)  # the reference's hd_decoded_data_byte: payload + the two CRC bytes
    want = ((ref_bytes[:, None] >> np.arange(8)[None, :]) & 1).reshape(-1)
    assert np.array_equal(bits, want)  # receive_bit: LSB first, CRC bytes included (telecom_system.cc:636-644)
    assert int(lines["after_bad_config"].split()[1]) == fb
