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


def run(exe, cfg, iters, x, capture=None):
    path = os.path.join(os.path.dirname(exe), f"frame{cfg}.bin")
    np.ascontiguousarray(x, np.complex128).tofile(path)
    args = [exe, _lib.LDPC_TABLES, str(cfg), str(iters), path]
    if capture is not None:
        cpath = os.path.join(os.path.dirname(exe), f"capture{cfg}.bin")
        np.ascontiguousarray(capture, np.float64).tofile(cpath)
        args.append(cpath)
    return subprocess.run(args, capture_output=True, text=True)


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
    t = lines["tx_side"].split()  # transmit_byte(NO_FILTER), generate/detect pattern, fir_tx_apply, get_configuration, get_active_nsymb
    assert int(t[1]) == (mb.MODES[cfg]["Nsymb"] + mb.MODES[cfg]["preamble_nSymb"]) * 1088 and int(t[3]) == 16 * 1088
    assert float(t[5]) > 14 and int(t[7]) == 16 and float(t[9]) < 6 and int(t[11]) == 11 and int(t[13]) == mb.MODES[cfg]["Nsymb"]


@pytest.mark.gpu
def test_cpp_mirror_whole_receive_byte_on_a_reference_capture(exe, golden_dir):
    """receive_byte(double* data, int* out) with the reference's own signature on the committed pass-band capture fixture."""
    g = np.load(os.path.join(golden_dir, "frontend_mode08_clean.npz"))
    x = np.load(os.path.join(golden_dir, "rx_mode08.npz"))["x"]
    r = run(exe, 8, 50, x, capture=g["capture"].astype(np.float64))
    assert r.returncode == 0, r.stderr
    lines = dict(l.split(" ", 1) for l in r.stdout.strip().splitlines())
    st = dict(zip(ref_fields(), g["stats"]))
    for call in (0, 1):  # the second call carries the first one's link state; trial 0 decodes, so the outcome is the same
        f = lines[f"capture{call}"].split()
        assert int(f[0]) == int(st["decoded"]) == 1 and int(f[2]) == int(st["delay"]) and int(f[4]) == int(st["sync_trials"])
        assert int(f[6]) == int(st["iterations"]) and int(f[8]) == int(st["crc"])
        assert abs(float(f[10]) - st["freq_offset"]) < 1e-8 and float(f[12]) == st["coarse_metric"]
        assert int(f[14]) == int(g["state_out"][0])
        assert np.array_equal(np.array(lines[f"capture{call}_bytes"].split(), int), g["rx_payload"].astype(int))


def ref_fields():
    from oracle.ref import STAT12
    return STAT12
