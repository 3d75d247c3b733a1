"""Multi-link batcher (csrc/mb_batcher.cpp): many threads making the reference's one-frame-per-call receive
pattern concurrently share one GPU batch.  CPU: the batching machinery in front of a test double, also under ThreadSanitizer;
GPU: real decodes, every link's result identical to a direct batch call."""
import os
import re
import subprocess

import pytest

from mercury_b200 import _lib

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
SRC = os.path.join(ROOT, "tests", "cpp", "batcher_test.cpp")
LIBDIR = os.path.join(ROOT, "mercury_b200")


def build(tmp, name, extra=(), extra_src=(), src=None):
    _lib.lib()
    out = str(tmp / name)
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-O1", "-pthread", *extra, "-I", os.path.join(ROOT, "include"), src or SRC,
                           *extra_src, "-o", out, "-L", LIBDIR, "-lmercury_b200", f"-Wl,-rpath,{LIBDIR}"])
    return out


def summary(stdout):
    return {k: float(v) for k, v in re.findall(r"(\w+) ([\d.]+)", stdout)}


def test_batching_machinery_with_a_test_double(tmp_path):
    exe = build(tmp_path, "batcher_test")
    # 64 links x 50 frames, batches of up to 32, 2 ms patience: every caller gets its own frame's result, frames are batched
    r = subprocess.run([exe, "mock", "64", "50", "32", "2000"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    s = summary(r.stdout)
    assert s["bad"] == 0 and s["frames"] == 3200 and s["mean_batch"] > 8 and s["batches"] < 3200 / 8
    # a single link cannot fill a batch: the patience timer must close it (no deadlock), one frame per batch
    r = subprocess.run([exe, "mock", "1", "20", "32", "500"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert summary(r.stdout)["batches"] == 20
    # more links than slots: callers queue for the next buffer
    r = subprocess.run([exe, "mock", "48", "20", "4", "100"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


def test_batching_machinery_under_thread_sanitizer(tmp_path):
    """The reference's README advertises sanitizer builds (README.md:64-69); the batcher is the one multi-threaded piece here."""
    try:
        # the batcher's own source is compiled INTO the instrumented binary (its definitions override the library's), so that
        # ThreadSanitizer sees every lock, wait and buffer hand-over of the machinery, not just the test's threads
        exe = build(tmp_path, "batcher_test_tsan", ("-fsanitize=thread", "-g"), (os.path.join(LIBDIR, "csrc", "mb_batcher.cpp"),))
    except subprocess.CalledProcessError:
        pytest.skip("ThreadSanitizer runtime not available")
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66")
    r = subprocess.run([exe, "mock", "16", "30", "8", "300"], capture_output=True, text=True, timeout=300, env=env)
    if "FATAL: ThreadSanitizer" in r.stderr and "unexpected memory mapping" in r.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow in this container")
    assert r.returncode == 0 and "WARNING: ThreadSanitizer" not in r.stderr, r.stdout + r.stderr[-3000:]


@pytest.mark.gpu
def test_links_share_gpu_batches_and_get_identical_results(tmp_path):
    exe = build(tmp_path, "batcher_test")
    r = subprocess.run([exe, "gpu", _lib.LDPC_TABLES, "8", "128", "40", "1024", "1000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    s = summary(r.stdout)
    assert s["bad"] == 0 and s["frames"] == 128 * 40 and s["mean_batch"] > 16


def test_passband_batcher_program_compiles(tmp_path):
    build(tmp_path, "batcher_passband_test", src=os.path.join(ROOT, "tests", "cpp", "batcher_passband_test.cpp"))


@pytest.mark.gpu
def test_links_share_whole_receive_byte_batches(tmp_path, golden_dir):
    """Pass-band flavour: 48 link threads x 6 synchronous receive_byte()-shaped calls on whole capture buffers (the committed reference
    capture at 288 different delays); every payload and stats record identical to one direct mercury_b200_receive_byte_batch call."""
    import numpy as np
    exe = build(tmp_path, "batcher_passband_test", src=os.path.join(ROOT, "tests", "cpp", "batcher_passband_test.cpp"))
    cap = np.load(os.path.join(golden_dir, "frontend_mode08_clean.npz"))["capture"].astype(np.float32)
    path = str(tmp_path / "capture.f32")
    cap.tofile(path)
    r = subprocess.run([exe, _lib.LDPC_TABLES, "8", path, "48", "6", "64", "2000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    s = summary(r.stdout)
    assert s["bad"] == 0 and s["calls"] == 288 and s["mean_batch"] > 4 and s["decoded"] >= 280
