"""C-ABI surface (CPU): the library loads, exports every symbol include/mercury_b200.h declares, and refuses to
compute without a device (no CPU fallback)."""
import ctypes as C
import re

import numpy as np
import pytest

from mercury_b200 import _lib


def _declared():
    src = open(_lib.HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mercury_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 25
    L = C.CDLL(_lib.SO_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    _lib.lib()  # binds argtypes for all of them


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    L = _lib.lib()
    h = C.c_void_p()
    assert L.mercury_b200_create(0, C.byref(h)) == -2 and not h.value  # MERCURY_B200_ENODEV
    assert b"no CPU fallback" in L.mercury_b200_strerror(-2)
    from mercury_b200 import MercuryB200Error, TelecomSystemB200
    with pytest.raises(MercuryB200Error):
        TelecomSystemB200(0)


def test_host_side_table_build_and_errors():
    L = _lib.lib()
    n = C.c_size_t(0)
    assert L.mercury_b200_build_tables_host(b"/nonexistent/ldpc.bin", None, C.byref(n)) == -5  # EIO
    assert L.mercury_b200_build_tables_host(_lib.LDPC_TABLES.encode(), None, C.byref(n)) == 0 and n.value > 100000
    small = np.zeros(16, np.uint8)
    m = C.c_size_t(16)
    assert L.mercury_b200_build_tables_host(_lib.LDPC_TABLES.encode(), small.ctypes.data_as(C.c_void_p), C.byref(m)) == -1
    assert L.mercury_b200_get_frame_size_bytes(None) == -4 and L.mercury_b200_kernel_launches(None) == 0
    out = np.zeros(8, np.float32)
    assert L.mercury_b200_synth_frames(_lib.LDPC_TABLES.encode(), 99, 1, 0, 300.0, None, out.ctypes.data_as(C.c_void_p), None, 1) == -1


def test_get_configuration_ladder_matches_the_reference_thresholds():
    """get_configuration(SNR) (telecom_system.cc:3036-3106): host-only, checked against the ladder restated from the reference."""
    from mercury_b200 import _lib
    L = _lib.lib()
    rungs = [(12.5, 15), (9, 14), (7.5, 13), (6.5, 12), (4, 11), (3, 10), (1.5, 9), (0.5, 8), (-0.5, 7), (-1.5, 6), (-2.5, 5), (-3.5, 4),
             (-4.5, 3), (-6, 2), (-7.5, 1)]
    for thr, cfg in rungs:
        assert L.mercury_b200_get_configuration(thr + 1e-9) == cfg and L.mercury_b200_get_configuration(thr) == cfg - 1
    assert L.mercury_b200_get_configuration(-50.0) == 0 and L.mercury_b200_get_configuration(99.0) == 15
