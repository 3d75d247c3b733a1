"""RX front-end row: the C restatement of the whole receive_byte() against the UNMODIFIED reference's
receive_byte() (oracle/_ref) on pass-band capture buffers, bit-exact in every reported field, and against the committed
fixtures generated from the reference (tests/golden/make_golden_frontend.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import port, ref
from tests import frontend_cases as fc


def _same(a, b):
    for k in ref.STAT12:
        assert a[k] == b[k], (k, a[k], b[k])
    assert np.array_equal(a["payload"], b["payload"])
    assert a["last_delay"] == b["last_delay"] and a["last_freq"] == b["last_freq"]


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libmercury_ref.so not built")
@pytest.mark.parametrize("cfg", list(range(17)))
def test_frontend_tables_and_receive_byte_bit_exact(cfg):
    r, p = ref.Ref(cfg, 50), port.Port(cfg, 50)
    tr, tp = r.frontend_tables(), p.frontend_tables()
    for k in tr:
        assert np.array_equal(np.asarray(tr[k]), np.asarray(tp[k])), k
    assert (r.buffer_Nsymb, r.total_frame_size) == (p.buffer_Nsymb, p.total_frame_size)
    cases = fc.CASES if cfg in (8, 16) else ["clean", "noise_light", "freq_offset", "late"]
    n_dec = 0
    for i, case in enumerate(cases):
        cap, pl, state = fc.make_capture(r, case, 100 * cfg + i)
        a = r.receive_byte2(cap, *state)
        b = p.receive_byte2(cap, *state)
        _same(a, b)
        if a["sync_trials"] > 0 or a["decoded"]:
            assert np.array_equal(a["baseband"], b["baseband"]), case
        if a["decoded"]:
            assert pl is None or np.array_equal(a["payload"], pl), case
            n_dec += 1
    assert n_dec >= 3


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "frontend_*.npz"))))
def test_port_against_frontend_fixture(path):
    g = np.load(path)
    p = port.Port(int(g["config"]), int(g["ldpc_iters"]))
    b = p.receive_byte2(g["capture"].astype(np.float64), int(g["state_in"][0]), float(g["state_in"][1]))
    for i, k in enumerate(ref.STAT12):
        assert b[k] == g["stats"][i], (k, b[k], g["stats"][i])
    assert np.array_equal(b["payload"], g["rx_payload"])
    assert b["last_delay"] == g["state_out"][0] and b["last_freq"] == g["state_out"][1]
    assert np.array_equal(b["baseband"], g["baseband"])


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libmercury_ref.so not built")
def test_optional_coarse_frequency_search_restated():
    """g_gui_state.coarse_freq_sync_enabled (off by default): the +-30 Hz search of trial 1 (telecom_system.cc:949-1013) is restated in the
    oracle and pinned here (the product's version: tests/test_gpu_frontend.py).  A carrier offset large enough to need it (> 23 Hz, half a
    carrier spacing) already fails the Schmidl-Cox gates before any trial runs -- the real part of the lag-1024 correlation turns negative --
    so the branch is only reachable where Moose alone would have coped."""
    r, p = ref.Ref(8, 50), port.Port(8, 50)
    rng = np.random.default_rng(3)
    n = r.capture_samples()
    try:
        for df in (28.0, 6.0, -12.0):
            for en in (True, False):
                r.set_coarse_freq_sync(en), p.set_coarse_freq_sync(en)
                pl = rng.integers(0, 256, r.frame_bytes)
                tx = r.transmit_byte(pl)
                d = int(rng.integers(6000, 30000))
                cap = np.zeros(n)
                cap[d:d + tx.size] += tx
                cap = (fc.freq_shift(cap, df) + rng.normal(0, 0.25 if df < 0 else 0.01, n)).astype(np.float32).astype(np.float64)
                _same(r.receive_byte2(cap), p.receive_byte2(cap))
    finally:
        r.set_coarse_freq_sync(False)
