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
@pytest.mark.parametrize("cfg", [0, 8, 13, 16])
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
