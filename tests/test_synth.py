"""The host TX synthesiser (csrc/mb_synth.cpp, driven by the RX index tables run backwards) against the oracle's TX chain."""
import numpy as np
import pytest

import mercury_b200 as mb
from oracle import port


@pytest.mark.parametrize("cfg", range(17))
def test_noise_free_waveform_equals_oracle_tx(cfg):
    p = port.Port(cfg)
    x, pl = mb.synth_frames(cfg, 3, seed=11 + cfg, esn0_db=300.0, n_threads=2)
    assert x.shape == (3, p.Nsymb, 272) and pl.shape == (3, p.frame_bytes)
    for f in range(3):
        ref = p.tx_baseband(pl[f].astype(np.int32))
        assert np.abs(x[f].reshape(-1) - ref).max() <= 2e-6 * np.abs(ref).max()
        o = p.rx_tail(x[f].reshape(-1).astype(np.complex128))
        assert o["decoded"] == 1 and o["iterations"] == 0 and np.array_equal(o["payload"], pl[f])


def test_noise_level_and_determinism():
    a, pa = mb.synth_frames(8, 8, seed=5, esn0_db=2.5, n_threads=1)
    b, pb = mb.synth_frames(8, 8, seed=5, esn0_db=2.5, n_threads=4)
    assert np.array_equal(a, b) and np.array_equal(pa, pb)
    clean, _ = mb.synth_frames(8, 8, seed=5, esn0_db=300.0)
    n = (a - clean).reshape(-1)
    sigma = 10 ** (-2.5 / 20) * 16  # telecom_system.cc:98,139-153
    assert abs(np.sqrt((np.abs(n) ** 2).mean()) / sigma - 1) < 0.02
    assert abs(n.real.std() / n.imag.std() - 1) < 0.03 and abs(n.mean()) < 0.05 * sigma
    p = port.Port(8)
    ok = sum(int(np.array_equal(p.rx_tail(a[f].reshape(-1).astype(np.complex128))["payload"], pa[f])) for f in range(8))
    assert ok == 8
