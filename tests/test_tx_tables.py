"""The product's host-built TX tables (csrc/mb_tx.cu: preamble, pre-equalisation channel, transmit FIR designs) against the oracle,
on the CPU.  FIRs and preamble are bit-exact; the pre-equalisation channel goes through 1000 symbols of TX filters -> RX filter ->
DFT with a different (direct) transform, so it is compared to 1e-12."""
import ctypes as C

import numpy as np
import pytest

from mercury_b200 import _lib
from oracle import port


@pytest.mark.parametrize("cfg", [0, 3, 8, 10, 12, 13, 14, 16])
def test_tx_tables_match_the_oracle(cfg):
    L = _lib.lib()
    p = port.Port(cfg, 50)
    t = p.tx_tables()
    pre_eq = np.zeros(50, np.complex128)
    preamble = np.zeros(4 * 50, np.complex128)
    c1, c2 = np.zeros(97), np.zeros(97)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.mercury_b200_build_tx_tables_host(_lib.LDPC_TABLES.encode(), cfg, vp(pre_eq), vp(preamble), vp(c1), vp(c2)) == 0
    assert np.array_equal(c1, t["tx1"]) and np.array_equal(c2, t["tx2"])
    assert np.array_equal(preamble[:p.preamble_nSymb * 50], t["preamble"])
    assert np.abs(pre_eq - t["pre_eq"]).max() <= 1e-12 * np.abs(t["pre_eq"]).max()


def test_frontend_constants_match_the_oracle_bit_exactly():
    """The RX front-end's host-derived constants (csrc/mb_frontend.cu: both receive FIR designs, carrier table, thresholds): the sync
    decisions are only bit-identical to the reference if these are."""
    L = _lib.lib()
    t = port.Port(8, 50).frontend_tables()
    a, b, k, car = np.zeros(33), np.zeros(33), np.zeros(8), np.zeros(2 * 4096)
    vp = lambda x: x.ctypes.data_as(C.c_void_p)
    assert L.mercury_b200_build_frontend_tables_host(vp(a), vp(b), vp(k), vp(car), 4096) == 0
    assert np.array_equal(a, t["ts"]) and np.array_equal(b, t["data"])
    assert list(k) == [t["fs"], t["fc"], t["amp"], t["bandwidth"], t["trials_max"], t["use_last_time"], t["use_last_freq"], t["ignore_limit"]]
    i = np.arange(4096, dtype=np.float64)
    ph = 2 * np.pi * t["fc"] * i * (1.0 / t["fs"])  # the reference's expression order (ofdm.cc:2332-2333); numpy's cos/sin are libm's here
    assert np.abs(car[0::2] - np.cos(ph)).max() <= 2e-16 and np.abs(car[1::2] - np.sin(ph)).max() <= 2e-16
