"""TX chain to pass-band: the C restatement of transmit_byte(SINGLE_MESSAGE) -- preamble, pre-equalisation
channel, transmit FIR designs, interpolation, mixing with the running carrier counter, PAPR clip, two FIRs -- against the UNMODIFIED
reference (oracle/_ref), bit-exact, and against the committed reference fixture."""
import os

import numpy as np
import pytest

from oracle import port, ref


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libmercury_ref.so not built")
@pytest.mark.parametrize("cfg", list(range(17)))
def test_tx_tables_and_passband_bit_exact(cfg):
    r, p = ref.Ref(cfg, 50), port.Port(cfg, 50)
    tr, tp = r.tx_tables(), p.tx_tables()
    for k in tr:
        assert np.array_equal(np.asarray(tr[k]), np.asarray(tp[k])), k
    rng = np.random.default_rng(31 + cfg)
    for start in (tr["start_sample_after_init"], 987654):
        pl = rng.integers(0, 256, r.frame_bytes - (3 if start > 2000 else 0))  # short payloads are zero padded before the CRC
        a, sa = r.transmit_byte2(pl, start)
        b, sb = p.transmit_byte2(pl, start)
        assert sa == sb == start + tr["total_frame_size"] and np.array_equal(a, b)
    # the reference's first transmit_byte after init starts one symbol into the carrier (get_pre_equalization_channel leaves it there)
    pl = rng.integers(0, 256, r.frame_bytes)
    assert np.array_equal(ref.Ref(cfg, 50).transmit_byte(pl), p.transmit_byte2(pl, tp["start_sample_after_init"])[0])


def test_port_against_tx_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "tx_mode16.npz"))
    p = port.Port(16, 50)
    out, after = p.transmit_byte2(g["payload"], int(g["start_sample"]))
    assert after == int(g["start_sample_after"]) and np.array_equal(out, g["passband"])
    t = p.tx_tables()
    assert np.array_equal(t["pre_eq"], g["pre_eq"]) and np.array_equal(t["preamble"], g["preamble"])
    assert np.array_equal(t["tx1"], g["tx1"]) and np.array_equal(t["tx2"], g["tx2"])


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libmercury_ref.so not built")
@pytest.mark.parametrize("cfg", [8, 16])
def test_streaming_message_locations_bit_exact(cfg):
    r, p = ref.Ref(cfg, 50), port.Port(cfg, 50)
    r.reset_tx_stream(), p.reset_tx_stream()
    rng = np.random.default_rng(cfg)
    sr = sp = 1088
    for loc in (0, 1, 1, 2):
        pl = rng.integers(0, 256, r.frame_bytes)
        a, sr = r.transmit_byte_loc(pl, sr, loc)
        b, sp = p.transmit_byte_loc(pl, sp, loc)
        assert sr == sp and np.array_equal(a, b), (cfg, loc)
