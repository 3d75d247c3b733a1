"""The drop-in claim, executed (-m gpu): the REFERENCE'S OWN cl_telecom_system::receive_byte(), compiled from the reference sources with
only the lines INTEGRATION.md names re-pointed at libmercury_b200.so, against the unmodified build on the same capture buffers.

  oracle/_ref/libmercury_ref.so        the unmodified reference physical layer                       (make -C oracle ref)
  oracle/_ref/libmercury_ref_tail.so   telecom_system.cc:1132-1341 (+ the SNR report :1362-1400) -> mercury_b200_receive_baseband:
                                       the reference's CPU front-end feeds the GPU tail                
  oracle/_ref/libmercury_ref_whole.so  the whole body of receive_byte() -> mercury_b200_receive_byte   
                                       (oracle/dropin/make_patched.py + `make -C oracle dropin`; every other source file unmodified)

All three are driven through the same extern "C" harness (oracle/ref_driver.cc) that calls the object's public members the way the datalink
layer does (arq_common.cc:2668): load_configuration(int), receive_byte(double*, int*), and reads st_receive_stats back.
Compared, per capture: payload bytes, every reported st_receive_stats field, and the link state the next call depends on
(delay_of_last_decoded_message, freq_offset_of_last_decoded_message).  Integers and the verdict exact; SNR within 2e-3 dB (fp32 tail);
with the tail variant delay / sync_trials / freq_offset / coarse_metric / signal strength are the reference's own code, hence exact.
"""
import os

import numpy as np
import pytest

from oracle import ref
from tests import frontend_cases as fc

pytestmark = pytest.mark.gpu

_HERE = os.path.dirname(os.path.abspath(__file__))
TABLES = os.path.normpath(os.path.join(_HERE, "..", "mercury_b200", "data", "ldpc_tables.bin"))


def _need(so):
    if not (ref.available() and os.path.exists(so)):
        pytest.skip(f"{os.path.basename(so)} did not travel to this box (make -C oracle dropin needs /root/reference)")
    os.environ["MERCURY_B200_LDPC_TABLES"] = TABLES


def _check(label, want, got, exact_sync):
    for k in ("decoded", "crc", "all_zeros", "iterations", "delay", "sync_trials"):
        assert got[k] == want[k], (label, k, got[k], want[k])
    assert np.array_equal(np.asarray(got["payload"], np.int32), np.asarray(want["payload"], np.int32)), label
    assert got["last_delay"] == want["last_delay"], label
    tol = 0.0 if exact_sync else 1e-9
    for k in ("freq_offset", "last_freq", "coarse_metric"):
        assert abs(got[k] - want[k]) <= tol * max(1.0, abs(want[k])), (label, k, got[k], want[k])
    assert got["signal_dbm"] == want["signal_dbm"] or abs(got["signal_dbm"] - want["signal_dbm"]) <= 1e-9, label
    assert abs(got["snr"] - want["snr"]) <= 2e-3 * max(1.0, abs(want["snr"])), (label, got["snr"], want["snr"])
    assert got["frame_overflow_symbols"] == want["frame_overflow_symbols"], label


@pytest.mark.parametrize("variant", ["tail", "whole"])
@pytest.mark.parametrize("cfg", [0, 8, 11, 13, 16])
def test_patched_reference_receive_byte_equals_the_unpatched_one(cfg, variant):
    so = ref.SO_DROPIN_TAIL if variant == "tail" else ref.SO_DROPIN_WHOLE
    _need(so)
    plain, patched = ref.Ref(cfg, 50), ref.Ref(cfg, 50, so=so)
    assert patched.capture_samples() == plain.capture_samples() and patched.frame_bytes == plain.frame_bytes
    n_dec = 0
    for i, case in enumerate(fc.CASES):
        cap, pl, state = fc.make_capture(plain, case, 7000 + 100 * cfg + i)
        want = plain.receive_byte2(cap, *state)
        got = patched.receive_byte2(cap, *state)
        _check((cfg, variant, case), want, got, exact_sync=(variant == "tail"))
        n_dec += want["decoded"]
    assert n_dec >= 6  # the scenario set exercises both verdicts


@pytest.mark.parametrize("variant", ["tail", "whole"])
def test_patched_reference_follows_load_configuration_like_the_arq_layer(variant):
    """ARQ flips between data and ack configurations on one long-lived object (arq_commander.cc:431,574,661): the patched object must
    follow load_configuration(int) without being told anything else."""
    so = ref.SO_DROPIN_TAIL if variant == "tail" else ref.SO_DROPIN_WHOLE
    _need(so)
    plain, patched = ref.Ref(8, 50), ref.Ref(8, 50, so=so)
    for k, cfg in enumerate((8, 0, 16, 8, 13)):
        plain.load_configuration(cfg), patched.load_configuration(cfg)
        cap, pl, state = fc.make_capture(plain, "noise_light", 8100 + k)
        want, got = plain.receive_byte2(cap, *state), patched.receive_byte2(cap, *state)
        _check((cfg, variant, "reconfigured"), want, got, exact_sync=(variant == "tail"))
        assert want["decoded"] == 1 and list(want["payload"]) == [int(v) for v in pl]


@pytest.mark.parametrize("variant", ["tail", "whole"])
@pytest.mark.parametrize("cfg", [100, 102])
def test_patched_reference_robust_modes(cfg, variant):
    """ROBUST (MFSK) configurations through the same patched members: clean, noisy, very noisy, a frame running
    past the end of the buffer (frame_overflow_symbols), a restricted search start."""
    so = ref.SO_DROPIN_TAIL if variant == "tail" else ref.SO_DROPIN_WHOLE
    _need(so)
    plain, patched = ref.Ref(cfg, 50), ref.Ref(cfg, 50, so=so)
    n = plain.capture_samples()
    rng = np.random.default_rng(cfg)
    for case, sigma in enumerate((1e-4, 0.05, 0.3, 0.01, 0.01)):
        pl = rng.integers(0, 256, plain.frame_bytes)
        tx = plain.transmit_byte(pl)
        d = int(rng.integers(6, plain.buffer_Nsymb - (plain.Nsymb + 4) - 2)) * 1088 + int(rng.integers(0, 60))
        if case == 3:
            d = (plain.buffer_Nsymb - (plain.Nsymb + 4) + 3) * 1088
        L = min(tx.size, n - d)
        cap = np.zeros(n)
        cap[d:d + L] += tx[:L]
        cap = (cap + rng.normal(0, sigma, n)).astype(np.float32).astype(np.float64)
        start = 3 if case == 4 else 0
        want = plain.receive_byte2(cap, search_start_symb=start)
        got = patched.receive_byte2(cap, search_start_symb=start)
        _check((cfg, variant, case), want, got, exact_sync=(variant == "tail"))


def test_patched_reference_forwards_the_coarse_frequency_search_switch():
    """g_gui_state.coarse_freq_sync_enabled is a global of the reference's GUI state; the whole-body variant hands it to the library
    (mercury_b200_set_coarse_freq_sync) on every call.  Captures whose outcome the search changes (carrier offset 11-13 Hz: trial 1 fails
    as well, trial 2 decodes) must come out of the patched object as they come out of the plain one, switch on and off."""
    _need(ref.SO_DROPIN_WHOLE)
    plain, patched = ref.Ref(8, 50), ref.Ref(8, 50, so=ref.SO_DROPIN_WHOLE)
    n = plain.capture_samples()
    rng = np.random.default_rng(77)
    caps = []
    for df in (12.1, -11.4, 13.6, 3.0):
        tx = plain.transmit_byte(rng.integers(0, 256, plain.frame_bytes))
        d = int(rng.integers(6000, 30000))
        cap = np.zeros(n)
        cap[d:d + tx.size] += tx
        caps.append((fc.freq_shift(cap, df) + rng.normal(0, 0.1, n)).astype(np.float32).astype(np.float64))
    try:
        trials = {}
        for enable in (True, False):
            plain.set_coarse_freq_sync(enable), patched.set_coarse_freq_sync(enable)
            for i, cap in enumerate(caps):
                want, got = plain.receive_byte2(cap), patched.receive_byte2(cap)
                _check((8, "whole", f"cfs{int(enable)}", i), want, got, exact_sync=False)
                trials[(enable, i)] = want["sync_trials"]
        assert any(trials[(True, i)] != trials[(False, i)] for i in range(len(caps)))
    finally:
        plain.set_coarse_freq_sync(False), patched.set_coarse_freq_sync(False)
