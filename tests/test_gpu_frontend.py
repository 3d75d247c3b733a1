"""GPU parity of the RX front-end row: mercury_b200_receive_byte(_batch) -- pass-band capture in, payload
out -- against the oracle's whole receive_byte() (the unmodified reference when oracle/_ref travelled to this box, else the C
restatement) on the capture scenarios of tests/frontend_cases.py and against the committed reference fixtures.

Bars: the sync decisions (delay, sync_trials), the Schmidl-Cox metric, the verdict fields and the payload bytes are bit-exact;
continuous reported values within the tolerances written below; the post-synchronisation base-band within 1e-9 relative."""
import glob
import os

import numpy as np
import pytest

import mercury_b200 as mb
from oracle import port, ref
from tests import frontend_cases as fc

pytestmark = pytest.mark.gpu

EXACT = ("delay", "sync_trials", "decoded", "crc", "all_zeros", "iterations", "coarse_metric")


@pytest.fixture(scope="module")
def ts():
    t = mb.TelecomSystemB200(0)
    yield t
    t.close()


def _compare(o, st, payload, bb, label):
    got = dict(delay=int(st["delay"]), sync_trials=int(st["sync_trials"]), decoded=int(st["message_decoded"]), crc=int(st["crc"]),
               all_zeros=int(st["all_zeros"]), iterations=int(st["iterations_done"]), coarse_metric=float(st["coarse_metric"]))
    for k in EXACT:
        assert got[k] == o[k], (label, k, got[k], o[k])
    assert int(st["delay_of_last_decoded_message"]) == o["last_delay"], label
    assert abs(float(st["freq_offset_of_last_decoded_message"]) - o["last_freq"]) <= 1e-9, label
    assert abs(float(st["freq_offset"]) - o["freq_offset"]) <= 1e-9, (label, float(st["freq_offset"]), o["freq_offset"])
    assert abs(float(st["signal_stregth_dbm"]) - o["signal_dbm"]) <= 1e-9 or (np.isinf(o["signal_dbm"]) and np.isinf(st["signal_stregth_dbm"])), label
    assert abs(float(st["SNR"]) - o["snr"]) <= 2e-3 * max(1.0, abs(o["snr"])), (label, float(st["SNR"]), o["snr"])
    assert np.array_equal(payload.astype(np.int32), np.asarray(o["payload"], np.int32)), label
    if bb is not None and (o["sync_trials"] > 0 or o["decoded"]):
        ref_bb = o["baseband"]
        err = np.abs(bb - ref_bb).max() / max(np.abs(ref_bb).max(), 1e-30)
        assert err <= 1e-9, (label, err)


@pytest.mark.parametrize("cfg", list(range(17)))
def test_receive_byte_scenarios(ts, cfg):
    if not ref.available():
        pytest.skip("scenario frames come from the reference's transmit_byte (oracle/_ref not on this box)")
    r = ref.Ref(cfg, 50)
    ts.load_configuration(cfg, 50)
    assert ts.get_capture_samples() == r.capture_samples()
    cases = fc.CASES if cfg in (8, 16) else ["clean", "noise_light", "freq_offset", "late", "silence", "last_good_state"]
    caps, oracle_out, states = [], [], mb.new_receive_stats(len(cases))
    for i, case in enumerate(cases):
        cap, pl, state = fc.make_capture(r, case, 100 * cfg + i)
        caps.append(cap)
        states["delay_of_last_decoded_message"][i] = state[0]
        states["freq_offset_of_last_decoded_message"][i] = state[1]
        oracle_out.append(r.receive_byte2(cap, *state))
    caps = np.stack(caps)
    # the whole batch in one call (exercises the per-capture state machine with captures in different phases) ...
    payload, st, bb = ts.receive_byte_batch(caps, states.copy(), want_baseband=True)
    for i, case in enumerate(cases):
        _compare(oracle_out[i], st[i], payload[i], bb[i], f"cfg{cfg}/{case}/batch")
    assert sum(o["decoded"] for o in oracle_out) >= 3
    # ... float32 samples (exactly the values the reference saw: the scenario captures are float32-representable) ...
    payload32, st32, _ = ts.receive_byte_batch(caps.astype(np.float32), states.copy())
    assert np.array_equal(payload32, payload) and st32.tobytes() == st.tobytes()
    # ... the PCM capture formats of the reference's audio layer (audioio.c:893-940): int16 (x / 32768.0) and int32 (x / INT_MAX),
    # converted on the device, against the double entry point fed with the host-converted values and (first captures) the oracle
    pcm16 = np.clip(np.rint(caps * 32768.0 * 0.5), -32768, 32767).astype(np.int16)
    pcm32 = np.clip(np.rint(caps * 2147483647.0 * 0.5), -2147483647, 2147483647).astype(np.int32)
    for pcm, conv in ((pcm16, pcm16.astype(np.float64) / 32768.0), (pcm32, pcm32.astype(np.float64) / float(2147483647))):
        p_i, st_i, _ = ts.receive_byte_batch(pcm, states.copy())
        p_d, st_d, _ = ts.receive_byte_batch(conv, states.copy())
        assert np.array_equal(p_i, p_d) and st_i.tobytes() == st_d.tobytes(), pcm.dtype
        for i in (0, 1):
            _compare(r.receive_byte2(conv[i], int(states["delay_of_last_decoded_message"][i]), float(states["freq_offset_of_last_decoded_message"][i])),
                     st_i[i], p_i[i], None, f"cfg{cfg}/{cases[i]}/{pcm.dtype}")
    # measure_signal_only (telecom_system.cc:1520-1541): the same mix + time-sync FIR + mean power, nothing else
    dbm = ts.measure_signal_only(caps)
    for i in range(len(cases)):
        want_dbm = oracle_out[i]["signal_dbm"]
        assert (np.isinf(dbm[i]) and np.isinf(want_dbm)) or abs(dbm[i] - want_dbm) <= 1e-9, cases[i]  # (a silent capture reads -inf dBm in both)
    # ... and the single-capture call in the reference's own types
    for i in (0, len(cases) - 1):
        one = states[i:i + 1].copy()
        out, one = ts.receive_byte(caps[i], one)
        _compare(oracle_out[i], one[0], out, None, f"cfg{cfg}/{cases[i]}/single")


@pytest.mark.parametrize("cfg", [8, 13])
def test_receive_byte_with_coarse_frequency_search(ts, cfg):
    """g_gui_state.coarse_freq_sync_enabled (off by default): when trial 0 fails, receive_byte() runs Schmidl-Cox with the time-sync filter at
    fc - 30, fc and fc + 30 Hz before trial 1 and fine-syncs on the time-sync base-band (telecom_system.cc:949-1013).  Captures with a
    carrier offset of 10-14 Hz are where that changes the outcome (trial 1 then fails too, trial 2 decodes); offsets that would need the
    +-30 Hz correction fail the Schmidl-Cox gate before any trial runs.  Every field against the reference with the flag on, then off again."""
    if not ref.available():
        pytest.skip("scenario frames come from the reference's transmit_byte (oracle/_ref not on this box)")
    r = ref.Ref(cfg, 50)
    ts.load_configuration(cfg, 50)
    n = r.capture_samples()
    rng = np.random.default_rng(40 + cfg)
    caps = []
    for df, sigma in ((12.1, 0.1), (-11.4, 0.1), (11.0, 0.1), (-10.7, 0.01), (13.6, 0.1), (-11.1, 0.1), (8.6, 0.1), (28.0, 0.01), (-12.0, 0.25), (0.0, 0.02)):
        tx = r.transmit_byte(rng.integers(0, 256, r.frame_bytes))
        d = int(rng.integers(6000, 30000))
        cap = np.zeros(n)
        cap[d:d + tx.size] += tx
        caps.append((fc.freq_shift(cap, df) + rng.normal(0, sigma, n)).astype(np.float32).astype(np.float64))
    for i, case in enumerate(("noise_heavy", "two_frames", "weak_frame", "tone_then_frame")):
        caps.append(fc.make_capture(r, case, 900 + 10 * cfg + i)[0])
    caps = np.stack(caps)
    states = mb.new_receive_stats(len(caps))
    states["delay_of_last_decoded_message"][:] = -1
    try:
        changed = 0
        for enable in (True, False):
            r.set_coarse_freq_sync(enable)
            ts.set_coarse_freq_sync(enable)
            want = [r.receive_byte2(c) for c in caps]
            payload, st, bb = ts.receive_byte_batch(caps, states.copy(), want_baseband=True)
            for i in range(len(caps)):
                _compare(want[i], st[i], payload[i], bb[i], f"cfg{cfg}/cfs{int(enable)}/{i}")
            if enable:
                with_search = [w["sync_trials"] for w in want]
            else:
                changed = sum(a != b["sync_trials"] for a, b in zip(with_search, want))
        assert changed >= 2  # the search was on the path of some captures
    finally:
        r.set_coarse_freq_sync(False)
        ts.set_coarse_freq_sync(False)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "frontend_*.npz"))))
def test_receive_byte_reference_fixture(ts, path):
    g = np.load(path)
    ts.load_configuration(int(g["config"]), int(g["ldpc_iters"]))
    st = mb.new_receive_stats(1)
    st["delay_of_last_decoded_message"][0] = int(g["state_in"][0])
    st["freq_offset_of_last_decoded_message"][0] = float(g["state_in"][1])
    payload, st, bb = ts.receive_byte_batch(g["capture"].astype(np.float64), st, want_baseband=True)
    o = {k: (float(g["stats"][i]) if k in ("snr", "freq_offset", "coarse_metric", "signal_dbm") else int(g["stats"][i])) for i, k in enumerate(ref.STAT12)}
    o.update(payload=g["rx_payload"], baseband=g["baseband"], last_delay=int(g["state_out"][0]), last_freq=float(g["state_out"][1]))
    _compare(o, st[0], payload[0], bb[0], os.path.basename(path))


def test_receive_byte_many_links_against_port(ts):
    """256 captures of one mode in one call (several chunks when MERCURY_B200_FE_CHUNK is small): every capture against the C restatement."""
    cfg = 16
    src = ref.Ref(cfg, 50) if ref.available() else None
    if src is None:
        pytest.skip("needs the reference's transmit_byte for frames")
    p = port.Port(cfg, 50)
    ts.load_configuration(cfg, 50)
    rng = np.random.default_rng(5)
    n = 96
    caps = np.stack([fc.make_capture(src, fc.CASES[int(rng.integers(0, 4))], 7000 + i)[0] for i in range(n)])
    payload, st, _ = ts.receive_byte_batch(caps)
    for i in range(n):
        o = p.receive_byte2(caps[i])
        _compare(o, st[i], payload[i], None, f"link{i}")
    assert int(st["message_decoded"].sum()) >= n // 2


def test_chunked_and_edge_arguments(ts, golden_dir, monkeypatch):
    """Chunk boundaries (a handle whose device and host chunks are 5 and 3 captures), empty batches and bad arguments."""
    g = np.load(os.path.join(golden_dir, "frontend_mode08_clean.npz"))
    cap = g["capture"].astype(np.float32)
    caps = np.stack([np.roll(cap, 53 * i) for i in range(13)])
    ts.load_configuration(8, 50)
    want_p, want_s, _ = ts.receive_byte_batch(caps)
    monkeypatch.setenv("MERCURY_B200_FE_CHUNK", "5")
    monkeypatch.setenv("MERCURY_B200_FE_HOST_CHUNK", "3")
    t2 = mb.TelecomSystemB200(0)
    try:
        t2.load_configuration(8, 50)
        p2, s2, _ = t2.receive_byte_batch(caps)
        assert np.array_equal(p2, want_p) and s2.tobytes() == want_s.tobytes()
        import torch
        d = torch.from_numpy(caps).cuda()
        d_st = torch.from_numpy(mb.new_receive_stats(13).view(np.uint8).reshape(13, -1)).cuda()
        d_out = torch.zeros((13, want_p.shape[1]), dtype=torch.uint8, device="cuda")
        t2.receive_byte_batch_device(d, mb.SAMPLES_F32, 13, d_out, d_st, stream=torch.cuda.current_stream().cuda_stream)
        assert np.array_equal(d_out.cpu().numpy(), want_p)
        assert d_st.cpu().numpy().view(mb.RECEIVE_STATS_DTYPE).reshape(-1).tobytes() == want_s.tobytes()
        # empty batch: nothing to do, no error
        e_p, e_s, _ = t2.receive_byte_batch(np.zeros((0, t2.get_capture_samples()), np.float32))
        assert e_p.shape[0] == 0 and e_s.shape[0] == 0
        with pytest.raises(TypeError):
            t2.receive_byte_batch(np.zeros(t2.get_capture_samples(), np.int64))
        with pytest.raises(ValueError):
            t2.receive_byte_batch(np.zeros(1000, np.float32))
        with pytest.raises(mb.MercuryB200Error):
            t2.load_configuration(55, 50)
    finally:
        t2.close()
    assert int(want_s["message_decoded"].sum()) == 13
