"""GPU parity of the MFSK row: the ROBUST_0..2 tail (FFT + cl_mfsk::demod + de-interleave + LDPC + CRC) through
the same batch entry points as the OFDM modes, and the tone-pattern detectors (time_sync_mfsk, detect_ack_pattern with the ACK and
BREAK tones), against the oracle (the unmodified reference when oracle/_ref travelled to this box, else the C restatement).

Bars: payload / CRC / decision exact, iteration count exact, LLRs within 1e-4 of max(|llr|, median|llr|) (the demodulator is fp32, the
reference double; LLRs are clamped to +-5), SNR exactly the reference's 0 / -99.9; sync delay and matched counts exact, pattern
metrics within 1e-9."""
import numpy as np
import pytest

import mercury_b200 as mb
from oracle import port, ref
from tests import mfsk_cases as mc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ts():
    t = mb.TelecomSystemB200(0)
    yield t
    t.close()


def _oracle(cfg):
    return ref.Ref(cfg, 50) if ref.available() else port.Port(cfg, 50)


@pytest.mark.parametrize("cfg", [100, 101, 102])
def test_robust_tail_against_oracle(ts, cfg):
    o = _oracle(cfg)
    g = ts.load_configuration(cfg, 50)
    R = mb.ROBUST_MODES[cfg]
    assert (g["Nsymb"], g["K"], g["frame_bytes"], g["nBits"], g["M"]) == (R["Nsymb"], R["K"], R["frame_bytes"], 1600, 200)
    assert (o.Nsymb, o.frame_bytes) == (g["Nsymb"], g["frame_bytes"])
    rng = np.random.default_rng(900 + cfg)
    sigmas = [0.0, 10.0, 25.0, 40.0, 55.0, 70.0, 90.0, 120.0]
    xs, pls = [], []
    for sg in sigmas:
        pl = rng.integers(0, 256, g["frame_bytes"])
        x = o.tx_baseband(pl)
        xs.append((x + sg * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))).astype(np.complex64))
        pls.append(pl)
    x = np.stack(xs).reshape(len(sigmas), g["Nsymb"], 272)
    payload, st, llr = ts.demod_decode_batch(x, want_llr=True)
    n_dec = 0
    for i in range(len(sigmas)):
        r = o.rx_tail(x[i].reshape(-1).astype(np.complex128))
        tol = 1e-4 * np.maximum(np.abs(r["llr_cw"]), np.median(np.abs(r["llr_cw"])))
        assert (np.abs(llr[i] - r["llr_cw"]) <= tol).all(), (cfg, sigmas[i], np.abs(llr[i] - r["llr_cw"]).max())
        assert int(st["message_decoded"][i]) == r["decoded"], (cfg, sigmas[i])
        assert int(st["iterations_done"][i]) == r["iterations"], (cfg, sigmas[i], int(st["iterations_done"][i]), r["iterations"])
        assert float(st["SNR"][i]) == pytest.approx(r["snr"], abs=1e-4)
        if r["decoded"]:
            assert int(st["crc"][i]) == 0 and np.array_equal(payload[i], np.asarray(pls[i], np.uint8)) and np.array_equal(payload[i], r["payload"].astype(np.uint8))
            n_dec += 1
    assert n_dec >= 3 and n_dec < len(sigmas)  # both outcomes exercised
    # the single-frame call in the reference's types
    out, s1 = ts.receive_baseband(x[1].reshape(-1).astype(np.complex128))
    assert s1["message_decoded"] == 1 and np.array_equal(out, np.asarray(pls[1], np.int32))


@pytest.mark.parametrize("cfg", [100, 101])
def test_pattern_detectors_against_oracle(ts, cfg):
    o = _oracle(cfg)
    ts.load_configuration(cfg, 50)
    kinds = ["ack", "break", "frame", "noise", "ack", "frame"]
    bufs, poss = zip(*[mc.pattern_buffer(o, k, 50 * cfg + i) for i, k in enumerate(kinds)])
    b = np.stack(bufs)
    for start in (0, 3):
        res = ts.mfsk_patterns_batch(b, search_start_symb=start)
        for i, k in enumerate(kinds):
            assert int(res["time_sync_delay"][i]) == o.time_sync_mfsk(b[i], start), (k, start)
            for brk, name in ((False, "ack"), (True, "break")):
                m, matched = o.detect_ack_pattern(b[i], brk)
                assert abs(float(res[name + "_metric"][i]) - m) <= 1e-9 * max(1.0, m), (k, name)
                assert int(res[name + "_matched"][i]) == matched, (k, name)
    res = ts.mfsk_patterns_batch(b)
    assert res["ack_metric"][0] > 8 and res["ack_matched"][0] >= 14 and res["break_metric"][0] < res["ack_metric"][0] / 2
    assert res["break_metric"][1] > 8 and res["ack_metric"][1] < res["break_metric"][1] / 2
    assert int(res["time_sync_delay"][2]) == (poss[2] // 1088) * 1088
    # complex64 buffers: compared with the oracle on exactly those values widened to double
    b32 = b.astype(np.complex64)
    res32 = ts.mfsk_patterns_batch(b32)
    for i in range(len(kinds)):
        w = b32[i].astype(np.complex128)
        assert int(res32["time_sync_delay"][i]) == o.time_sync_mfsk(w, 0)
        assert abs(float(res32["ack_metric"][i]) - o.detect_ack_pattern(w, False)[0]) <= 1e-9 * max(1.0, float(res32["ack_metric"][i]))


@pytest.mark.parametrize("cfg", [100, 101, 102])
def test_robust_transmit_byte_against_oracle(ts, cfg):
    """transmit_byte(SINGLE_MESSAGE) in the ROBUST modes: tone preamble, cl_mfsk::mod, drive-level boost, no pre-equalisation."""
    o, p = _oracle(cfg), port.Port(cfg, 50)
    g = ts.load_configuration(cfg, 50)
    assert ts.get_total_frame_size() == o.total_frame_size
    rng = np.random.default_rng(300 + cfg)
    pl = rng.integers(0, 256, (3, g["frame_bytes"])).astype(np.uint8)
    starts = np.array([0, 777, 123456789], np.uint64)
    out, cw = ts.transmit_byte_batch(pl, starts, want_codeword=True)
    for i in range(3):
        want, after = o.transmit_byte2(pl[i], int(starts[i]))
        _, aux = p.tx_baseband(pl[i], want_aux=True)
        assert np.array_equal(cw[i], aux["codeword"].astype(np.uint8)), (cfg, i)
        assert np.abs(out[i] - want).max() <= 1e-9 * np.abs(want).max(), (cfg, i)
    one, after = ts.transmit_byte([int(v) for v in pl[0]])  # default counter = a freshly initialised reference object (0 in MFSK modes)
    assert after == o.total_frame_size and np.abs(one - out[0]).max() == 0


@pytest.mark.parametrize("cfg,n", [(101, 48), (100, 24)])
def test_robust_tx_channel_rx_round_trip_on_the_device(ts, cfg, n):
    import torch
    dev = torch.device("cuda", 0)
    g = ts.load_configuration(cfg, 50)
    fb, L, buf = g["frame_bytes"], ts.get_total_frame_size(), ts.get_capture_samples()
    gen = torch.Generator(device=dev)
    gen.manual_seed(99 + cfg)
    d_pl = torch.randint(0, 256, (n, fb), device=dev, dtype=torch.uint8, generator=gen)
    d_tx = torch.empty((n, L), device=dev, dtype=torch.float32)
    ts.transmit_byte_batch_device(d_pl, None, n, d_tx, mb.SAMPLES_F32, stream=torch.cuda.current_stream().cuda_stream)
    symb = torch.randint(6, buf // 1088 - (g["Nsymb"] + 4) - 2, (n,), device=dev, generator=gen)
    delays = symb * 1088 + torch.randint(0, 60, (n,), device=dev, generator=gen)
    caps = torch.randn((n, buf), device=dev, dtype=torch.float32, generator=gen) * 0.05
    caps[torch.arange(n, device=dev)[:, None], delays[:, None] + torch.arange(L, device=dev)[None, :]] += d_tx
    d_st = torch.from_numpy(mb.new_receive_stats(n).view(np.uint8).reshape(n, -1)).to(dev)
    d_out = torch.zeros((n, fb), device=dev, dtype=torch.uint8)
    ts.receive_byte_batch_device(caps, mb.SAMPLES_F32, n, d_out, d_st, stream=torch.cuda.current_stream().cuda_stream)
    st = d_st.cpu().numpy().view(mb.RECEIVE_STATS_DTYPE).reshape(-1)
    assert int(st["message_decoded"].sum()) == n and torch.equal(d_out, d_pl)
    assert np.array_equal(st["delay"], (symb * 1088).cpu().numpy())  # the tone-preamble sync works on the symbol grid


@pytest.mark.parametrize("cfg", [100, 101, 102])
def test_robust_receive_byte_against_reference(ts, cfg):
    """The MFSK branch of the whole receive_byte(): pass-band captures holding the reference's own transmit_byte frame (clean, noisy,
    very noisy, running past the end of the buffer, restricted search start) -- delay, verdict, payload, overflow count exact."""
    if not ref.available():
        pytest.skip("frames come from the reference's transmit_byte (oracle/_ref not on this box)")
    r = ref.Ref(cfg, 50)
    ts.load_configuration(cfg, 50)
    n = r.capture_samples()
    assert ts.get_capture_samples() == n
    rng = np.random.default_rng(cfg)
    caps, wants, states = [], [], mb.new_receive_stats(5)
    for case, sigma in enumerate((1e-4, 0.05, 0.3, 0.01, 0.01)):
        pl = rng.integers(0, 256, r.frame_bytes)
        tx = r.transmit_byte(pl)
        d = int(rng.integers(6, r.buffer_Nsymb - (r.Nsymb + 4) - 2)) * 1088 + int(rng.integers(0, 60))
        if case == 3:
            d = (r.buffer_Nsymb - (r.Nsymb + 4) + 3) * 1088
        L = min(tx.size, n - d)
        cap = np.zeros(n)
        cap[d:d + L] += tx[:L]
        cap = (cap + rng.normal(0, sigma, n)).astype(np.float32)
        start = 3 if case == 4 else 0
        states["mfsk_search_or_overflow"][case] = start
        caps.append(cap)
        wants.append((r.receive_byte2(cap.astype(np.float64), search_start_symb=start), pl))
    payload, st, _ = ts.receive_byte_batch(np.stack(caps), states)
    n_dec = 0
    for i, (o, pl) in enumerate(wants):
        assert int(st["delay"][i]) == o["delay"] and int(st["message_decoded"][i]) == o["decoded"], (cfg, i)
        assert int(st["sync_trials"][i]) == o["sync_trials"] and int(st["iterations_done"][i]) == o["iterations"], (cfg, i)
        assert int(st["crc"][i]) == o["crc"] and int(st["all_zeros"][i]) == o["all_zeros"]
        assert int(st["mfsk_search_or_overflow"][i]) == o["frame_overflow_symbols"], (cfg, i)
        assert int(st["delay_of_last_decoded_message"][i]) == o["last_delay"]
        assert float(st["SNR"][i]) == pytest.approx(o["snr"], abs=1e-6)
        assert abs(float(st["signal_stregth_dbm"][i]) - o["signal_dbm"]) <= 1e-9
        assert np.array_equal(payload[i].astype(np.int32), o["payload"]), (cfg, i)
        if o["decoded"]:
            assert np.array_equal(payload[i], np.asarray(pl, np.uint8))
            n_dec += 1
    assert n_dec == 4 and int(st["mfsk_search_or_overflow"][3]) == 3


@pytest.mark.parametrize("cfg", [100, 101, 102])
def test_robust_fixed_delay_overflow_recapture(ts, cfg):
    """mfsk_fixed_delay (telecom_system.h:110, .cc:663-673) = MERCURY_B200_MFSK_FIXED_DELAY(d) in the record: the ARQ layer's overflow recapture
    (arq_common.cc:2815-2841).  Call 1 reports frame_overflow_symbols; the caller shifts the buffer by overflow + 4 symbols and calls again with
    the known delay.  One batch mixes searching and fixed-delay captures; every field equals the unmodified reference's."""
    if not ref.available():
        pytest.skip("frames come from the reference's transmit_byte (oracle/_ref not on this box)")
    r = ref.Ref(cfg, 50)
    ts.load_configuration(cfg, 50)
    n, sym = r.capture_samples(), 1088
    rng = np.random.default_rng(300 + cfg)
    pl = rng.integers(0, 256, r.frame_bytes)
    tx = r.transmit_byte(pl)
    d = (r.buffer_Nsymb - (r.Nsymb + 4) + 3) * sym
    stream = np.zeros(n + 16 * sym)
    stream[d:d + tx.size] += tx
    stream = (stream + rng.normal(0, 0.01, stream.size)).astype(np.float32)
    payload, st, _ = ts.receive_byte_batch(stream[None, :n])
    o = r.receive_byte2(stream[:n].astype(np.float64))
    assert int(st["mfsk_search_or_overflow"][0]) == o["frame_overflow_symbols"] == 3 and int(st["delay"][0]) == o["delay"] == d
    shift = 3 + 4
    fixed = max(d - shift * sym, 0)
    cap2 = stream[shift * sym:shift * sym + n]
    cases = [fixed, fixed + 3, fixed + 5 * sym, 2 * sym, 0, n - 10 * sym, -1]  # -1: the same capture with the ordinary search
    states = mb.new_receive_stats(len(cases))
    for i, fd in enumerate(cases):
        states["mfsk_search_or_overflow"][i] = mb.mfsk_fixed_delay(fd) if fd >= 0 else 0
    payload, st, _ = ts.receive_byte_batch(np.stack([cap2] * len(cases)), states)
    n_dec = 0
    for i, fd in enumerate(cases):
        o = r.receive_byte2(cap2.astype(np.float64), mfsk_fixed_delay=fd)
        assert int(st["delay"][i]) == o["delay"] and int(st["message_decoded"][i]) == o["decoded"], (cfg, fd)
        assert int(st["sync_trials"][i]) == o["sync_trials"] and int(st["iterations_done"][i]) == o["iterations"], (cfg, fd)
        assert int(st["crc"][i]) == o["crc"] and int(st["all_zeros"][i]) == o["all_zeros"]
        assert int(st["mfsk_search_or_overflow"][i]) == o["frame_overflow_symbols"], (cfg, fd)  # the fixed delay is consumed, never echoed
        assert int(st["delay_of_last_decoded_message"][i]) == o["last_delay"]
        assert float(st["SNR"][i]) == pytest.approx(o["snr"], abs=1e-6)
        if fd >= 0:
            assert float(st["signal_stregth_dbm"][i]) == o["signal_dbm"] == 0.0
        else:
            assert abs(float(st["signal_stregth_dbm"][i]) - o["signal_dbm"]) <= 1e-9
        assert np.array_equal(payload[i].astype(np.int32), o["payload"]), (cfg, fd)
        if o["decoded"]:
            assert np.array_equal(payload[i], np.asarray(pl, np.uint8))
            n_dec += 1
    assert n_dec == 3  # the exact delay, three samples late (inside the guard interval), and the ordinary search


@pytest.mark.parametrize("cfg", [8, 16, 101])
def test_arq_tone_pattern_calls_in_any_configuration(ts, cfg):
    """generate_ack/break_pattern_passband and detect_ack/break_pattern_from_passband (telecom_system.h:122-130): the ARQ layer's
    acknowledgement tones, config independent (dedicated 16-MFSK plan), here in an OFDM and a ROBUST configuration."""
    o = _oracle(cfg)
    ts.load_configuration(cfg, 50)
    rng = np.random.default_rng(40 + cfg)
    pats = {}
    for brk in (False, True):
        got, after = ts.generate_pattern_passband(brk, 4321)
        want, wafter = o.generate_pattern_passband(brk, 4321)
        assert after == wafter == 4321 + 16 * 1088 and got.size == want.size
        assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max(), (cfg, brk)
        pats[brk] = want
    n = 40 * 1088
    bufs = []
    for kind in ("ack", "break", "noise", "ack_weak"):
        b = rng.normal(0, 0.05, n)
        pos = int(rng.integers(2, 20)) * 1088 + int(rng.integers(0, 50))
        if kind != "noise":
            b[pos:pos + 16 * 1088] += pats[kind == "break"] * (0.15 if kind == "ack_weak" else 1.0)
        bufs.append(b.astype(np.float32).astype(np.float64))
    bufs = np.stack(bufs)
    res = ts.detect_patterns_from_passband_batch(bufs)
    for i in range(len(bufs)):
        for brk, name in ((False, "ack"), (True, "break")):
            m, matched = o.detect_pattern_from_passband(bufs[i], brk)
            assert abs(float(res[name + "_metric"][i]) - m) <= 1e-9 * max(1.0, m), (cfg, i, name)
            assert int(res[name + "_matched"][i]) == matched, (cfg, i, name)
    assert res["ack_metric"][0] > 12 and res["break_metric"][1] > 12 and res["ack_metric"][2] < 6
    res32 = ts.detect_patterns_from_passband_batch(bufs.astype(np.float32))  # float32 captures: the same values
    assert res32.tobytes() == res.tobytes()


@pytest.mark.parametrize("cfg", [100, 101, 102])
def test_mfsk_control_frames(ts, cfg):
    """set_mfsk_ctrl_mode(true): shortened control frames (ROBUST_0 240 of 320 symbols, ROBUST_1 175 of 200, ROBUST_2 unchanged) through
    transmit_byte, the tail and the whole receive_byte, against the oracle in the same mode."""
    o, p = _oracle(cfg), port.Port(cfg, 50)
    g = ts.load_configuration(cfg, 50)
    na = ts.set_mfsk_ctrl_mode(True)
    assert na == o.set_mfsk_ctrl_mode(True) == p.set_mfsk_ctrl_mode(True) == {100: 240, 101: 175, 102: 200}[cfg] == ts.get_active_nsymb()
    rng = np.random.default_rng(70 + cfg)
    pl = rng.integers(0, 256, g["frame_bytes"]).astype(np.uint8)
    # tail: base-band frames whose symbols past the active ones are silence
    x = o.tx_baseband(pl)
    x[na * 272:] = 0
    xs = np.stack([(x + sg * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))).astype(np.complex64) for sg in (0.0, 20.0, 45.0, 70.0)])
    payload, st, llr = ts.demod_decode_batch(xs.reshape(4, g["Nsymb"], 272), want_llr=True)
    for i in range(4):
        r = o.rx_tail(xs[i].astype(np.complex128))
        tol = 1e-4 * np.maximum(np.abs(r["llr_cw"]), np.median(np.abs(r["llr_cw"])) + 1e-3)
        assert (np.abs(llr[i] - r["llr_cw"]) <= tol).all(), (cfg, i)
        assert int(st["message_decoded"][i]) == r["decoded"] and int(st["iterations_done"][i]) == r["iterations"], (cfg, i)
        if r["decoded"]:
            assert np.array_equal(payload[i], pl)
    # transmit: only the active symbols are modulated
    tx = ts.transmit_byte_batch(pl[None, :], np.array([0], np.uint64))[0]
    want, _ = o.transmit_byte2(pl, 0)
    L = (4 + na) * 1088
    # the reference leaves whatever its TX buffer held before behind the active part (stale samples of earlier frames, which its
    # FIRs also smear over the last 96 active samples); here silence follows.  Compare the part that is defined.
    assert np.abs(tx[:L - 100] - want[:L - 100]).max() <= 1e-9 * np.abs(want).max()
    assert L + 200 >= tx.size or np.abs(tx[L + 200:]).max() <= 1e-12
    # the running carrier counter advances by (preamble + ACTIVE symbols) * 1088 (telecom_system.cc:531-532, ofdm.cc:2313): three frames in a
    # row through the single-frame entry point, counter and samples against the oracle's, which is called the same way
    cnt_gpu = cnt_ref = 12345
    for k in range(3):
        plk = rng.integers(0, 256, g["frame_bytes"]).astype(np.uint8)
        got, cnt_gpu = ts.transmit_byte([int(v) for v in plk], cnt_gpu)
        exp, cnt_ref = o.transmit_byte2(plk, cnt_ref)
        assert cnt_gpu == cnt_ref == 12345 + (k + 1) * (4 + na) * 1088, (cfg, k, cnt_gpu, cnt_ref)
        assert np.abs(got[:L - 100] - exp[:L - 100]).max() <= 1e-9 * np.abs(exp).max(), (cfg, k)
    # the whole receive_byte on a capture holding that control frame
    if ref.available():
        n = o.capture_samples()
        cap = np.zeros(n)
        d = 30 * 1088 + 21
        cap[d:d + L - 100] += want[:L - 100]
        cap = (cap + rng.normal(0, 0.05, n)).astype(np.float32)
        rr = o.receive_byte2(cap.astype(np.float64))
        pay, rs, _ = ts.receive_byte_batch(cap)
        assert int(rs["delay"][0]) == rr["delay"] and int(rs["message_decoded"][0]) == rr["decoded"] == 1
        assert int(rs["iterations_done"][0]) == rr["iterations"] and np.array_equal(pay[0], pl)
    # load_configuration switches the mode off, like the reference (telecom_system.cc:2989)
    ts.load_configuration(cfg, 50)
    assert ts.get_active_nsymb() == g["Nsymb"]
