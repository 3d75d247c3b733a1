"""GPU parity of the TX chain row: mercury_b200_transmit_byte(_batch) against the oracle's transmit_byte
(the unmodified reference when oracle/_ref travelled to this box, else the C restatement) and the committed reference fixture.

Bars: LDPC codeword bits exact; pass-band samples within 1e-9 of the frame's peak (fp64 chain, different operation order);
float32 output = the float64 output rounded; and the size-independent property: frames made by the GPU TX chain, dropped at random
delays into noisy capture buffers, come back bit-exact through the GPU receive_byte() (TX -> channel -> RX entirely on the device)."""
import os

import numpy as np
import pytest

import mercury_b200 as mb
from oracle import port, ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ts():
    t = mb.TelecomSystemB200(0)
    yield t
    t.close()


def _oracle(cfg):
    return ref.Ref(cfg, 50) if ref.available() else port.Port(cfg, 50)


@pytest.mark.parametrize("cfg", list(range(17)))
def test_transmit_byte_batch_against_oracle(ts, cfg):
    o, p = _oracle(cfg), port.Port(cfg, 50)
    g = ts.load_configuration(cfg, 50)
    assert ts.get_total_frame_size() == o.total_frame_size
    rng = np.random.default_rng(77 + cfg)
    n = 6
    pl = rng.integers(0, 256, (n, g["frame_bytes"])).astype(np.uint8)
    pl[1, -3:] = 0  # what a short message looks like after the zero padding
    starts = np.array([1088, 0, 123457, 987654321, 5, 4096], np.uint64)
    out, cw = ts.transmit_byte_batch(pl, starts, want_codeword=True)
    for i in range(n):
        want, after = o.transmit_byte2(pl[i], int(starts[i]))
        assert after == int(starts[i]) + o.total_frame_size
        _, aux = p.tx_baseband(pl[i], want_aux=True)
        assert np.array_equal(cw[i], aux["codeword"].astype(np.uint8)), (cfg, i)
        err = np.abs(out[i] - want).max() / np.abs(want).max()
        assert err <= 1e-9, (cfg, i, err)
    out32 = ts.transmit_byte_batch(pl, starts, dtype=np.float32)
    assert np.abs(out32 - out.astype(np.float32)).max() <= 1e-6 * np.abs(out).max()
    # the single call in the reference's own types: int* data, nBytes shorter than the frame, counter advanced like ofdm.passband_start_sample
    short = [int(v) for v in pl[2, :g["frame_bytes"] - 2]]
    one, after = ts.transmit_byte(short, 777)
    want, wafter = o.transmit_byte2(short, 777)
    assert after == wafter and np.abs(one - want).max() <= 1e-9 * np.abs(want).max()
    fresh, _ = ts.transmit_byte(short)  # default counter = a freshly initialised reference object
    assert np.abs(fresh - o.transmit_byte2(short, 1088)[0]).max() <= 1e-9 * np.abs(want).max()


def test_transmit_byte_reference_fixture(ts, golden_dir):
    g = np.load(os.path.join(golden_dir, "tx_mode16.npz"))
    ts.load_configuration(16, 50)
    out, after = ts.transmit_byte([int(v) for v in g["payload"]], int(g["start_sample"]))
    assert after == int(g["start_sample_after"])
    assert np.abs(out - g["passband"]).max() <= 1e-9 * np.abs(g["passband"]).max()


@pytest.mark.parametrize("cfg,n", [(8, 1024), (16, 512), (0, 256)])
def test_tx_channel_rx_round_trip_on_the_device(ts, cfg, n):
    """Property test at batch size: GPU TX -> random delay + white noise (torch, on the device) -> GPU receive_byte()."""
    import torch
    dev = torch.device("cuda", 0)
    g = ts.load_configuration(cfg, 50)
    fb, L, buf = g["frame_bytes"], ts.get_total_frame_size(), ts.get_capture_samples()
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + cfg)
    d_pl = torch.randint(0, 256, (n, fb), device=dev, dtype=torch.uint8, generator=gen)
    d_start = torch.randint(0, 1 << 40, (n,), device=dev, dtype=torch.int64, generator=gen)
    d_tx = torch.empty((n, L), device=dev, dtype=torch.float32)
    ts.transmit_byte_batch_device(d_pl, d_start, n, d_tx, mb.SAMPLES_F32, stream=torch.cuda.current_stream().cuda_stream)
    lo, hi = (g["preamble_nSymb"] + 1) * 1088 + 10, buf - L - 2000
    delays = torch.randint(lo, hi, (n,), device=dev, generator=gen)
    # mode 16 (32QAM, rate 14/16, zero-forcing) only works as a hard-decision pass-through: keep its channel nearly clean
    caps = torch.randn((n, buf), device=dev, dtype=torch.float32, generator=gen) * (0.0005 if cfg == 16 else 0.01)
    caps[torch.arange(n, device=dev)[:, None], delays[:, None] + torch.arange(L, device=dev)[None, :]] += d_tx
    d_st = torch.from_numpy(mb.new_receive_stats(n).view(np.uint8).reshape(n, -1)).to(dev)
    d_out = torch.zeros((n, fb), device=dev, dtype=torch.uint8)
    ts.receive_byte_batch_device(caps, mb.SAMPLES_F32, n, d_out, d_st, stream=torch.cuda.current_stream().cuda_stream)
    st = d_st.cpu().numpy().view(mb.RECEIVE_STATS_DTYPE).reshape(-1)
    dec = st["message_decoded"] == 1
    # mode 16 is a hard-decision pass-through behind a zero-forcing estimate: a few sub-symbol sync offsets defeat it, in the reference
    # exactly as here (the RX side is bit-identical to the reference, tests/test_gpu_frontend.py); every other mode decodes every frame
    assert dec.mean() >= (0.85 if cfg == 16 else 1.0), dec.mean()
    assert np.array_equal(d_out.cpu().numpy()[dec], d_pl.cpu().numpy()[dec])
    assert int(np.abs(st["delay"][dec] - delays.cpu().numpy()[dec]).max()) <= 64  # within the guard interval (16 samples x 4)


@pytest.mark.parametrize("cfg", [8, 16, 101])
def test_arq_batch_transmit_path(ts, cfg):
    """What the ARQ layer does to send a batch (arq_common.cc:2224-2247): transmit_byte(..., NO_FILTER_MESSAGE) per frame with the carrier
    counter running on, pad one frame on each side, FIR_tx1 + FIR_tx2 over the whole buffer."""
    o = _oracle(cfg)
    g = ts.load_configuration(cfg, 50)
    L = ts.get_total_frame_size()
    rng = np.random.default_rng(cfg)
    pl = rng.integers(0, 256, (3, g["frame_bytes"])).astype(np.uint8)
    starts = np.array([5000 + i * L for i in range(3)], np.uint64)
    raw = ts.transmit_byte_batch(pl, starts, message_location=4)
    want, counter = [], 5000
    for i in range(3):
        w, counter = o.transmit_byte2(pl[i], counter, no_filter=True)
        want.append(w)
        assert np.abs(raw[i] - w).max() <= 1e-9 * np.abs(w).max(), (cfg, i)
    assert counter == 5000 + 3 * L
    batch = np.concatenate([raw[0]] + list(raw) + [raw[2]])
    got = ts.fir_tx_apply(batch)
    ref_out = o.fir_tx_apply(np.concatenate([want[0]] + want + [want[2]]))
    assert np.abs(got - ref_out).max() <= 1e-9 * np.abs(ref_out).max()


@pytest.mark.parametrize("cfg", list(range(17)) + [100, 101, 102])
def test_every_configuration_round_trips_on_the_device(ts, cfg):
    """All 17 OFDM configurations and the 3 ROBUST ones: GPU transmit_byte -> random delay + white noise -> GPU receive_byte returns
    the payloads (every frame; the two zero-forcing modes 15/16 on a nearly clean channel, >= 85 % there, see the mode-16 note above)."""
    import torch
    dev = torch.device("cuda", 0)
    g = ts.load_configuration(cfg, 50)
    n = 16 if cfg >= 100 else 64
    fb, L, buf = g["frame_bytes"], ts.get_total_frame_size(), ts.get_capture_samples()
    gen = torch.Generator(device=dev)
    gen.manual_seed(4000 + cfg)
    d_pl = torch.randint(0, 256, (n, fb), device=dev, dtype=torch.uint8, generator=gen)
    d_tx = torch.empty((n, L), device=dev, dtype=torch.float32)
    ts.transmit_byte_batch_device(d_pl, None, n, d_tx, mb.SAMPLES_F32, stream=torch.cuda.current_stream().cuda_stream)
    lo, hi = (g["preamble_nSymb"] + 2) * 1088 + 10, buf - L - 2200
    delays = torch.randint(lo, hi, (n,), device=dev, generator=gen)
    if cfg >= 100:  # the MFSK tone-preamble sync works on the capture's symbol grid (time_sync_mfsk): frames start within a guard interval of it
        delays = (delays // 1088) * 1088 + torch.randint(0, 60, (n,), device=dev, generator=gen)
    sigma = 0.0005 if cfg in (15, 16) else (0.003 if cfg in (12, 14) else 0.01)  # rate 14/16 leaves little margin for sub-sample sync offsets
    caps = torch.randn((n, buf), device=dev, dtype=torch.float32, generator=gen) * sigma
    caps[torch.arange(n, device=dev)[:, None], delays[:, None] + torch.arange(L, device=dev)[None, :]] += d_tx
    d_st = torch.from_numpy(mb.new_receive_stats(n).view(np.uint8).reshape(n, -1)).to(dev)
    d_out = torch.zeros((n, fb), device=dev, dtype=torch.uint8)
    ts.receive_byte_batch_device(caps, mb.SAMPLES_F32, n, d_out, d_st, stream=torch.cuda.current_stream().cuda_stream)
    st = d_st.cpu().numpy().view(mb.RECEIVE_STATS_DTYPE).reshape(-1)
    dec = st["message_decoded"] == 1
    # short preambles (2 symbols in modes 13-15, 1 in mode 16) and rate 14/16 leave the reference's receiver -- and therefore this one, which
    # reproduces its decisions -- a few per cent of sync-offset losses even on a clean channel
    # (mode 14, 8PSK 14/16 behind a 2-symbol preamble, is the extreme: the unmodified reference decodes 15 of 40 such captures of its own frames)
    assert dec.mean() >= {14: 0.3}.get(cfg, 0.85 if cfg in (12, 13, 15, 16) else 1.0), (cfg, dec.mean())
    assert np.array_equal(d_out.cpu().numpy()[dec], d_pl.cpu().numpy()[dec])


@pytest.mark.parametrize("cfg", [8, 16])
def test_streaming_message_locations(ts, cfg):
    """FIRST / MIDDLE / MIDDLE / FLUSH (TX_TEST, telecom_system.cc:2033-2038,559-594): the three-frame filter buffer lives on the device."""
    o = _oracle(cfg)
    g = ts.load_configuration(cfg, 50)
    o.reset_tx_stream(), ts.reset_tx_stream()
    rng = np.random.default_rng(cfg)
    sr = sg = 1088
    for loc in (0, 1, 1, 2):
        pl = rng.integers(0, 256, g["frame_bytes"])
        want, sr = o.transmit_byte_loc(pl, sr, loc)
        got, sg = ts.transmit_byte_loc(pl, sg, loc)
        assert sg == sr and np.abs(got - want).max() <= 1e-9 * np.abs(want).max(), (cfg, loc)
