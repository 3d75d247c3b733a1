"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the oracle on identical inputs.

Bar (BASELINE.json north_star): bit-exact decoded payload + CRC against the reference C++ path, LLRs within
1e-4 relative.  The tolerance used for LLRs is |d| <= 1e-4 * max(|LLR_ref|, median|LLR_ref|) (SURVEY.md 7:
an LLR is a difference of two distances and may be arbitrarily close to 0).  For the ZF modes 15/16 the
reference's noise variance is rounding residue, so raw LLRs are not comparable; what the decoder
can observe -- their signs -- is compared instead.

The oracle is oracle/_ref (the unmodified reference) when that .so travelled to this box, else the C restatement
that is pinned to it bit-exactly (tests/test_oracle_vs_ref.py).
"""
import os

import numpy as np
import pytest

import mercury_b200 as mb
from oracle import port, ref

pytestmark = pytest.mark.gpu

THRESH = mb.THRESH_DB
ITERS = {16: 20}


def oracle_for(cfg, iters):
    return ref.Ref(cfg, iters) if ref.available() else port.Port(cfg, iters)


@pytest.fixture(scope="module")
def ts():
    t = mb.TelecomSystemB200(0)
    yield t
    t.close()


def llr_close(got, want):
    tol = 1e-4 * np.maximum(np.abs(want), np.median(np.abs(want)))
    return np.abs(got - want) <= tol


@pytest.mark.parametrize("cfg", range(17))
def test_golden_vectors_every_stage(ts, cfg, golden_dir):
    """Committed reference vectors: stage tensors, LLRs, payload, CRC, iteration count."""
    import torch
    g = np.load(os.path.join(golden_dir, f"rx_mode{cfg:02d}.npz"))
    geom = ts.load_configuration(cfg, int(g["ldpc_iters"]))
    S = geom["Nsymb"]
    x = torch.from_numpy(g["x"].reshape(1, S, 272)).cuda()
    dY, dH, dZ = (torch.zeros(1, S, 50, dtype=torch.complex64, device="cuda") for _ in range(3))
    d_pay = torch.zeros(1, geom["frame_bytes"], dtype=torch.uint8, device="cuda")
    d_st = torch.zeros(1, 32, dtype=torch.uint8, device="cuda")
    d_llr = torch.zeros(1, 1600, dtype=torch.float32, device="cuda")
    ts.set_debug_capture(dY, dH, dZ)
    try:
        ts.demod_decode_batch_device(x, 1, d_pay, d_st, d_llr, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    finally:
        ts.set_debug_capture(None, None, None)
    st = d_st.cpu().numpy().view(mb.STATS_DTYPE)[0]
    for name, t in (("Y", dY), ("H", dH), ("Z", dZ)):
        got, want = t.cpu().numpy()[0], g[name]
        assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max()), name
    llr = d_llr.cpu().numpy()[0]
    assert abs(st["SNR"] - float(g["snr"])) <= 1e-3  # ZF modes: the re-encode path (telecom_system.cc:1376-1400)
    if cfg < 15:
        assert llr_close(llr, g["llr_cw"]).all()
        assert abs(st["variance"] / float(g["variance"]) - 1) <= 1e-4
    else:
        nz = np.abs(g["llr_cw"]) > 0
        assert np.array_equal(np.signbit(llr[nz]), np.signbit(g["llr_cw"][nz]))
    assert abs(st["mean_H"] - float(g["mean_H"])) <= 1e-4
    assert np.array_equal(d_pay.cpu().numpy()[0], g["rx_payload"]) and np.array_equal(g["rx_payload"], g["payload"].astype(np.uint8))
    assert (st["iterations_done"], st["crc"], st["all_zeros"], st["message_decoded"]) == (
        int(g["iterations"]), int(g["crc"]), int(g["all_zeros"]), int(g["decoded"]))


def test_config1_loopback_frame_through_receive_baseband(ts, golden_dir):
    """BASELINE config #1: the frame the reference's own receive_byte() synchronised out of a TX_TEST pass-band capture."""
    g = np.load(os.path.join(golden_dir, "loopback_mode00.npz"))
    ts.load_configuration(0, 50)
    out, st = ts.receive_baseband(g["x"].astype(np.complex128))
    assert out.dtype == np.int32 and np.array_equal(out, g["payload"]) and np.array_equal(out.astype(np.uint8), g["rx_payload_receive_byte"])
    assert st["message_decoded"] == mb.YES == int(g["decoded_receive_byte"]) and st["crc"] == 0
    assert st["iterations_done"] == int(g["iterations"]) == int(g["iterations_receive_byte"])
    assert abs(st["SNR"] - float(g["snr"])) < 0.01


@pytest.mark.parametrize("cfg", range(17))
def test_seeded_batch_vs_oracle(ts, cfg):
    """Fresh seeded inputs around the operating point; host-buffer batch path; every frame compared with the oracle."""
    iters = ITERS.get(cfg, 50)
    geom = ts.load_configuration(cfg, iters)
    o = oracle_for(cfg, iters)
    n_per = 24
    offs = (2.0, 0.5, -1.5) if cfg < 15 else (16.0, 11.5, 8.0)
    xs, pls = [], []
    for i, off in enumerate(offs):
        x, pl = mb.synth_frames(cfg, n_per, seed=100 * cfg + i, esn0_db=THRESH[cfg] + off)
        xs.append(x), pls.append(pl)
    x, pl = np.concatenate(xs), np.concatenate(pls)
    payload, stats, llr = ts.demod_decode_batch(x, want_llr=True)
    n_dec = n_same_iter = n_llr_ok = 0
    for f in range(x.shape[0]):
        r = o.rx_tail(x[f].reshape(-1).astype(np.complex128))
        if cfg < 15:
            n_llr_ok += int(llr_close(llr[f], r["llr_cw"]).all())
        else:
            nz = np.abs(r["llr_cw"]) > 0
            n_llr_ok += int(np.array_equal(np.signbit(llr[f][nz]), np.signbit(r["llr_cw"][nz])))
        assert stats["message_decoded"][f] == r["decoded"], (f, stats[f], r["iterations"])
        if r["decoded"]:
            n_dec += 1
            assert np.array_equal(payload[f], r["payload"].astype(np.uint8)) and np.array_equal(payload[f], pl[f])
            assert stats["crc"][f] == 0 == r["crc"] and stats["all_zeros"][f] == 0
            assert abs(stats["SNR"][f] - r["snr"]) < 2e-3
        else:
            assert stats["SNR"][f] == np.float32(-99.9)
        n_same_iter += int(stats["iterations_done"][f] == r["iterations"])
    n = x.shape[0]
    assert n_llr_ok == n
    assert n_dec >= n_per  # the +2 dB third decodes
    # the fp32 decoder must take the reference's (double-precision) number of iterations on every frame, including the ones that do not converge
    assert n_same_iter == n, f"iteration counts equal on {n_same_iter}/{n} frames"


@pytest.mark.parametrize("fmt", ["int16", "float16"])
@pytest.mark.parametrize("cfg", [0, 8, 11, 13, 16])
def test_narrow_sample_formats_vs_oracle(ts, cfg, fmt):
    """mercury_b200_demod_decode_batch_fmt: complex int16 / fp16 base-band samples (half the PCIe bytes of complex64).  The kernel widens
    them in its load, so the result must be the reference's on exactly those quantised values (fed to it as doubles)."""
    iters = ITERS.get(cfg, 50)
    ts.load_configuration(cfg, iters)
    o = oracle_for(cfg, iters)
    x, pl = mb.synth_frames(cfg, 24, seed=900 + cfg, esn0_db=THRESH[cfg] + (1.0 if cfg < 15 else 14.0))
    xf = x.view(np.float32).reshape(x.shape + (2,))
    if fmt == "int16":
        scale = float(np.abs(xf).max()) / 32000.0
        q = np.rint(xf / scale).astype(np.int16)
        wide = (q.astype(np.float32) * np.float32(scale)).astype(np.float64)   # the kernel's own widening: float(int) x float32 scale, in fp32
        payload, stats, llr = ts.demod_decode_batch(q, want_llr=True, scale=scale)
    else:
        q = xf.astype(np.float16)
        wide = q.astype(np.float64)
        payload, stats, llr = ts.demod_decode_batch(q, want_llr=True)
    xr = wide[..., 0] + 1j * wide[..., 1]
    for f in range(x.shape[0]):
        r = o.rx_tail(xr[f].reshape(-1))
        if cfg < 15:
            assert llr_close(llr[f], r["llr_cw"]).all()
        else:
            nz = np.abs(r["llr_cw"]) > 0
            assert np.array_equal(np.signbit(llr[f][nz]), np.signbit(r["llr_cw"][nz]))
        assert stats["message_decoded"][f] == r["decoded"] and stats["iterations_done"][f] == r["iterations"]
        if r["decoded"]:
            assert np.array_equal(payload[f], r["payload"].astype(np.uint8)) and np.array_equal(payload[f], pl[f])
    assert (stats["message_decoded"] == 1).sum() >= 12
    # the device entry point with the same buffers
    import torch
    geom = ts.geometry
    d_q = torch.from_numpy(q).cuda()
    d_pay = torch.zeros(x.shape[0], geom["frame_bytes"], dtype=torch.uint8, device="cuda")
    d_st = torch.zeros(x.shape[0], 32, dtype=torch.uint8, device="cuda")
    ts.demod_decode_batch_device(d_q, x.shape[0], d_pay, d_st, None, stream=torch.cuda.current_stream().cuda_stream,
                                 sample_format=mb.BASEBAND_CI16 if fmt == "int16" else mb.BASEBAND_CF16, scale=scale if fmt == "int16" else 1.0)
    torch.cuda.synchronize()
    assert np.array_equal(d_pay.cpu().numpy(), payload) and np.array_equal(d_st.cpu().numpy().view(mb.STATS_DTYPE).reshape(-1), stats)
    with pytest.raises(mb.MercuryB200Error):   # the ROBUST (MFSK) tail takes complex64 only
        ts.load_configuration(100, 50)
        ts.demod_decode_batch(np.zeros((1, 320, 272, 2), np.int16))
    ts.load_configuration(8, 50)


def test_shards_decode_like_the_whole_batch(ts):
    """Multi-GPU partitioning (mercury_b200/dist.py shard_range): a batch decoded as contiguous shards -- what each rank of a sharded
    job does -- gives exactly the whole-batch payloads and records, for every world size the bench runs."""
    from mercury_b200.dist import shard_range
    ts.load_configuration(8, 50)
    n = 3001
    x, pl = mb.synth_frames(8, n, seed=77, esn0_db=THRESH[8] + 0.3)   # a few frames fail here: both outcomes are compared
    payload, stats, _ = ts.demod_decode_batch(x)
    assert 0 < (stats["message_decoded"] == 0).sum() < n // 2
    for world in (2, 4, 8):
        got_p, got_s = [], []
        for rank in range(world):
            lo, hi = shard_range(n, rank, world)
            p, s, _ = ts.demod_decode_batch(x[lo:hi])
            got_p.append(p), got_s.append(s)
        assert np.array_equal(np.concatenate(got_p), payload) and np.array_equal(np.concatenate(got_s), stats)


def test_large_batch_properties_mode8(ts):
    """BASELINE config #2 shape at test size: size-independent properties over a whole batch (round trip, CRC invariant,
    batch-order independence, chunk pipeline == single launch)."""
    geom = ts.load_configuration(8, 50)
    n = 6000  # > 2 pipeline chunks of the host path
    x, pl = mb.synth_frames(8, n, seed=7, esn0_db=THRESH[8] + 2.0)
    payload, stats, _ = ts.demod_decode_batch(x)
    dec = stats["message_decoded"] == 1
    assert dec.mean() > 0.995
    assert np.array_equal(payload[dec], pl[dec])                      # encode -> channel -> decode round trip
    assert (stats["crc"][dec] == 0).all() and (stats["all_zeros"] == 0).all()
    assert (stats["iterations_done"][dec] <= 50).all() and (stats["iterations_done"][~dec] >= 0).all()
    assert (stats["SNR"][~dec] == np.float32(-99.9)).all()
    assert abs(float(np.median(stats["SNR"][dec])) - (THRESH[8] + 2.0)) < 1.0
    perm = np.random.default_rng(0).permutation(n)[:1500]
    p2, s2, _ = ts.demod_decode_batch(x[perm])
    assert np.array_equal(p2, payload[perm]) and np.array_equal(s2, stats[perm])  # frames are independent; results deterministic


@pytest.mark.parametrize("cfg", [8, 16])
def test_gated_frames_inside_a_busy_queue(ts, cfg):
    """Frames the demodulator gates out (mean|H| < 0.3, telecom_system.cc:1268-1280) in the middle of a batch larger than the decoder's
    resident slots: they are skipped by the refill of a slot whose previous frame has just been finished (epilogue and refill share one
    function in the LS modes; the ZF modes run their epilogue first), several in a row included, and every frame around them is decoded as if they were not there."""
    geom = ts.load_configuration(cfg, 50)
    S = geom["Nsymb"]
    n = 3000
    x, pl = mb.synth_frames(cfg, n, seed=21, esn0_db=THRESH[8] + 2.0 if cfg == 8 else 32.0)  # (ZF modes decode by hard decision: SURVEY.md 7)
    x = np.ascontiguousarray(x)
    rng = np.random.default_rng(5)
    noise_at = np.zeros(n, bool)
    noise_at[::5] = True
    noise_at[1001:1009] = True          # a run of gated frames: the refill loops over them
    noise_at[-3:] = True                # and the queue ends on gated frames
    k = int(noise_at.sum())
    x[noise_at] = (0.05 * (rng.standard_normal((k, S, 272)) + 1j * rng.standard_normal((k, S, 272)))).astype(np.complex64)
    p, s, _ = ts.demod_decode_batch(x)
    gated = s["mean_H"] < 0.3
    if cfg == 8:  # LS estimate of noise: weak.  (ZF: H = Y / p under the AGC never is -- there the noise frames run all iterations and fail.)
        assert gated[noise_at].all()
    assert not gated[~noise_at].any()
    assert (s["iterations_done"][gated] == -1).all() and (s["message_decoded"][gated] == 0).all() and not p[gated].any()
    assert (s["crc"][gated] == 0).all() and (s["all_zeros"][gated] == 0).all() and np.allclose(s["SNR"][gated], -99.9)
    good = ~noise_at
    assert (s["message_decoded"][good] == 1).mean() > 0.99
    dec = good & (s["message_decoded"] == 1)
    assert np.array_equal(p[dec], pl[dec])
    sub = np.r_[995:1015, n - 8:n]      # the same frames in a batch of their own: identical records
    p2, s2, _ = ts.demod_decode_batch(x[sub])
    assert np.array_equal(p2, p[sub]) and np.array_equal(s2, s[sub])


def test_edge_cases(ts):
    geom = ts.load_configuration(8, 50)
    S, fb = geom["Nsymb"], geom["frame_bytes"]
    # empty batch
    p, s, _ = ts.demod_decode_batch(np.zeros((0, S, 272), np.complex64))
    assert p.shape == (0, fb) and s.shape == (0,)
    # ragged sizes around the launch / chunk boundaries give the same per-frame answers
    x, pl = mb.synth_frames(8, 37, seed=9, esn0_db=4.0)
    full = ts.demod_decode_batch(x)
    for k in (1, 2, 3, 36):
        part = ts.demod_decode_batch(x[:k])
        assert np.array_equal(part[0], full[0][:k]) and np.array_equal(part[1], full[1][:k])
    # noise only: the reference skips the decode when mean|H| < 0.3 (telecom_system.cc:1268-1280)
    rng = np.random.default_rng(3)
    noise = (rng.standard_normal((4, S, 272)) + 1j * rng.standard_normal((4, S, 272))).astype(np.complex64)
    p, s, _ = ts.demod_decode_batch(noise)
    o = oracle_for(8, 50)
    for f in range(4):
        r = o.rx_tail(noise[f].reshape(-1).astype(np.complex128))
        assert abs(s["mean_H"][f] - r["mean_H"]) < 1e-4
        if r["mean_H"] < 0.3:
            assert s["iterations_done"][f] == -1 and s["message_decoded"][f] == 0 and not p[f].any()
    # an all-zero decoded frame is rejected like the reference does (telecom_system.cc:1319-1327,1343)
    geom = ts.load_configuration(0, 50)
    scr_free = np.zeros((1, geom["frame_bytes"]), np.uint8)
    x0, _ = mb.synth_frames(0, 1, seed=1, esn0_db=300.0, payload=scr_free)
    p, s, _ = ts.demod_decode_batch(x0)
    r = oracle_for(0, 50).rx_tail(x0[0].reshape(-1).astype(np.complex128))
    assert (s["all_zeros"][0], s["message_decoded"][0], s["crc"][0]) == (r["all_zeros"], r["decoded"], r["crc"])
    # O(1) mode switching with all tables resident (ARQ flips data/ack configurations every batch)
    for cfg in (16, 0, 13, 8, 10):
        gg = ts.load_configuration(cfg, 50)
        xx, pp = mb.synth_frames(cfg, 5, seed=cfg, esn0_db=300.0)
        got, st, _ = ts.demod_decode_batch(xx)
        assert np.array_equal(got, pp) and (st["iterations_done"] == 0).all() and (st["message_decoded"] == 1).all()
    # error behaviour of the boundary
    with pytest.raises(mb.MercuryB200Error):
        ts.load_configuration(17)
    assert ts.load_configuration(8, 500)["ldpc_iters"] == 50 and ts.load_configuration(8, 1)["ldpc_iters"] == 5  # main.cc:303-311


@pytest.mark.parametrize("cfg", [0, 6, 8, 12, 13])
def test_minsum_decoder_agrees_where_both_converge(ts, cfg):
    """north_star's min-sum decoder: identical payloads wherever it and the reference's SPA both converge."""
    ts.load_configuration(cfg, 50)
    x, pl = mb.synth_frames(cfg, 256, seed=55 + cfg, esn0_db=THRESH[cfg] + 2.5)
    ts.set_decoder(mb.DECODER_SPA)
    p_spa, s_spa, _ = ts.demod_decode_batch(x)
    ts.set_decoder(mb.DECODER_MINSUM)
    try:
        p_ms, s_ms, _ = ts.demod_decode_batch(x)
    finally:
        ts.set_decoder(mb.DECODER_SPA)
    both = (s_spa["message_decoded"] == 1) & (s_ms["message_decoded"] == 1)
    assert both.mean() > 0.9
    assert np.array_equal(p_spa[both], p_ms[both]) and np.array_equal(p_ms[both], pl[both])


def test_ldpc_stage_alone_rate_sweep(ts):
    """BASELINE config #4 shape: the decoder stage by itself on all 8 rates (BPSK geometry + mode 12 for 14/16)."""
    import torch
    for cfg in (0, 1, 2, 3, 4, 5, 6, 12):
        geom = ts.load_configuration(cfg, 50)
        o = oracle_for(cfg, 50)
        x, pl = mb.synth_frames(cfg, 16, seed=900 + cfg, esn0_db=THRESH[cfg] + 1.5)
        dx = torch.from_numpy(x).cuda()
        d_llr = torch.zeros(16, mb.HANDOFF_FLOATS, device="cuda")
        d_st = torch.zeros(16, 32, dtype=torch.uint8, device="cuda")
        d_pay = torch.zeros(16, geom["frame_bytes"], dtype=torch.uint8, device="cuda")
        s = torch.cuda.current_stream().cuda_stream
        ts.demod_batch_device(dx, 16, d_llr, d_st, None, stream=s)
        ts.ldpc_decode_batch_device(d_llr, 16, d_pay, d_st, stream=s)
        torch.cuda.synchronize()
        st = d_st.cpu().numpy().view(mb.STATS_DTYPE).reshape(-1)
        for f in range(16):
            r = o.rx_tail(x[f].reshape(-1).astype(np.complex128))
            assert st["message_decoded"][f] == r["decoded"]
            if r["decoded"]:
                assert np.array_equal(d_pay[f].cpu().numpy(), r["payload"].astype(np.uint8))


def test_all_zeros_frame_is_rejected(ts):
    """telecom_system.cc:1319-1327,1343: a frame whose de-scrambled bytes are all zero is dropped (all_zeros = YES, crc = 0).
    Crafted at the decoder stage: saturated LLRs equal to the scrambler sequence (pass-through, like the ZF modes)."""
    import torch
    from tests import blob_emulator as be
    geom = ts.load_configuration(8, 50)
    blob = be.Blob(ts.export_tables())
    m = blob.mode(8)
    r = blob.rate(m["rate_idx"])
    cw = np.full(1600, 1e20, np.float32)
    cw[: geom["nReal"]] = np.where(m["scr"][: geom["nReal"]] == 1, -1e20, 1e20)
    L = np.zeros((1, mb.HANDOFF_FLOATS), np.float32)
    L[0, be.handoff(r["var_of_cw"].astype(int))] = cw  # the kernels' hand-off layout
    st = np.zeros(1, mb.STATS_DTYPE)
    st["mean_H"] = 1.0
    d_llr, d_st = torch.from_numpy(L).cuda(), torch.from_numpy(st.view(np.uint8).reshape(1, 32)).cuda()
    d_pay = torch.full((1, geom["frame_bytes"]), 255, dtype=torch.uint8, device="cuda")
    ts.ldpc_decode_batch_device(d_llr, 1, d_pay, d_st, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = d_st.cpu().numpy().view(mb.STATS_DTYPE)[0]
    o = oracle_for(8, 50)
    it, bits = o.ldpc_decode(cw)
    assert out["all_zeros"] == 1 and out["crc"] == 0 and out["message_decoded"] == 0 and out["SNR"] == np.float32(-99.9)
    assert out["iterations_done"] == it and not d_pay.cpu().numpy().any()
