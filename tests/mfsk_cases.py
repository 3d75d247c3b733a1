"""Inputs for the MFSK row, shared by the oracle pin test and the GPU parity test."""
import numpy as np


def upsample4(x):
    """Base-band symbols (decimated rate) -> a pass-band-rate base-band buffer by x4 linear interpolation (what the pattern functions
    of ofdm.cc:1969-2186 are handed: baseband_data_interpolated; they read every 4th sample)."""
    n = x.size
    t = np.arange(4 * n) / 4.0
    i0 = np.minimum(np.floor(t).astype(int), n - 2)
    f = t - i0
    return x[i0] + (x[i0 + 1] - x[i0]) * f


def pattern_buffer(o, kind, seed, nsymb=64):
    """A buffer of nsymb symbols (pass-band rate, complex128) holding the ACK / BREAK pattern or an MFSK preamble + data at a
    symbol-aligned position (+ a few samples of offset) in white noise.  kind: 'ack', 'break', 'frame', 'noise'."""
    rng = np.random.default_rng(seed)
    sym = o.Nofdm * 4
    buf = (rng.standard_normal(nsymb * sym) + 1j * rng.standard_normal(nsymb * sym)) * 0.05
    pos = None
    if kind in ("ack", "break"):
        pat = upsample4(o.ack_pattern_baseband(kind == "break")) / 8.0
        pos = int(rng.integers(2, nsymb - 18)) * sym + int(rng.integers(0, 40))
        buf[pos:pos + pat.size] += pat
    elif kind == "frame":
        # preamble tones + the first data symbols of a real frame
        t = o.mfsk_tables()
        pre = np.zeros((4, o.Nc), np.complex128)
        for s in range(4):
            for st in range(t["nStreams"]):
                pre[s, t["stream_offsets"][st] + t["preamble_tones"][s]] = np.sqrt(o.Nc / t["nStreams"])
        x, aux = o.tx_baseband(rng.integers(0, 256, o.frame_bytes), want_aux=True)
        # symbol_mod of the preamble rows through an ACK-pattern-style helper is not exported: build the time signal by an inverse DFT
        k = np.arange(256)
        bins = np.where(np.arange(o.Nc) < o.Nc // 2, np.arange(o.Nc) + 256 - o.Nc // 2, np.arange(o.Nc) - o.Nc // 2 + 1)
        td = []
        for s in range(4):
            v = (pre[s][None, :] * np.exp(2j * np.pi * np.outer(k, bins) / 256)).sum(axis=1)
            td.append(np.concatenate([v[-16:], v]))
        sig = np.concatenate(td + [x[:20 * o.Nofdm]])
        sig = upsample4(sig) / 8.0
        pos = int(rng.integers(2, nsymb - 26)) * sym + int(rng.integers(0, 40))
        buf[pos:pos + sig.size] += sig
    return buf, pos
