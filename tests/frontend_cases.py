"""Capture-buffer scenarios for the RX front-end row: shared by the oracle-vs-reference pin test and the
GPU parity test.  A capture is what receive_byte() is handed: Nofdm*buffer_Nsymb*4 real pass-band samples at 48 kHz
(telecom_system.cc:646, data_container.cc:133-153).  Frames come from the TX side of the object passed in (the unmodified
reference's transmit_byte when oracle/_ref is built)."""
import numpy as np

SYM = 1088  # pass-band samples per OFDM symbol (Nofdm 272 x interpolation 4)


def freq_shift(x, df, fs=48000.0):
    """Shift a real pass-band signal by df Hz through its analytic signal."""
    n = x.size
    X = np.fft.fft(x)
    h = np.zeros(n)
    h[0] = 1
    h[1:(n + 1) // 2] = 2
    if n % 2 == 0:
        h[n // 2] = 1
    a = np.fft.ifft(X * h)
    return np.real(a * np.exp(2j * np.pi * df * np.arange(n) / fs))


CASES = ["clean", "noise_light", "noise_heavy", "freq_offset", "too_early", "late", "weak_frame", "noise_only", "silence",
         "tone_then_frame", "last_good_state", "two_frames", "clean_sym_aligned", "tiny_noise_floor"]


def make_capture(tx, case, seed):
    """-> (capture float64 [n] holding float32-representable values, payload or None, state (last_delay, last_freq)).
    `tx` needs .transmit_byte(payload), .frame_bytes, .preamble_nSymb, .Nsymb, .buffer_Nsymb."""
    rng = np.random.default_rng(seed)
    n = tx.Nofdm * tx.buffer_Nsymb * 4
    pl = rng.integers(0, 256, tx.frame_bytes)
    frame = tx.transmit_byte(pl)
    L = frame.size
    lo, hi = (tx.preamble_nSymb + 1) * SYM + 10, n - L - 2000
    d = int(rng.integers(lo, hi))
    cap = np.zeros(n)
    state = (-1, 0.0)
    if case == "clean":
        cap[d:d + L] += frame
    elif case == "clean_sym_aligned":
        d = (d // SYM) * SYM
        cap[d:d + L] += frame
    elif case == "noise_light":
        cap[d:d + L] += frame
        cap += rng.normal(0, 0.02, n)
    elif case == "tiny_noise_floor":
        cap[d:d + L] += frame
        cap += rng.normal(0, 1e-10, n)
    elif case == "noise_heavy":
        cap[d:d + L] += frame
        cap += rng.normal(0, 0.22, n)
    elif case == "freq_offset":
        cap[d:d + L] += frame
        cap = freq_shift(cap, 6.5) + rng.normal(0, 0.01, n)
    elif case == "too_early":
        d = 100
        cap[d:d + L] += frame
    elif case == "late":
        d = n - L - 3
        cap[d:d + L] += frame
    elif case == "weak_frame":
        cap[d:d + L] += 0.01 * frame
    elif case == "noise_only":
        cap += rng.normal(0, 0.1, n)
        pl = None
    elif case == "silence":
        pl = None
    elif case == "tone_then_frame":
        d = max(d, 12 * SYM)
        d = min(d, hi)
        t = np.arange(3 * SYM)
        cap[2 * SYM:5 * SYM] += 0.3 * np.cos(2 * np.pi * 1500.0 * t / 48000.0)
        cap[d:d + L] += frame
    elif case == "last_good_state":
        cap[d:d + L] += frame
        cap += rng.normal(0, 0.18, n)
        state = (d - 7, 0.2)
    elif case == "two_frames":
        L2 = min(L, n - L - lo - 100)
        d = lo
        cap[d:d + L] += frame
        cap[d + L + 50:d + L + 50 + L2] += tx.transmit_byte(rng.integers(0, 256, tx.frame_bytes))[:L2]
        pl = None  # either frame may win the coarse search
    else:
        raise ValueError(case)
    return cap.astype(np.float32).astype(np.float64), pl, state
