"""The 17 OFDM configurations (reference: telecom_system.cc:2506-2624; geometry SURVEY.md 8a).

Static description only (sizes for buffer allocation, Es/N0 operating points); the authoritative tables are
built by the C++ side (csrc/mb_tables.cpp) and cross-checked against this list in tests/test_tables.py.
"""

# FER < 0.1 thresholds in dB, include/common/common_defines.h:130-147
THRESH_DB = [-10, -7.5, -6, -4.5, -3.5, -2.5, -1.5, -0.5, 0.5, 1.5, 3, 4, 6.5, 7.5, 9, 12.5, 13.5]
LDPC_EDGES = {1: 3574, 2: 3859, 3: 4439, 4: 4651, 5: 5409, 6: 5616, 8: 6049, 14: 6604}

_DEF = [(2, 1, 4, 1), (2, 2, 4, 1), (2, 3, 4, 1), (2, 4, 4, 1), (2, 5, 4, 1), (2, 6, 4, 1), (2, 8, 4, 1), (4, 5, 4, 1),
        (4, 6, 4, 1), (4, 8, 4, 1), (8, 6, 3, 1), (8, 8, 3, 1), (4, 14, 3, 1), (16, 8, 2, 1), (8, 14, 2, 1), (16, 14, 2, 0),
        (32, 14, 1, 0)]
_NSYMB = {2: 48, 4: 24, 8: 16, 16: 12, 32: 9}


def _mode(cfg):
    M, rate, pre, est = _DEF[cfg]
    S = _NSYMB[M]
    cells = S * 50
    n_pil = sum(1 for s in range(S) for c in range(50) if s % 3 == c % 3)
    n_data = cells - n_pil
    bps = M.bit_length() - 1
    n_bits = n_data * bps
    K = 100 * rate
    P = 1600 - K
    n_real = n_bits - P
    return dict(config=cfg, M=M, bps=bps, rate_num=rate, Nsymb=S, nData=n_data, nPilots=n_pil, nBits=n_bits, K=K, P=P,
                nReal=n_real, nVirtual=1600 - n_bits, frame_bytes=(n_real - 16) // 8, preamble_nSymb=pre, estimator=est,
                phase_only=int(M in (2, 4, 8)), edges=LDPC_EDGES[rate], thresh_db=THRESH_DB[cfg],
                # algorithmic bytes of the FFT/equalise stage (samples incl. GI in, LLRs out)
                demod_bytes=S * 272 * 8 + n_bits * 4)


MODES = [_mode(c) for c in range(17)]

# ROBUST_0..2 (MFSK, include/common/common_defines.h:63-65): tones per stream, streams, LDPC rate
ROBUST_MODES = {}
for _cfg, (_M, _streams, _rate) in {100: (32, 1, 1), 101: (16, 2, 1), 102: (16, 2, 4)}.items():
    _bps = (_M.bit_length() - 1) * _streams
    _K = 100 * _rate
    ROBUST_MODES[_cfg] = dict(config=_cfg, M=200, mfsk_M=_M, nStreams=_streams, bps=_bps, rate_num=_rate, Nsymb=1600 // _bps, nData=1600 // _bps,
                              nPilots=0, nBits=1600, K=_K, P=1600 - _K, nReal=_K, nVirtual=0, frame_bytes=(_K - 16) // 8, preamble_nSymb=4,
                              estimator=1, phase_only=0, edges=LDPC_EDGES[_rate])
