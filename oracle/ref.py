"""ctypes binding of oracle/_ref/libmercury_ref.so -- TEST INFRASTRUCTURE ONLY.

The .so is the UNMODIFIED reference physical layer (compiled by oracle/Makefile from
/root/reference) plus oracle/ref_driver.cc.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libmercury_ref.so")

GEOM_FIELDS = [
    "Nsymb", "Nc", "Nfft", "Ngi", "Nofdm", "nData", "nPilots", "nBits", "N", "K", "P", "M",
    "preamble_nSymb", "frame_bytes", "estimator", "amp_restore", "bit_il_block", "tf_il_block",
    "interp_rate", "buffer_Nsymb", "total_frame_size", "Cwidth", "Vwidth", "dwidth", "ldpc_iters",
    "outer_reserved", "ls_win_h", "ls_win_w",
]

RATE_OF_CONFIG = {0: 1, 1: 2, 2: 3, 3: 4, 4: 5, 5: 6, 6: 8, 7: 5, 8: 6, 9: 8, 10: 6, 11: 8, 12: 14, 13: 8, 14: 14, 15: 14, 16: 14}


def available():
    return os.path.exists(_SO)


class _RxOut(C.Structure):
    _fields_ = [
        ("Y", C.c_void_p), ("H", C.c_void_p), ("Z", C.c_void_p), ("llr_demod", C.c_void_p),
        ("llr_cw", C.c_void_p), ("bits", C.c_void_p), ("bytes", C.c_void_p), ("payload", C.c_void_p),
        ("stats", C.c_void_p),
    ]


_libs = {}
# the same reference with receive_byte() re-pointed at libmercury_b200.so (oracle/dropin/make_patched.py, `make -C oracle dropin`)
SO_DROPIN_TAIL = os.path.join(_HERE, "_ref", "libmercury_ref_tail.so")
SO_DROPIN_WHOLE = os.path.join(_HERE, "_ref", "libmercury_ref_whole.so")


def lib(so=None):
    so = so or _SO
    if so not in _libs:
        if not os.path.exists(so):
            raise RuntimeError(f"{so} not built (make -C oracle ref / dropin; needs /root/reference)")
        L = C.CDLL(so)
        L.mref_create.restype = C.c_void_p
        L.mref_create.argtypes = [C.c_int, C.c_int]
        L.mref_destroy.argtypes = [C.c_void_p]
        if hasattr(L, "mref_load_configuration"):
            L.mref_load_configuration.argtypes = [C.c_void_p, C.c_int]
        L.mref_geometry.argtypes = [C.c_void_p, C.c_void_p]
        L.mref_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.mref_ldpc_tables.argtypes = [C.c_int] + [C.c_void_p] * 5
        L.mref_ldpc_tables.restype = C.c_int
        L.mref_random.argtypes = [C.c_uint, C.c_int, C.c_void_p]
        L.mref_crc16.argtypes = [C.c_void_p, C.c_int]
        L.mref_crc16.restype = C.c_int
        L.mref_tx_baseband.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.mref_rx_tail.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mref_rx_tail_timed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mref_rx_tail_timed.restype = C.c_double
        L.mref_ldpc_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mref_ldpc_decode.restype = C.c_int
        L.mref_transmit_byte.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.mref_transmit_byte.restype = C.c_int
        L.mref_receive_byte.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.mref_receive_byte2.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.mref_receive_byte_timed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.mref_receive_byte_timed.restype = C.c_double
        L.mref_frontend_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.mref_tx_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.mref_time_sync_mfsk.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.mref_time_sync_mfsk.restype = C.c_int
        L.mref_detect_ack_pattern.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.mref_detect_ack_pattern.restype = C.c_double
        L.mref_ack_pattern_baseband.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.mref_mfsk_tables.argtypes = [C.c_void_p, C.c_void_p]
        L.mref_set_mfsk_ctrl_mode.argtypes = [C.c_void_p, C.c_int]
        L.mref_set_mfsk_ctrl_mode.restype = C.c_int
        L.mref_set_coarse_freq_sync.argtypes = [C.c_int]
        L.mref_transmit_byte_loc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.mref_transmit_byte_loc.restype = C.c_int
        L.mref_reset_tx_stream.argtypes = [C.c_void_p]
        L.mref_generate_pattern_passband.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.mref_generate_pattern_passband.restype = C.c_int
        L.mref_detect_pattern_from_passband.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.mref_detect_pattern_from_passband.restype = C.c_double
        L.mref_transmit_byte2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.mref_transmit_byte2.restype = C.c_int
        L.mref_transmit_byte_nofilter.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.mref_transmit_byte_nofilter.restype = C.c_int
        L.mref_fir_tx_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _libs[so] = L
    return _libs[so]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def ref_random(seed, n):
    out = np.zeros(n, np.int32)
    lib().mref_random(seed, n, _p(out))
    return out


def ref_crc16(data):
    a = np.asarray(list(data), np.int32)
    return lib().mref_crc16(_p(a), len(a))


def ldpc_tables(rate_num):
    """-> dict(C[P,Cw], V[N,Vw], d[dw], Enc[P,Cw-1]) as held by the reference."""
    dims = np.zeros(3, np.int32)
    assert lib().mref_ldpc_tables(rate_num, _p(dims), None, None, None, None) == 0
    Cw, Vw, dw = (int(x) for x in dims)
    N, K = 1600, 100 * rate_num
    P = N - K
    Cm = np.zeros((P, Cw), np.int32)
    Vm = np.zeros((N, Vw), np.int32)
    d = np.zeros(dw, np.int32)
    E = np.zeros((P, Cw - 1), np.int32)
    lib().mref_ldpc_tables(rate_num, _p(dims), _p(Cm), _p(Vm), _p(d), _p(E))
    return dict(C=Cm, V=Vm, d=d, Enc=E, N=N, K=K, P=P)


STAT12 = ("iterations", "crc", "all_zeros", "decoded", "snr", "delay", "sync_trials", "freq_offset", "coarse_metric",
          "signal_dbm", "buffer_samples", "frame_bytes")


class FrontEndMixin:
    """The whole receive_byte() (pass-band capture in), shared by Ref (mref_*) and Port (mo_*); `_fe` = symbol prefix."""

    def capture_samples(self):
        return self.Nofdm * self.buffer_Nsymb * self.interp_rate

    def receive_byte2(self, passband, last_delay=-1, last_freq=0.0, search_start_symb=0, mfsk_fixed_delay=-1):
        L = self._felib()
        n = self.capture_samples()
        pb = np.ascontiguousarray(passband, np.float64)
        assert pb.size == n, (pb.size, n)
        out = np.zeros(self.frame_bytes, np.int32)
        st = np.zeros(12, np.float64)
        state = np.array([last_delay, last_freq, search_start_symb, 0, mfsk_fixed_delay], np.float64)
        bb = np.zeros((self.Nsymb + self.preamble_nSymb) * self.Nofdm, np.complex128)
        getattr(L, self._fe + "receive_byte2" if self._fe == "mref_" else self._fe + "receive_byte")(
            self.h, _p(pb), _p(out), _p(st), _p(state), _p(bb))
        r = {k: (float(st[i]) if k in ("snr", "freq_offset", "coarse_metric", "signal_dbm") else int(st[i])) for i, k in enumerate(STAT12)}
        r.update(payload=out, baseband=bb, last_delay=int(state[0]), last_freq=float(state[1]), frame_overflow_symbols=int(state[3]), mfsk_fixed_delay_after=int(state[4]))
        return r

    def receive_byte_timed(self, captures):
        pb = np.ascontiguousarray(captures, np.float64)
        n = pb.size // self.capture_samples()
        dec = np.zeros(n, np.int32)
        secs = getattr(self._felib(), self._fe + "receive_byte_timed")(self.h, _p(pb), n, _p(dec))
        return secs, dec

    def tx_tables(self):
        """Preamble carriers, pre-equalisation channel, transmit FIRs, TX constants (row 2)."""
        n = self.preamble_nSymb * self.Nc
        pre = np.zeros(n, np.complex128)
        typ = np.zeros(n, np.int32)
        peq = np.zeros(self.Nc, np.complex128)
        nt = np.zeros(2, np.int32)
        c1 = np.zeros(128, np.float64)
        c2 = np.zeros(128, np.float64)
        k = np.zeros(8, np.float64)
        getattr(self._felib(), self._fe + "tx_tables")(self.h, _p(pre), _p(typ), _p(peq), _p(nt), _p(c1), _p(c2), _p(k))
        return dict(preamble=pre, preamble_is_carrier=(typ == int(k[6])).astype(np.int32), pre_eq=peq, tx1=c1[:nt[0]].copy(), tx2=c2[:nt[1]].copy(),
                    output_power=k[0], preamble_boost=k[1], preamble_papr=k[2], data_papr=k[3], start_sample_after_init=int(k[4]),
                    total_frame_size=int(k[5]))

    def fir_tx_apply(self, x):
        """ofdm.FIR_tx1.apply then ofdm.FIR_tx2.apply over a buffer of any length (the ARQ layer's batch filtering, arq_common.cc:2243-2246)."""
        a = np.ascontiguousarray(x, np.float64)
        out = np.zeros(a.size, np.float64)
        getattr(self._felib(), self._fe + "fir_tx_apply")(self.h, _p(a), a.size, _p(out))
        return out

    def transmit_byte2(self, payload, start_sample, no_filter=False):
        """transmit_byte(SINGLE_MESSAGE | NO_FILTER_MESSAGE) from a chosen running carrier sample counter -> (passband[total_frame_size], counter after)."""
        pl = np.asarray(list(payload), np.int32)
        out = np.zeros(self.total_frame_size + 16, np.float64)
        st = np.array([float(start_sample)], np.float64)
        if no_filter:
            name = self._fe + "transmit_byte_nofilter"
        else:
            name = "mref_transmit_byte2" if self._fe == "mref_" else "mo_transmit_byte"
        n = getattr(self._felib(), name)(self.h, _p(pl), len(pl), _p(out), _p(st))
        return out[:n], int(st[0])

    # ---- MFSK pattern functions (row 3): bbi = complex128 buffer at the pass-band rate --------------------------------
    def time_sync_mfsk(self, bbi, search_start_symb=0):
        b = np.ascontiguousarray(bbi, np.complex128)
        return int(getattr(self._felib(), self._fe + "time_sync_mfsk")(self.h, _p(b), b.size, int(search_start_symb)))

    def detect_ack_pattern(self, bbi, use_break_tones=False):
        b = np.ascontiguousarray(bbi, np.complex128)
        m = np.zeros(1, np.int32)
        v = getattr(self._felib(), self._fe + "detect_ack_pattern")(self.h, _p(b), b.size, int(use_break_tones), _p(m))
        return float(v), int(m[0])

    def ack_pattern_baseband(self, use_break_tones=False):
        out = np.zeros(16 * self.Nofdm, np.complex128)
        getattr(self._felib(), self._fe + "ack_pattern_baseband")(self.h, int(use_break_tones), _p(out))
        return out

    def mfsk_tables(self):
        t = np.zeros(32, np.int32)
        getattr(self._felib(), self._fe + "mfsk_tables")(self.h, _p(t))
        return dict(M=int(t[0]), nBits=int(t[1]), nStreams=int(t[2]), tone_hop_step=int(t[3]), stream_offsets=t[4:8].copy(),
                    preamble_tones=t[8:12].copy(), ack_tones=t[12:20].copy(), break_tones=t[20:28].copy())

    # ---- the ARQ-facing tone-pattern calls (any configuration; dedicated 16-MFSK plan) -------------------------------
    def generate_pattern_passband(self, use_break_tones=False, start_sample=0):
        out = np.zeros(16 * self.Nofdm * 4, np.float64)
        st = np.array([float(start_sample)], np.float64)
        n = getattr(self._felib(), self._fe + "generate_pattern_passband")(self.h, int(use_break_tones), _p(out), _p(st))
        return out[:n], int(st[0])

    def detect_pattern_from_passband(self, data, use_break_tones=False):
        d = np.ascontiguousarray(data, np.float64)
        m = np.zeros(1, np.int32)
        v = getattr(self._felib(), self._fe + "detect_pattern_from_passband")(self.h, _p(d), d.size, int(use_break_tones), _p(m))
        return float(v), int(m[0])

    def reset_tx_stream(self):
        getattr(self._felib(), self._fe + "reset_tx_stream")(self.h)

    def transmit_byte_loc(self, payload, start_sample, message_location):
        """transmit_byte with message_location FIRST 0 / MIDDLE 1 / FLUSH 2 / SINGLE 3 / NO_FILTER 4 -> (passband, counter after)."""
        pl = np.asarray(list(payload), np.int32)
        out = np.zeros(self.total_frame_size + 16, np.float64)
        st = np.array([float(start_sample)], np.float64)
        n = getattr(self._felib(), self._fe + "transmit_byte_loc")(self.h, _p(pl), len(pl), _p(out), _p(st), int(message_location))
        return out[:n], int(st[0])

    def set_coarse_freq_sync(self, enable):
        """g_gui_state.coarse_freq_sync_enabled: the optional +-30 Hz search of trial 1 (GLOBAL in the reference, per mode object in the port)."""
        if self._fe == "mref_":
            self._felib().mref_set_coarse_freq_sync(int(bool(enable)))
        else:
            self._felib().mo_set_coarse_freq_sync(self.h, int(bool(enable)))

    def set_mfsk_ctrl_mode(self, enable):
        """set_mfsk_ctrl_mode(bool) -> get_active_nsymb() (shortened control frames in ROBUST_0 / ROBUST_1)."""
        return int(getattr(self._felib(), self._fe + "set_mfsk_ctrl_mode")(self.h, int(bool(enable))))

    def frontend_tables(self):
        nt = np.zeros(2, np.int32)
        a = np.zeros(64, np.float64)
        b = np.zeros(64, np.float64)
        c = np.zeros(8, np.float64)
        getattr(self._felib(), self._fe + "frontend_tables")(self.h, _p(nt), _p(a), _p(b), _p(c))
        return dict(ts=a[:nt[0]].copy(), data=b[:nt[1]].copy(), fs=c[0], fc=c[1], amp=c[2], bandwidth=c[3], trials_max=int(c[4]),
                    use_last_time=int(c[5]), use_last_freq=int(c[6]), ignore_limit=c[7])


class Ref(FrontEndMixin):
    """One reference cl_telecom_system loaded with CONFIG_<config> and -I <ldpc_iters>."""

    _fe = "mref_"

    def _felib(self):
        return self._L

    def __init__(self, config, ldpc_iters=50, so=None):
        self._L = lib(so)
        self.h = self._L.mref_create(config, ldpc_iters)
        self.config = config
        g = np.zeros(64, np.int32)
        self._L.mref_geometry(self.h, _p(g))
        self.geom = {k: int(g[i]) for i, k in enumerate(GEOM_FIELDS)}
        self.__dict__.update(self.geom)
        self.nReal = self.nBits - self.P
        self.nVirtual = self.N - self.nBits

    def load_configuration(self, config):
        """cl_telecom_system::load_configuration(int) on this object (telecom_system.cc:2487); refreshes the geometry attributes."""
        self._L.mref_load_configuration(self.h, config)
        self.config = config
        g = np.zeros(64, np.int32)
        self._L.mref_geometry(self.h, _p(g))
        self.geom = {k: int(g[i]) for i, k in enumerate(GEOM_FIELDS)}
        self.__dict__.update(self.geom)
        self.nReal = self.nBits - self.P
        self.nVirtual = self.N - self.nBits

    def close(self):
        if self.h:
            self._L.mref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tables(self):
        ct = np.zeros(self.Nsymb * self.Nc, np.int32)
        ps = np.zeros(self.nPilots, np.float64)
        sc = np.zeros(self.N, np.int32)
        co = np.zeros(2 * self.M, np.float64)
        boost = np.zeros(1, np.float64)
        self._L.mref_tables(self.h, _p(ct), _p(ps), _p(sc), _p(co), _p(boost))
        return dict(carrier_type=ct.reshape(self.Nsymb, self.Nc), pilot_seq=ps, scrambler=sc,
                    constellation=co.view(np.complex128), pilot_boost=float(boost[0]))

    def tx_baseband(self, payload, want_aux=False):
        pl = np.asarray(list(payload), np.int32)
        out = np.zeros(self.Nsymb * self.Nofdm, np.complex128)
        info = np.zeros(self.nReal, np.int32)
        cw = np.zeros(self.N, np.int32)
        framed = np.zeros(self.Nsymb * self.Nc, np.complex128)
        self._L.mref_tx_baseband(self.h, _p(pl), len(pl), _p(out), _p(info), _p(cw), _p(framed))
        if want_aux:
            return out, dict(info_bits=info, codeword=cw, framed=framed.reshape(self.Nsymb, self.Nc))
        return out

    def rx_tail(self, baseband):
        bb = np.ascontiguousarray(baseband, np.complex128).reshape(-1)
        assert bb.size == self.Nsymb * self.Nofdm
        cells = self.Nsymb * self.Nc
        r = dict(
            Y=np.zeros(cells, np.complex128), H=np.zeros(cells, np.complex128), Z=np.zeros(cells, np.complex128),
            llr_demod=np.zeros(self.nBits, np.float32), llr_cw=np.zeros(self.N, np.float32),
            bits=np.zeros(self.K, np.int32), bytes=np.zeros(self.nReal // 8, np.int32),
            payload=np.zeros(self.frame_bytes, np.int32), stats=np.zeros(8, np.float64),
        )
        o = _RxOut(*[_p(r[k]) for k in ("Y", "H", "Z", "llr_demod", "llr_cw", "bits", "bytes", "payload", "stats")])
        self._L.mref_rx_tail(self.h, _p(bb), C.byref(o))
        st = r.pop("stats")
        r.update(iterations=int(st[0]), crc=int(st[1]), all_zeros=int(st[2]), decoded=int(st[3]), snr=float(st[4]),
                 variance=np.float32(st[5]), mean_H=float(st[7]))
        for k in ("Y", "H", "Z"):
            r[k] = r[k].reshape(self.Nsymb, self.Nc)
        return r

    def rx_tail_timed(self, baseband_batch):
        bb = np.ascontiguousarray(baseband_batch, np.complex128)
        n = bb.size // (self.Nsymb * self.Nofdm)
        pay = np.zeros((n, self.frame_bytes), np.int32)
        dec = np.zeros(n, np.int32)
        its = np.zeros(n, np.int32)
        secs = self._L.mref_rx_tail_timed(self.h, _p(bb), n, _p(pay), _p(dec), _p(its))
        return secs, pay, dec, its

    def ldpc_decode(self, llr_cw):
        l = np.ascontiguousarray(llr_cw, np.float32)
        bits = np.zeros(self.K, np.int32)
        it = self._L.mref_ldpc_decode(self.h, _p(l), _p(bits))
        return it, bits

    def transmit_byte(self, payload):
        pl = np.asarray(list(payload), np.int32)
        out = np.zeros(self.total_frame_size + 16, np.float64)
        n = self._L.mref_transmit_byte(self.h, _p(pl), len(pl), _p(out))
        return out[:n]

    def receive_byte(self, passband):
        n = self.Nofdm * self.buffer_Nsymb * self.interp_rate
        pb = np.ascontiguousarray(passband, np.float64)
        assert pb.size == n, (pb.size, n)
        out = np.zeros(self.frame_bytes, np.int32)
        st = np.zeros(8, np.float64)
        bb = np.zeros((self.Nsymb + self.preamble_nSymb) * self.Nofdm, np.complex128)
        self._L.mref_receive_byte(self.h, _p(pb), _p(out), _p(st), _p(bb))
        return dict(payload=out, iterations=int(st[0]), crc=int(st[1]), all_zeros=int(st[2]), decoded=int(st[3]),
                    snr=float(st[4]), delay=int(st[5]), sync_trials=int(st[6]), freq_offset=float(st[7]), baseband=bb)
