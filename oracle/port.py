"""ctypes binding of oracle/libmercury_oracle.so (the plain-C restatement) -- TEST INFRASTRUCTURE ONLY.

Same surface as oracle/ref.py (class Port mirrors class Ref) so tests can run either against the other.
Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmercury_oracle.so")
LDPC_BLOB = os.path.normpath(os.path.join(_HERE, "..", "mercury_b200", "data", "ldpc_tables.bin"))

from .ref import GEOM_FIELDS, FrontEndMixin, _RxOut, _p  # noqa: E402  (shared record layouts)


def build(force=False):
    src = os.path.join(_HERE, "mercury_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.mo_mode_new.restype = C.c_void_p
        L.mo_mode_new.argtypes = [C.c_int, C.c_int, C.c_char_p]
        L.mo_mode_delete.argtypes = [C.c_void_p]
        L.mo_geometry.argtypes = [C.c_void_p, C.c_void_p]
        L.mo_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.mo_ldpc_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.mo_random_seq.argtypes = [C.c_uint, C.c_int, C.c_void_p]
        L.mo_crc16.argtypes = [C.c_void_p, C.c_int]
        L.mo_crc16.restype = C.c_int
        L.mo_tx_baseband.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.mo_rx_tail.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mo_rx_tail_timed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mo_rx_tail_timed.restype = C.c_double
        L.mo_ldpc_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mo_ldpc_decode.restype = C.c_int
        L.mo_receive_byte.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.mo_receive_byte_timed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.mo_receive_byte_timed.restype = C.c_double
        L.mo_frontend_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.mo_tx_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.mo_time_sync_mfsk.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.mo_time_sync_mfsk.restype = C.c_int
        L.mo_detect_ack_pattern.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.mo_detect_ack_pattern.restype = C.c_double
        L.mo_ack_pattern_baseband.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.mo_mfsk_tables.argtypes = [C.c_void_p, C.c_void_p]
        L.mo_set_mfsk_ctrl_mode.argtypes = [C.c_void_p, C.c_int]
        L.mo_set_mfsk_ctrl_mode.restype = C.c_int
        L.mo_set_coarse_freq_sync.argtypes = [C.c_void_p, C.c_int]
        L.mo_transmit_byte_loc.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.mo_transmit_byte_loc.restype = C.c_int
        L.mo_reset_tx_stream.argtypes = [C.c_void_p]
        L.mo_generate_pattern_passband.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.mo_generate_pattern_passband.restype = C.c_int
        L.mo_detect_pattern_from_passband.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.mo_detect_pattern_from_passband.restype = C.c_double
        L.mo_transmit_byte.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.mo_transmit_byte.restype = C.c_int
        L.mo_transmit_byte_nofilter.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.mo_transmit_byte_nofilter.restype = C.c_int
        L.mo_fir_tx_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def port_random(seed, n):
    out = np.zeros(n, np.int32)
    lib().mo_random_seq(seed, n, _p(out))
    return out


def port_crc16(data):
    a = np.asarray(list(data), np.int32)
    return lib().mo_crc16(_p(a), len(a))


class Port(FrontEndMixin):
    """The C restatement loaded with CONFIG_<config> and -I <ldpc_iters>."""

    _fe = "mo_"

    @staticmethod
    def _felib():
        return lib()

    def __init__(self, config, ldpc_iters=50):
        self.h = lib().mo_mode_new(config, ldpc_iters, LDPC_BLOB.encode())
        if not self.h:
            raise RuntimeError(f"mo_mode_new({config}) failed (ldpc blob {LDPC_BLOB})")
        self.config = config
        g = np.zeros(64, np.int32)
        lib().mo_geometry(self.h, _p(g))
        self.geom = {k: int(g[i]) for i, k in enumerate(GEOM_FIELDS)}
        self.__dict__.update(self.geom)
        self.nReal = self.nBits - self.P
        self.nVirtual = self.N - self.nBits

    def close(self):
        if self.h:
            lib().mo_mode_delete(self.h)
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
        lib().mo_tables(self.h, _p(ct), _p(ps), _p(sc), _p(co), _p(boost))
        return dict(carrier_type=ct.reshape(self.Nsymb, self.Nc), pilot_seq=ps, scrambler=sc,
                    constellation=co.view(np.complex128), pilot_boost=float(boost[0]))

    def ldpc_tables(self):
        dims = np.zeros(3, np.int32)
        lib().mo_ldpc_tables(self.h, _p(dims), None, None, None, None)
        Cw, Vw, dw = (int(x) for x in dims)
        Cm = np.zeros((self.P, Cw), np.int32)
        Vm = np.zeros((self.N, Vw), np.int32)
        d = np.zeros(dw, np.int32)
        E = np.zeros((self.P, Cw - 1), np.int32)
        lib().mo_ldpc_tables(self.h, _p(dims), _p(Cm), _p(Vm), _p(d), _p(E))
        return dict(C=Cm, V=Vm, d=d, Enc=E, N=self.N, K=self.K, P=self.P)

    def tx_baseband(self, payload, want_aux=False):
        pl = np.asarray(list(payload), np.int32)
        out = np.zeros(self.Nsymb * self.Nofdm, np.complex128)
        info = np.zeros(self.nReal, np.int32)
        cw = np.zeros(self.N, np.int32)
        framed = np.zeros(self.Nsymb * self.Nc, np.complex128)
        lib().mo_tx_baseband(self.h, _p(pl), len(pl), _p(out), _p(info), _p(cw), _p(framed))
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
        lib().mo_rx_tail(self.h, _p(bb), C.byref(o))
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
        secs = lib().mo_rx_tail_timed(self.h, _p(bb), n, _p(pay), _p(dec), _p(its))
        return secs, pay, dec, its

    def ldpc_decode(self, llr_cw):
        l = np.ascontiguousarray(llr_cw, np.float32)
        bits = np.zeros(self.K, np.int32)
        it = lib().mo_ldpc_decode(self.h, _p(l), _p(bits))
        return it, bits
