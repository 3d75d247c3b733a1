"""Host-side mirror of the reference's physical-layer RX interface over the C ABI.

Reference: class cl_telecom_system (include/physical_layer/telecom_system.h:85-199) -- the three calls the
datalink layer makes into the physical layer are load_configuration(int), receive_byte(double*, int*) and
get_frame_size_bytes()/bits() (INTERNALS:11, source/datalink_layer/arq_common.cc:627,2668).  TelecomSystemB200
keeps those names, argument meaning and error behaviour (no exceptions from the receive call: a frame that
does not decode comes back with message_decoded == NO and SNR == -99.9), for the part of receive_byte() that
north_star moves to the GPU: everything after time/frequency synchronisation (telecom_system.cc:1132-1429).

All arithmetic happens in mercury_b200/libmercury_b200.so (CUDA, sm_100a).  numpy / torch are used only to
hold buffers.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Geometry, ReceiveStats, RxStats

YES, NO = 1, 0
DECODER_SPA, DECODER_MINSUM = 0, 1
BASEBAND_C64, BASEBAND_CI16, BASEBAND_CF16 = 0, 1, 2  # MERCURY_B200_BASEBAND_*: base-band sample formats of the batch calls
HANDOFF_FLOATS = 2400  # MERCURY_B200_HANDOFF_FLOATS: float32 per frame of the stage hand-off buffer between the two kernels

STATS_DTYPE = np.dtype([("iterations_done", "<i4"), ("crc", "<i4"), ("all_zeros", "<i4"), ("message_decoded", "<i4"),
                        ("SNR", "<f4"), ("variance", "<f4"), ("mean_H", "<f4"), ("reserved", "<i4")])
assert STATS_DTYPE.itemsize == C.sizeof(RxStats) == 32
# mercury_b200_receive_stats (the whole receive_byte(), pass-band in): st_receive_stats of the OFDM branch
RECEIVE_STATS_DTYPE = np.dtype([("iterations_done", "<i4"), ("delay", "<i4"), ("delay_of_last_decoded_message", "<i4"), ("sync_trials", "<i4"),
                                ("message_decoded", "<i4"), ("crc", "<i4"), ("all_zeros", "<i4"), ("mfsk_search_or_overflow", "<i4"),
                                ("freq_offset", "<f8"), ("freq_offset_of_last_decoded_message", "<f8"), ("SNR", "<f8"),
                                ("signal_stregth_dbm", "<f8"), ("coarse_metric", "<f8")])
assert RECEIVE_STATS_DTYPE.itemsize == C.sizeof(ReceiveStats) == 72
MFSK_PATTERN_DTYPE = np.dtype([("time_sync_delay", "<i4"), ("ack_matched", "<i4"), ("break_matched", "<i4"), ("reserved", "<i4"),
                               ("ack_metric", "<f8"), ("break_metric", "<f8")])
SAMPLES_F64, SAMPLES_F32, SAMPLES_I16, SAMPLES_I32 = 0, 1, 2, 3
MFSK_FIXED_DELAY_FLAG = 0x40000000  # MERCURY_B200_MFSK_FIXED_DELAY_FLAG


def mfsk_fixed_delay(d):
    """MERCURY_B200_MFSK_FIXED_DELAY(d): the IN value of a record's mfsk_search_or_overflow that stands for the reference's one-shot member
    mfsk_fixed_delay = d (telecom_system.h:110, .cc:663-673; set by the ARQ layer's overflow recapture, arq_common.cc:2830-2833)."""
    return MFSK_FIXED_DELAY_FLAG | max(int(d), 0)
_SAMPLE_FORMATS = {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.int16): 2, np.dtype(np.int32): 3}


def new_receive_stats(n):
    """n link-state records of a fresh link (telecom_system.cc:42,49: no last decoded message yet)."""
    st = np.zeros(n, RECEIVE_STATS_DTYPE)
    st["delay_of_last_decoded_message"] = -1
    return st


class MercuryB200Error(RuntimeError):
    pass


def _vp(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):  # torch tensor (host or device)
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class TelecomSystemB200:
    """One long-lived RX object per GPU, like the reference's cl_telecom_system."""

    def __init__(self, device=0, ldpc_table_path=None, tables_blob=None):
        self._L = _lib.lib()
        h = C.c_void_p()
        rc = self._L.mercury_b200_create(int(device), C.byref(h))
        if rc != 0:
            raise MercuryB200Error(f"mercury_b200_create(device={device}): {self._L.mercury_b200_strerror(rc).decode()}")
        self._h = h
        self.device = int(device)
        if tables_blob is not None:
            buf = np.frombuffer(bytes(tables_blob), np.uint8) if not isinstance(tables_blob, np.ndarray) else tables_blob
            self._check(self._L.mercury_b200_import_tables(self._h, _vp(buf), buf.size))
        else:
            path = ldpc_table_path or _lib.LDPC_TABLES
            self._check(self._L.mercury_b200_load_tables(self._h, path.encode()))
        self.current_configuration = -1

    def close(self):
        if getattr(self, "_h", None):
            self._L.mercury_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self._L.mercury_b200_last_error(self._h).decode()
            raise MercuryB200Error(f"{self._L.mercury_b200_strerror(rc).decode()}: {msg}")

    # ---- reference surface -------------------------------------------------------------------------------
    def load_configuration(self, configuration, ldpc_nIteration_max=50):
        """cl_telecom_system::load_configuration(int) (+ the CLI's -I). O(1): all 17 modes stay resident."""
        self._check(self._L.mercury_b200_load_configuration(self._h, int(configuration), int(ldpc_nIteration_max)))
        self.current_configuration = int(configuration)
        g = Geometry()
        self._check(self._L.mercury_b200_get_geometry(self._h, C.byref(g)))
        self.geometry = {n: getattr(g, n) for n, _ in Geometry._fields_}
        return self.geometry

    def set_decoder(self, decoder):
        self._check(self._L.mercury_b200_set_decoder(self._h, int(decoder)))
        if self.current_configuration >= 0:
            self.geometry["decoder"] = int(decoder)

    def get_frame_size_bytes(self):
        return self._L.mercury_b200_get_frame_size_bytes(self._h)

    def get_frame_size_bits(self):
        return self._L.mercury_b200_get_frame_size_bits(self._h)

    def receive_baseband(self, baseband):
        """The tail of receive_byte() on one synchronised frame: baseband = Nsymb*272 complex128 samples.
        Returns (out, stats): out = one int per payload byte (like `int* out`), stats = dict of st_receive_stats fields."""
        g = self.geometry
        bb = np.ascontiguousarray(baseband, np.complex128).reshape(-1)
        if bb.size != g["Nsymb"] * g["Nofdm"]:
            raise ValueError("baseband must hold Nsymb*Nofdm complex samples")
        out = np.zeros(g["frame_bytes"], np.int32)
        st = RxStats()
        self._check(self._L.mercury_b200_receive_baseband(self._h, _vp(bb), _vp(out), C.byref(st)))
        return out, {n: getattr(st, n) for n, _ in RxStats._fields_ if n != "reserved"}

    def get_configuration(self, SNR):
        """char cl_telecom_system::get_configuration(double SNR) (telecom_system.cc:3036-3106): the gear-shift ladder."""
        return int(self._L.mercury_b200_get_configuration(float(SNR)))

    def measure_signal_only(self, captures):
        """double measure_signal_only(double* data) (telecom_system.cc:1520-1541) on [n, capture_samples] captures -> dBm per capture."""
        x = np.ascontiguousarray(captures)
        if x.dtype not in _SAMPLE_FORMATS:
            raise TypeError("captures must be float64, float32, int16 or int32")
        x = x.reshape(-1, self.get_capture_samples())
        out = np.zeros(x.shape[0], np.float64)
        self._check(self._L.mercury_b200_measure_signal_only_batch(self._h, _vp(x), _SAMPLE_FORMATS[x.dtype], x.shape[0], _vp(out)))
        return out

    def set_coarse_freq_sync(self, enable):
        """g_gui_state.coarse_freq_sync_enabled (gui_state.h:143): the optional +-30 Hz search before trial 1 of receive_byte()
        (telecom_system.cc:949-1013).  Off by default."""
        self._check(self._L.mercury_b200_set_coarse_freq_sync(self._h, int(bool(enable))))

    def set_mfsk_ctrl_mode(self, enable):
        """void set_mfsk_ctrl_mode(bool) (telecom_system.cc:1572-1575) -> get_active_nsymb()."""
        return int(self._L.mercury_b200_set_mfsk_ctrl_mode(self._h, int(bool(enable))))

    def get_active_nsymb(self):
        return int(self._L.mercury_b200_get_active_nsymb(self._h))

    def get_capture_samples(self):
        """Pass-band samples receive_byte() is handed: Nofdm * buffer_Nsymb * 4 (data_container.cc:133-153)."""
        return self._L.mercury_b200_get_capture_samples(self._h)

    def receive_byte(self, data, stats=None):
        """st_receive_stats cl_telecom_system::receive_byte(double* data, int* out) (telecom_system.cc:646-1518), whole: `data` is
        one pass-band capture (float64).  `stats` carries the link state between calls (new_receive_stats(1) for a new link).
        Returns (out, stats): out = one int per payload byte."""
        pb = np.ascontiguousarray(data, np.float64).reshape(-1)
        if pb.size != self.get_capture_samples():
            raise ValueError("data must hold get_capture_samples() doubles")
        st = new_receive_stats(1) if stats is None else stats
        out = np.zeros(self.geometry["frame_bytes"], np.int32)
        self._check(self._L.mercury_b200_receive_byte(self._h, _vp(pb), _vp(out), st.ctypes.data_as(C.POINTER(ReceiveStats))))
        return out, st

    def receive_byte_batch(self, captures, stats=None, want_baseband=False):
        """n captures of independent links, host buffers [n, capture_samples]: float64 / float32, or the PCM formats of the reference's
        audio layer int16 (x / 32768.0) / int32 (x / INT_MAX) converted on the device like audioio.c:893-940.
        Returns (payload[n, frame_bytes] u8, stats[n], baseband[n, (pre+Nsymb)*272] complex128 | None)."""
        x = np.ascontiguousarray(captures)
        if x.dtype not in _SAMPLE_FORMATS:
            raise TypeError("captures must be float64, float32, int16 or int32")
        cs = self.get_capture_samples()
        if x.size % cs:
            raise ValueError("captures size is not a whole number of capture buffers")
        n = x.size // cs
        st = new_receive_stats(n) if stats is None else stats
        payload = np.zeros((n, self.geometry["frame_bytes"]), np.uint8)
        g = self.geometry
        bb = np.zeros((n, (g["preamble_nSymb"] + g["Nsymb"]) * g["Nofdm"]), np.complex128) if want_baseband else None
        self._check(self._L.mercury_b200_receive_byte_batch(self._h, _vp(x), _SAMPLE_FORMATS[x.dtype], n,
                                                            _vp(payload), _vp(st), _vp(bb)))
        return payload, st, bb

    def receive_byte_batch_device(self, d_captures, sample_format, n, d_payload, d_stats, stream=0):
        self._check(self._L.mercury_b200_receive_byte_batch_device(self._h, _vp(d_captures), int(sample_format), int(n), _vp(d_payload),
                                                                   _vp(d_stats), C.c_void_p(stream)))

    # ---- MFSK tone-pattern detectors --------------------------------------------------
    def mfsk_patterns_batch(self, bbi, search_start_symb=0):
        """bbi [n_buffers, n_samples] complex128 / complex64 base-band at the pass-band rate -> structured array per buffer:
        time_sync_delay (cl_ofdm::time_sync_mfsk), ack_metric / ack_matched, break_metric / break_matched (cl_ofdm::detect_ack_pattern)."""
        b = np.ascontiguousarray(bbi)
        if b.dtype not in (np.complex128, np.complex64):
            raise TypeError("bbi must be complex128 or complex64")
        b = b.reshape(-1, b.shape[-1])
        out = np.zeros(b.shape[0], MFSK_PATTERN_DTYPE)
        self._check(self._L.mercury_b200_mfsk_patterns_batch(self._h, _vp(b), SAMPLES_F32 if b.dtype == np.complex64 else SAMPLES_F64, b.shape[0],
                                                             b.shape[1], int(search_start_symb), _vp(out)))
        return out

    def generate_pattern_passband(self, use_break_tones=False, passband_start_sample=0):
        """generate_ack_pattern_passband / generate_break_pattern_passband (telecom_system.cc:1589-1689) -> (samples float64[17408], counter after)."""
        out = np.zeros(16 * 1088, np.float64)
        st = np.array([int(passband_start_sample)], np.uint64)
        n = self._L.mercury_b200_generate_pattern_passband(self._h, int(use_break_tones), _vp(out), _vp(st))
        if n < 0:
            self._check(n)
        return out[:n], int(st[0])

    def detect_patterns_from_passband_batch(self, passband):
        """detect_ack_pattern_from_passband + detect_break_pattern_from_passband on [n_buffers, n_samples] pass-band buffers (any of the capture
        sample formats) -> structured array (ack_metric, ack_matched, break_metric, break_matched)."""
        x = np.ascontiguousarray(passband)
        if x.dtype not in _SAMPLE_FORMATS:
            raise TypeError("passband must be float64, float32, int16 or int32")
        x = x.reshape(-1, x.shape[-1])
        out = np.zeros(x.shape[0], MFSK_PATTERN_DTYPE)
        self._check(self._L.mercury_b200_detect_patterns_from_passband_batch(self._h, _vp(x), _SAMPLE_FORMATS[x.dtype], x.shape[0], x.shape[1], _vp(out)))
        return out

    # ---- TX chain --------------------------------------------------------------------
    def get_total_frame_size(self):
        """Pass-band samples of one transmitted frame: (preamble_nSymb + Nsymb) * Nofdm * 4 (data_container.total_frame_size)."""
        return self._L.mercury_b200_get_total_frame_size(self._h)

    def transmit_byte(self, data, passband_start_sample=None):
        """void cl_telecom_system::transmit_byte(int* data, int nBytes, double* out, SINGLE_MESSAGE) (telecom_system.cc:342-553).
        Returns (out float64[total_frame_size], counter after); passband_start_sample=None = a freshly initialised reference object."""
        d = np.asarray(list(data), np.int32)
        out = np.zeros(self.get_total_frame_size(), np.float64)
        fresh = 0 if self.geometry["M"] == 200 else 1088  # where a freshly initialised reference object's counter stands (MFSK: no pre-equalisation pass)
        st = np.array([fresh if passband_start_sample is None else int(passband_start_sample)], np.uint64)
        self._check(self._L.mercury_b200_transmit_byte(self._h, _vp(d), int(d.size), _vp(out), _vp(st)))
        return out, int(st[0])

    def transmit_byte_loc(self, data, passband_start_sample, message_location):
        """transmit_byte(data, nBytes, out, message_location) with FIRST 0 / MIDDLE 1 / FLUSH 2 (streaming, one frame of latency) / SINGLE 3 /
        NO_FILTER 4 -> (out float64[total_frame_size], counter after)."""
        d = np.asarray(list(data), np.int32)
        out = np.zeros(self.get_total_frame_size(), np.float64)
        st = np.array([int(passband_start_sample)], np.uint64)
        self._check(self._L.mercury_b200_transmit_byte_loc(self._h, _vp(d), int(d.size), _vp(out), _vp(st), int(message_location)))
        return out, int(st[0])

    def reset_tx_stream(self):
        self._check(self._L.mercury_b200_reset_tx_stream(self._h))

    def fir_tx_apply(self, x):
        """ofdm.FIR_tx1.apply + ofdm.FIR_tx2.apply over a host buffer of any length (the ARQ layer's batch filtering, arq_common.cc:2243-2246)."""
        a = np.ascontiguousarray(x, np.float64)
        out = np.zeros(a.size, np.float64)
        self._check(self._L.mercury_b200_fir_tx_apply(self._h, _vp(a), a.size, _vp(out)))
        return out

    def transmit_byte_batch(self, payload, start_sample=None, dtype=np.float64, want_codeword=False, message_location=3, out=None):
        """payload [n, frame_bytes] uint8 -> pass-band frames [n, total_frame_size] (float64 or float32) [, codewords [n, 1600] u8].
        message_location: 3 = SINGLE_MESSAGE (filtered), 4 = NO_FILTER_MESSAGE (clipped, before the transmit FIRs).
        out: optional preallocated [n, total_frame_size] array of `dtype` (pinned memory lets the D2H copy overlap the kernels)."""
        pl = np.ascontiguousarray(payload, np.uint8).reshape(-1, self.geometry["frame_bytes"])
        n = pl.shape[0]
        if out is None:
            out = np.zeros((n, self.get_total_frame_size()), dtype)
        elif out.dtype != np.dtype(dtype) or out.shape != (n, self.get_total_frame_size()) or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous [n, total_frame_size] array of dtype")
        st = None if start_sample is None else np.ascontiguousarray(start_sample, np.uint64)
        cw = np.zeros((n, self.geometry["N"]), np.uint8) if want_codeword else None
        self._check(self._L.mercury_b200_transmit_byte_batch_ex(self._h, _vp(pl), _vp(st), n, _vp(out), _SAMPLE_FORMATS[np.dtype(dtype)],
                                                                int(message_location), _vp(cw)))
        return (out, cw) if want_codeword else out

    def transmit_byte_batch_device(self, d_payload, d_start_sample, n, d_passband, out_format, stream=0):
        self._check(self._L.mercury_b200_transmit_byte_batch_device(self._h, _vp(d_payload), _vp(d_start_sample), int(n), _vp(d_passband),
                                                                    int(out_format), C.c_void_p(stream)))

    # ---- batched entry points ----------------------------------------------------------------------------
    def demod_decode_batch(self, baseband, want_llr=False, out=None, scale=1.0):
        """Host buffers: baseband [B, Nsymb, 272] complex64 (or float32 [..., 2]), or the 4-byte formats int16 [..., 2] (sample = integer x
        scale) / float16 [..., 2].  Returns (payload[B,frame_bytes] u8, stats[B], llr_cw|None)."""
        g = self.geometry
        bb = np.ascontiguousarray(baseband)
        if bb.dtype == np.complex64:
            bb = bb.view(np.float32)
        fmt = {np.dtype(np.float32): BASEBAND_C64, np.dtype(np.int16): BASEBAND_CI16, np.dtype(np.float16): BASEBAND_CF16}.get(bb.dtype)
        if fmt is None:
            raise TypeError("baseband must be complex64 / float32, int16 or float16 (re, im interleaved)")
        per = g["Nsymb"] * g["Nofdm"] * 2
        if bb.size % per:
            raise ValueError("baseband size is not a whole number of frames")
        n = bb.size // per
        if out is None:
            payload = np.zeros((n, g["frame_bytes"]), np.uint8)
            stats = np.zeros(n, STATS_DTYPE)
        else:
            payload, stats = out
        llr = np.zeros((n, g["N"]), np.float32) if want_llr else None
        self._check(self._L.mercury_b200_demod_decode_batch_fmt(self._h, _vp(bb), fmt, float(scale), n, _vp(payload), _vp(stats), _vp(llr)))
        return payload, stats, llr

    def demod_decode_batch_device(self, d_baseband, n_frames, d_payload, d_stats, d_llr_cw=None, stream=0, sample_format=0, scale=1.0):
        self._check(self._L.mercury_b200_demod_decode_batch_device_fmt(self._h, _vp(d_baseband), int(sample_format), float(scale), int(n_frames),
                                                                       _vp(d_payload), _vp(d_stats), _vp(d_llr_cw), C.c_void_p(stream)))

    def demod_batch_device(self, d_baseband, n_frames, d_llr, d_stats, d_llr_cw=None, stream=0):
        self._check(self._L.mercury_b200_demod_batch_device(self._h, _vp(d_baseband), int(n_frames), _vp(d_llr), _vp(d_stats),
                                                            _vp(d_llr_cw), C.c_void_p(stream)))

    def ldpc_decode_batch_device(self, d_llr, n_frames, d_payload, d_stats, stream=0):
        self._check(self._L.mercury_b200_ldpc_decode_batch_device(self._h, _vp(d_llr), int(n_frames), _vp(d_payload), _vp(d_stats),
                                                                  C.c_void_p(stream)))

    def set_debug_capture(self, d_Y=None, d_H=None, d_Z=None):
        self._check(self._L.mercury_b200_set_debug_capture(self._h, _vp(d_Y), _vp(d_H), _vp(d_Z)))

    def export_tables(self):
        n = C.c_size_t(0)
        self._check(self._L.mercury_b200_export_tables(self._h, None, C.byref(n)))
        buf = np.zeros(n.value, np.uint8)
        self._check(self._L.mercury_b200_export_tables(self._h, _vp(buf), C.byref(n)))
        return buf

    def synchronize(self):
        self._check(self._L.mercury_b200_synchronize(self._h))

    @property
    def kernel_launches(self):
        return int(self._L.mercury_b200_kernel_launches(self._h))


def build_tables_host(ldpc_table_path=None):
    """The table blob, built on the host (no device needed)."""
    L = _lib.lib()
    path = (ldpc_table_path or _lib.LDPC_TABLES).encode()
    n = C.c_size_t(0)
    rc = L.mercury_b200_build_tables_host(path, None, C.byref(n))
    if rc != 0:
        raise MercuryB200Error(L.mercury_b200_strerror(rc).decode())
    buf = np.zeros(n.value, np.uint8)
    rc = L.mercury_b200_build_tables_host(path, _vp(buf), C.byref(n))
    if rc != 0:
        raise MercuryB200Error(L.mercury_b200_strerror(rc).decode())
    return buf


def synth_frames(config, n_frames, seed=0, esn0_db=300.0, payload=None, n_threads=0, out=None, ldpc_table_path=None):
    """Synthetic baseband frames (host TX chain + AWGN). Returns (baseband[B,Nsymb,272] complex64, payload[B,frame_bytes] u8)."""
    import os
    L = _lib.lib()
    from .modes import MODES
    m = MODES[config]
    fb = m["frame_bytes"]
    if out is None:
        out = np.zeros((n_frames, m["Nsymb"], 272), np.complex64)
    pl_out = np.zeros((n_frames, fb), np.uint8)
    pin = None
    if payload is not None:
        pin = np.ascontiguousarray(payload, np.uint8).reshape(n_frames, fb)
    if n_threads <= 0:
        n_threads = max(1, (os.cpu_count() or 1))
    rc = L.mercury_b200_synth_frames((ldpc_table_path or _lib.LDPC_TABLES).encode(), int(config), int(n_frames), int(seed),
                                     float(esn0_db), _vp(pin), _vp(out), _vp(pl_out), int(n_threads))
    if rc != 0:
        raise MercuryB200Error(L.mercury_b200_strerror(rc).decode())
    return out, pl_out
