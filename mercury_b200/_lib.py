"""ctypes loader of the C-ABI library (mercury_b200/libmercury_b200.so, built by mercury_b200/csrc/Makefile).

The library is the product: hand-written sm_100a kernels behind include/mercury_b200.h.  There is no Python
or CPU implementation of the RX path in this package; if the library is missing, importing fails loudly.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("MERCURY_B200_SO") or os.path.join(_HERE, "libmercury_b200.so")  # the override is for kernel tuning builds only
LDPC_TABLES = os.path.join(_HERE, "data", "ldpc_tables.bin")
HEADER = os.path.normpath(os.path.join(_HERE, "..", "include", "mercury_b200.h"))


class Geometry(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "config", "M", "bits_per_symbol", "ldpc_rate_num", "Nsymb", "Nc", "Nfft", "Ngi", "Nofdm", "nData", "nPilots",
        "nBits", "N", "K", "P", "nReal", "nVirtual", "preamble_nSymb", "frame_bytes", "estimator", "phase_only",
        "ldpc_iters", "ldpc_edges", "decoder")]


class RxStats(C.Structure):
    _fields_ = [("iterations_done", C.c_int32), ("crc", C.c_int32), ("all_zeros", C.c_int32), ("message_decoded", C.c_int32),
                ("SNR", C.c_float), ("variance", C.c_float), ("mean_H", C.c_float), ("reserved", C.c_int32)]


class ReceiveStats(C.Structure):
    """mercury_b200_receive_stats = st_receive_stats (telecom_system.h:63-82), fields of the OFDM branch."""
    _fields_ = [("iterations_done", C.c_int32), ("delay", C.c_int32), ("delay_of_last_decoded_message", C.c_int32), ("sync_trials", C.c_int32),
                ("message_decoded", C.c_int32), ("crc", C.c_int32), ("all_zeros", C.c_int32), ("mfsk_search_or_overflow", C.c_int32),
                ("freq_offset", C.c_double), ("freq_offset_of_last_decoded_message", C.c_double), ("SNR", C.c_double),
                ("signal_stregth_dbm", C.c_double), ("coarse_metric", C.c_double)]


def build(force=False):
    """Compile the library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcdir = os.path.join(_HERE, "csrc")
    newest = max(os.path.getmtime(os.path.join(srcdir, f)) for f in os.listdir(srcdir) if f.endswith((".cu", ".cpp", ".h", ".cuh")))
    newest = max(newest, os.path.getmtime(HEADER))
    if force or not os.path.exists(SO_PATH) or os.path.getmtime(SO_PATH) < newest:
        subprocess.check_call(["make", "-s", "-C", srcdir])
    return SO_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: build it with `make -C mercury_b200/csrc` "
                          "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    L = C.CDLL(SO_PATH)
    vp, sz, i32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
    sig = {
        "mercury_b200_version": (C.c_char_p, []),
        "mercury_b200_strerror": (C.c_char_p, [i32]),
        "mercury_b200_create": (i32, [i32, C.POINTER(vp)]),
        "mercury_b200_destroy": (None, [vp]),
        "mercury_b200_last_error": (C.c_char_p, [vp]),
        "mercury_b200_load_tables": (i32, [vp, C.c_char_p]),
        "mercury_b200_export_tables": (i32, [vp, vp, C.POINTER(sz)]),
        "mercury_b200_import_tables": (i32, [vp, vp, sz]),
        "mercury_b200_build_tables_host": (i32, [C.c_char_p, vp, C.POINTER(sz)]),
        "mercury_b200_load_configuration": (i32, [vp, i32, i32]),
        "mercury_b200_set_decoder": (i32, [vp, i32]),
        "mercury_b200_get_geometry": (i32, [vp, C.POINTER(Geometry)]),
        "mercury_b200_get_frame_size_bytes": (i32, [vp]),
        "mercury_b200_get_frame_size_bits": (i32, [vp]),
        "mercury_b200_broadcast_tables": (i32, [vp, vp, i32, vp]),
        "mercury_b200_demod_decode_batch": (i32, [vp, vp, sz, vp, vp, vp]),
        "mercury_b200_demod_decode_batch_device": (i32, [vp, vp, sz, vp, vp, vp, vp]),
        "mercury_b200_demod_decode_batch_fmt": (i32, [vp, vp, i32, C.c_float, sz, vp, vp, vp]),
        "mercury_b200_demod_decode_batch_device_fmt": (i32, [vp, vp, i32, C.c_float, sz, vp, vp, vp, vp]),
        "mercury_b200_demod_batch_device": (i32, [vp, vp, sz, vp, vp, vp, vp]),
        "mercury_b200_ldpc_decode_batch_device": (i32, [vp, vp, sz, vp, vp, vp]),
        "mercury_b200_set_debug_capture": (i32, [vp, vp, vp, vp]),
        "mercury_b200_receive_baseband": (i32, [vp, vp, vp, C.POINTER(RxStats)]),
        "mercury_b200_get_capture_samples": (i32, [vp]),
        "mercury_b200_set_mfsk_ctrl_mode": (i32, [vp, i32]),
        "mercury_b200_set_coarse_freq_sync": (i32, [vp, i32]),
        "mercury_b200_get_active_nsymb": (i32, [vp]),
        "mercury_b200_get_configuration": (i32, [C.c_double]),
        "mercury_b200_measure_signal_only_batch": (i32, [vp, vp, i32, sz, vp]),
        "mercury_b200_build_frontend_tables_host": (i32, [vp, vp, vp, vp, i32]),
        "mercury_b200_receive_byte": (i32, [vp, vp, vp, C.POINTER(ReceiveStats)]),
        "mercury_b200_receive_byte_batch": (i32, [vp, vp, i32, sz, vp, vp, vp]),
        "mercury_b200_receive_byte_batch_device": (i32, [vp, vp, i32, sz, vp, vp, vp]),
        "mercury_b200_get_total_frame_size": (i32, [vp]),
        "mercury_b200_build_tx_tables_host": (i32, [C.c_char_p, i32, vp, vp, vp, vp]),
        "mercury_b200_transmit_byte": (i32, [vp, vp, i32, vp, vp]),
        "mercury_b200_transmit_byte_batch": (i32, [vp, vp, vp, sz, vp, i32, vp]),
        "mercury_b200_transmit_byte_batch_ex": (i32, [vp, vp, vp, sz, vp, i32, i32, vp]),
        "mercury_b200_fir_tx_apply": (i32, [vp, vp, sz, vp]),
        "mercury_b200_transmit_byte_loc": (i32, [vp, vp, i32, vp, vp, i32]),
        "mercury_b200_reset_tx_stream": (i32, [vp]),
        "mercury_b200_transmit_byte_batch_device": (i32, [vp, vp, vp, sz, vp, i32, vp]),
        "mercury_b200_mfsk_patterns_batch": (i32, [vp, vp, i32, sz, i32, i32, vp]),
        "mercury_b200_generate_pattern_passband": (i32, [vp, i32, vp, vp]),
        "mercury_b200_detect_patterns_from_passband_batch": (i32, [vp, vp, i32, sz, i32, vp]),
        "mercury_b200_host_alloc": (vp, [sz]),
        "mercury_b200_host_free": (None, [vp]),
        "mercury_b200_device_alloc": (vp, [vp, sz]),
        "mercury_b200_device_free": (None, [vp, vp]),
        "mercury_b200_memcpy_h2d": (i32, [vp, vp, vp, sz]),
        "mercury_b200_memcpy_d2h": (i32, [vp, vp, vp, sz]),
        "mercury_b200_synchronize": (i32, [vp]),
        "mercury_b200_kernel_launches": (u64, [vp]),
        "mercury_b200_batcher_create": (i32, [vp, sz, C.c_uint, C.POINTER(vp)]),
        "mercury_b200_batcher_receive_baseband": (i32, [vp, vp, vp, C.POINTER(RxStats)]),
        "mercury_b200_batcher_create_passband": (i32, [vp, i32, sz, C.c_uint, C.POINTER(vp)]),
        "mercury_b200_batcher_receive_byte": (i32, [vp, vp, vp, vp]),
        "mercury_b200_batcher_get_counters": (i32, [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "mercury_b200_batcher_destroy": (None, [vp]),
        "mercury_b200_batcher_create_with_backend": (i32, [sz, sz, sz, C.c_uint, vp, vp, C.POINTER(vp)]),
        "mercury_b200_synth_frames": (i32, [C.c_char_p, i32, sz, u64, C.c_double, vp, vp, vp, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = the .so does not export what the header declares
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


EXPORTS = None  # filled lazily by tests: symbols declared in include/mercury_b200.h
