"""mercury_b200 -- B200-native (sm_100a) receive hot path of the Mercury HF modem's physical layer.

Product = mercury_b200/libmercury_b200.so (C ABI in include/mercury_b200.h). This package is the host-side
mirror of the reference's interface for that path; see DESIGN.md.
"""
from .modes import MODES, ROBUST_MODES, THRESH_DB  # noqa: F401
from .telecom_system import (BASEBAND_C64, BASEBAND_CF16, BASEBAND_CI16, DECODER_MINSUM, DECODER_SPA, HANDOFF_FLOATS, NO, MFSK_PATTERN_DTYPE, RECEIVE_STATS_DTYPE, SAMPLES_F32, SAMPLES_F64, SAMPLES_I16, SAMPLES_I32,  # noqa: F401
                             STATS_DTYPE, YES, MercuryB200Error, TelecomSystemB200, build_tables_host, mfsk_fixed_delay, new_receive_stats, synth_frames)
