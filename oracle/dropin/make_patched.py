#!/usr/bin/env python3
"""oracle/dropin/make_patched.py -- TEST INFRASTRUCTURE ONLY: the reference's own cl_telecom_system with its receive path re-pointed at the
C ABI of libmercury_b200.so, exactly as INTEGRATION.md sections 2 and 2c tell a maintainer to do it.

  make_patched.py <reference root> <out dir>
      -> <out dir>/telecom_system_tail.cc    receive_byte() keeps its CPU front-end; lines 1132-1341 (demodulate .. CRC) and the SNR
                                             report of the success branch (:1362-1400) call mercury_b200_receive_baseband  
      -> <out dir>/telecom_system_whole.cc   the whole body of receive_byte() calls mercury_b200_receive_byte               

Nothing of the reference is stored in this repository: the script reads telecom_system.cc where it lies, checks that the lines it is about
to replace are the ones INTEGRATION.md names (anchors below; it refuses to patch a different revision), and writes the patched copies into
a build directory that is git-ignored (oracle/_ref/).  oracle/Makefile (target `dropin`) compiles them with the rest of the UNMODIFIED
physical layer into oracle/_ref/libmercury_ref_tail.so / libmercury_ref_whole.so, linked against libmercury_b200.so;
tests/test_gpu_dropin.py then runs the patched reference's own receive_byte() next to the unpatched one.
"""
import os
import sys

# what the replaced region must look like (1-based line -> substring), Rhizomatica/mercury @ c91aa4b
ANCHORS = {
    646: "st_receive_stats cl_telecom_system::receive_byte(double *data, int* out)",
    647: "{",
    1132: "{",
    1134: "int rx_nsymb = get_active_nsymb();",
    1337: "receive_stats.crc=0;",
    1341: "}",
    1343: "if(receive_stats.all_zeros==YES ||",
    1360: "else",
    1362: "if(M == MOD_MFSK)",
    1403: "receive_stats.message_decoded=YES;",
}

PRELUDE = r'''
// ---- added by oracle/dropin/make_patched.py (INTEGRATION.md section 2: "new members" kept out of the header by a side table) ----
#include "mercury_b200.h"
#include <map>
namespace {
struct B200Link { mercury_b200_t *h = nullptr; int config = -999, iters = -1, ctrl = -1, cfs = -1; };
mercury_b200_t *b200_of(cl_telecom_system *ts)
{
	static std::map<cl_telecom_system *, B200Link> links;
	B200Link &l = links[ts];
	if (!l.h) {
		if (mercury_b200_create(0, &l.h) != MERCURY_B200_OK) { fprintf(stderr, "mercury_b200_create: no B200\n"); exit(1); }
		const char *tables = getenv("MERCURY_B200_LDPC_TABLES");
		if (!tables || mercury_b200_load_tables(l.h, tables) != MERCURY_B200_OK) { fprintf(stderr, "mercury_b200_load_tables failed (MERCURY_B200_LDPC_TABLES)\n"); exit(1); }
	}
	if (l.config != ts->current_configuration || l.iters != ts->ldpc.nIteration_max) {  // cl_telecom_system::load_configuration(int), O(1) here
		if (mercury_b200_load_configuration(l.h, ts->current_configuration, ts->ldpc.nIteration_max) != MERCURY_B200_OK) {
			fprintf(stderr, "mercury_b200_load_configuration(%d): %s\n", ts->current_configuration, mercury_b200_last_error(l.h));
			exit(1);
		}
		l.config = ts->current_configuration, l.iters = ts->ldpc.nIteration_max, l.ctrl = -1;
	}
	if (l.cfs != (int)g_gui_state.coarse_freq_sync_enabled.load()) {  // the GUI's switch for the +-30 Hz search of trial 1 (:949)
		l.cfs = (int)g_gui_state.coarse_freq_sync_enabled.load();
		mercury_b200_set_coarse_freq_sync(l.h, l.cfs);
	}
	if (l.ctrl != (int)ts->mfsk_ctrl_mode) {
		mercury_b200_set_mfsk_ctrl_mode(l.h, ts->mfsk_ctrl_mode ? 1 : 0);
		l.ctrl = (int)ts->mfsk_ctrl_mode;
	}
	return l.h;
}
}  // namespace
// ---- end of the addition ----
'''

# replaces telecom_system.cc:1132-1341 inside the sync-trial loop
TAIL = r'''			// ---- telecom_system.cc:1132-1341 re-pointed at the B200 path (INTEGRATION.md section 2) ----
			mercury_b200_rx_stats b200_rs;
			{
				const std::complex<double> *b200_frame = &data_container.baseband_data[data_container.Nofdm * data_container.preamble_nSymb];
				if (mercury_b200_receive_baseband(b200_of(this), reinterpret_cast<const double *>(b200_frame), out, &b200_rs) != MERCURY_B200_OK) {
					fprintf(stderr, "mercury_b200_receive_baseband: %s\n", mercury_b200_last_error(b200_of(this)));
					exit(1);
				}
			}
			receive_stats.iterations_done = b200_rs.iterations_done;  // -1: skipped by the mean|H| < 0.3 gate (:1271)
			receive_stats.crc = b200_rs.crc;
			receive_stats.all_zeros = b200_rs.all_zeros;
			variance = b200_rs.variance;
			if (b200_rs.iterations_done < 0) {  // as :1268-1280
				skip_h_count++;
				receive_stats.sync_trials++;
				continue;
			}
			// ---- end of the re-pointed region; the decision below (:1343-1359) is the reference's own ----
'''

# the success branch's SNR report (:1362-1400) reads CPU-side arrays the re-pointed tail no longer fills: the library reports it
SNR = r'''				receive_stats.SNR = b200_rs.SNR;  // :1362-1400 (LS: pilot variance; ZF: re-encode; MFSK: 0), computed by the library
'''

# INTEGRATION.md section 2c (+ 2e for the ROBUST configurations): the whole body of receive_byte
WHOLE = r'''	// ---- the whole receive_byte() re-pointed at the B200 path (INTEGRATION.md sections 2c / 2e) ----
	{
		mercury_b200_receive_stats b200_rs = {};
		b200_rs.delay_of_last_decoded_message = receive_stats.delay_of_last_decoded_message;              // in: link state (:945-947)
		b200_rs.freq_offset_of_last_decoded_message = receive_stats.freq_offset_of_last_decoded_message;  //     (:1108-1110)
		if (M == MOD_MFSK) {
			b200_rs.mfsk_search_or_overflow = mfsk_fixed_delay >= 0 ? MERCURY_B200_MFSK_FIXED_DELAY(mfsk_fixed_delay)
										  : receive_stats.mfsk_search_raw - (int)data_container.nUnder_processing_events;
			mfsk_fixed_delay = -1;
		}
		if (mercury_b200_receive_byte(b200_of(this), data, out, &b200_rs) != MERCURY_B200_OK) {
			fprintf(stderr, "mercury_b200_receive_byte: %s\n", mercury_b200_last_error(b200_of(this)));
			exit(1);
		}
		receive_stats.iterations_done = b200_rs.iterations_done;
		receive_stats.delay = b200_rs.delay;
		receive_stats.sync_trials = b200_rs.sync_trials;
		receive_stats.message_decoded = b200_rs.message_decoded;
		receive_stats.crc = b200_rs.crc;
		receive_stats.all_zeros = b200_rs.all_zeros;
		receive_stats.SNR = b200_rs.SNR;
		receive_stats.freq_offset = b200_rs.freq_offset;
		receive_stats.signal_stregth_dbm = b200_rs.signal_stregth_dbm;
		receive_stats.coarse_metric = b200_rs.coarse_metric;
		receive_stats.delay_of_last_decoded_message = b200_rs.delay_of_last_decoded_message;              // out: link state (:1423-1427)
		receive_stats.freq_offset_of_last_decoded_message = b200_rs.freq_offset_of_last_decoded_message;
		receive_stats.frame_overflow_symbols = M == MOD_MFSK ? b200_rs.mfsk_search_or_overflow : 0;
		return receive_stats;
	}
	// ---- (the reference's own body follows, unreachable) ----
'''


def main():
    ref_root, out_dir = sys.argv[1], sys.argv[2]
    path = os.path.join(ref_root, "source", "physical_layer", "telecom_system.cc")
    lines = open(path).read().split("\n")
    for ln, want in ANCHORS.items():
        got = lines[ln - 1]
        if (want not in got) if len(want) > 1 else (got.strip() != want):
            raise SystemExit(f"{path}:{ln}: expected {want!r}, found {got.strip()!r} -- not the revision INTEGRATION.md describes, refusing to patch")
    inc = max(i for i, l in enumerate(lines[:60]) if l.startswith("#include"))  # after the last #include of the file's head
    os.makedirs(out_dir, exist_ok=True)

    # tail variant: [1132, 1341] -> TAIL, [1362, 1401] (the if / else-if chain that computes SNR) -> SNR
    end_snr = 1403 - 1  # 1-based line of message_decoded=YES, minus one
    while lines[end_snr - 1].strip() == "":
        end_snr -= 1  # last non-blank line before it = the closing brace of the ZF branch
    tail = lines[:inc + 1] + PRELUDE.split("\n") + lines[inc + 1:1131] + TAIL.split("\n") + lines[1341:1361] + SNR.split("\n") + lines[end_snr:]
    open(os.path.join(out_dir, "telecom_system_tail.cc"), "w").write("\n".join(tail))

    # whole variant: right after the opening brace of receive_byte (:647)
    whole = lines[:inc + 1] + PRELUDE.split("\n") + lines[inc + 1:647] + WHOLE.split("\n") + lines[647:]
    open(os.path.join(out_dir, "telecom_system_whole.cc"), "w").write("\n".join(whole))
    print(f"patched copies of {path} written to {out_dir}")


if __name__ == "__main__":
    main()
