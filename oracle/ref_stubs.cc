/*
 * oracle/ref_stubs.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Link-time stubs that let the UNMODIFIED reference physical layer
 * (/root/reference/source/physical_layer/*.cc + source/common/*.cc) link without
 * its audio, GUI and CLI translation units.  Nothing here is reference code: these
 * are definitions of the external symbols the physical layer expects its host
 * program to provide (SURVEY.md section 8c lists them):
 *
 *   - source/main.cc globals            (carrier_frequency_offset, radio_type, ...)
 *   - source/audioio/audioio.c globals  (capture_buffer, capture_prep_mutex, tx_transfer, ...)
 *   - source/gui/* entry points         (get_gui_state, waterfall_push_samples)
 */
#include <pthread.h>
#include <cstddef>
#include "common/ring_buffer_posix.h"
#include "gui/gui_state.h"

extern "C" {
double carrier_frequency_offset = 0.0;
double test_tx_carrier_offset = 0.0;
int radio_type = 0;
char *input_dev = nullptr;
char *output_dev = nullptr;
bool shutdown_ = false;
}

int g_verbose = 0;

cbuf_handle_t capture_buffer = nullptr;
cbuf_handle_t playback_buffer = nullptr;
pthread_mutex_t capture_prep_mutex = PTHREAD_MUTEX_INITIALIZER;

int tx_transfer(double *, size_t) { return 0; }
int rx_transfer(double *, size_t) { return 0; }

st_gui_state &get_gui_state()
{
	static st_gui_state state;
	return state;
}

extern "C" void waterfall_push_samples(const double *, int) {}
