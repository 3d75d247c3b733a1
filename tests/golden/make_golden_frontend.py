#!/usr/bin/env python3
"""Generate tests/golden/frontend_*.npz from the UNMODIFIED reference (oracle/_ref) -- run in the build container only.

Each fixture is one pass-band capture buffer (float32-representable values; the reference runs on them widened to double)
built by tests/frontend_cases.py from the reference's own transmit_byte(SINGLE_MESSAGE), plus everything the reference's
receive_byte() (telecom_system.cc:646-1518) returned for it: the 12 reported fields, the payload bytes, the link state it
left behind and the post-synchronisation baseband_data its tail consumed.

Usage: python tests/golden/make_golden_frontend.py   (needs oracle/_ref/libmercury_ref.so, i.e. /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.normpath(os.path.join(HERE, "..", "..")))
from oracle import ref  # noqa: E402
from tests import frontend_cases as fc  # noqa: E402

FIXTURES = [(16, "freq_offset", 7), (8, "clean", 8), (16, "last_good_state", 9), (13, "tone_then_frame", 10)]


def main():
    for cfg, case, seed in FIXTURES:
        r = ref.Ref(cfg, 50)
        cap, pl, state = fc.make_capture(r, case, seed)
        a = r.receive_byte2(cap, *state)
        np.savez_compressed(
            os.path.join(HERE, f"frontend_mode{cfg:02d}_{case}.npz"), config=cfg, ldpc_iters=50, case=case,
            capture=cap.astype(np.float32), tx_payload=np.asarray(pl if pl is not None else [], np.int32),
            state_in=np.asarray(state, np.float64), state_out=np.asarray([a["last_delay"], a["last_freq"]], np.float64),
            stats=np.asarray([a[k] for k in ref.STAT12], np.float64), rx_payload=a["payload"], baseband=a["baseband"])
        print(cfg, case, {k: a[k] for k in ("decoded", "delay", "sync_trials", "iterations", "freq_offset")})
        r.close()


if __name__ == "__main__" and "--tx" not in sys.argv:
    main()


def tx_fixture():
    """tx_mode16.npz: one reference transmit_byte(SINGLE_MESSAGE) frame with its TX tables."""
    r = ref.Ref(16, 50)
    t = r.tx_tables()
    pl = np.random.default_rng(16).integers(0, 256, r.frame_bytes).astype(np.int32)
    out, after = r.transmit_byte2(pl, 4321)
    np.savez_compressed(os.path.join(HERE, "tx_mode16.npz"), config=16, payload=pl, start_sample=4321, start_sample_after=after, passband=out,
                        pre_eq=t["pre_eq"], preamble=t["preamble"], tx1=t["tx1"], tx2=t["tx2"])
    print("tx fixture", out.size, after)


if __name__ == "__main__" and "--tx" in sys.argv:
    tx_fixture()
