#!/usr/bin/env python3
"""tools/single_call_probe.py -- one link, the reference's own call shape: N mercury_b200_receive_byte() calls on one mode-8 capture
(the committed reference fixture).  Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel list of ONE call, or alone
for the host-side latency split (time.perf_counter around the call)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
import mercury_b200 as mb  # noqa: E402

n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = np.load(os.path.join(ROOT, "tests", "golden", "frontend_mode08_clean.npz"))
cap = g["capture"].astype(np.float64)
rng = np.random.default_rng(1)
cap = (cap + rng.normal(0, 0.02, cap.size)).astype(np.float32).astype(np.float64)
ts = mb.TelecomSystemB200(0)
ts.load_configuration(8, 50)
lat = []
for i in range(n_calls):
    t0 = time.perf_counter()
    out, st = ts.receive_byte(cap)
    lat.append((time.perf_counter() - t0) * 1e6)
print("decoded", int(st["message_decoded"][0]), "delay", int(st["delay"][0]), "launches/call", ts.kernel_launches // n_calls,
      "median_us %.1f min_us %.1f" % (float(np.median(lat)), float(np.min(lat))))
ts.close()
