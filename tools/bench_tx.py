#!/usr/bin/env python3
"""tools/bench_tx.py -- frames/s through the GPU TX chain: payload bytes in, pass-band frames out
(transmit_byte(SINGLE_MESSAGE), telecom_system.cc:342-553), next to the unmodified reference's transmit_byte on one host core.

  python tools/bench_tx.py [--config 8 --frames 4096 --steps 5 --warmup 2 --out f32|f64 --cpu-frames 64]
Prints ONE JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=8)
    ap.add_argument("--frames", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--out", default="f32", choices=["f32", "f64"])
    ap.add_argument("--cpu-frames", type=int, default=64)
    a = ap.parse_args()
    import torch

    import mercury_b200 as mb
    dev = torch.device("cuda", 0)
    ts = mb.TelecomSystemB200(0)
    g = ts.load_configuration(a.config, 50)
    n, fb, L = a.frames, g["frame_bytes"], ts.get_total_frame_size()
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    d_pl = torch.randint(0, 256, (n, fb), device=dev, dtype=torch.uint8, generator=gen)
    d_start = (torch.arange(n, device=dev, dtype=torch.int64) * L + 1088)
    dt = torch.float32 if a.out == "f32" else torch.float64
    d_out = torch.empty((n, L), device=dev, dtype=dt)
    fmt = mb.SAMPLES_F32 if a.out == "f32" else mb.SAMPLES_F64
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        ts.transmit_byte_batch_device(d_pl, d_start, n, d_out, fmt, stream=stream)

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    l0 = ts.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    launches = (ts.kernel_launches - l0) // a.steps
    # e2e: payload bytes from host memory, pass-band frames back to pinned host memory (two device slots: the copy of one chunk runs under
    # the kernels of the next); the same call with a fresh pageable numpy array is timed next to it
    h_pl = d_pl.cpu().numpy()
    h_start = d_start.cpu().numpy().astype(np.uint64)
    np_dt = np.float32 if a.out == "f32" else np.float64
    h_out_t = torch.empty((n, L), dtype=torch.float32 if a.out == "f32" else torch.float64, pin_memory=True)
    host_out = h_out_t.numpy()
    ts.transmit_byte_batch(h_pl, h_start, dtype=np_dt, out=host_out)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ts.transmit_byte_batch(h_pl, h_start, dtype=np_dt, out=host_out)
    dt_e2e = (time.perf_counter() - t0) / a.steps
    t0 = time.perf_counter()
    pageable = ts.transmit_byte_batch(h_pl, h_start, dtype=np_dt)
    dt_pageable = time.perf_counter() - t0
    assert np.array_equal(pageable, host_out)
    same = bool(np.array_equal(host_out, d_out.cpu().numpy()))
    cpu = None
    if a.cpu_frames > 0:
        from oracle import port, ref
        o, kind = (ref.Ref(a.config, 50), "reference") if ref.available() else (port.Port(a.config, 50), "port")
        k = min(a.cpu_frames, n)
        t0 = time.perf_counter()
        worst = 0.0
        for i in range(k):
            want, _ = o.transmit_byte2(h_pl[i], int(h_start[i]))
            worst = max(worst, float(np.abs(host_out[i] - want).max() / np.abs(want).max()))
        secs = time.perf_counter() - t0
        cpu = {"value": k / secs, "unit": "frames/s", "cores": 1, "kind": kind,
               "sample": f"first {k} frames of the same batch through the {'unmodified ' if kind == 'reference' else 'restated '}transmit_byte(SINGLE_MESSAGE), "
                         f"{secs:.1f} s; worst relative sample difference to the GPU's {a.out} output {worst:.2e}"}
    out_bytes = L * (4 if a.out == "f32" else 8)
    print(json.dumps({
        "metric": "frames_per_s_transmit_byte_passband", "value": n / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms, "higher_is_better": True, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"mode {a.config}, {n} frames of {fb} payload bytes -> {L} pass-band samples ({a.out}) each",
                   "l2_policy": f"outputs {n * out_bytes / 1e9:.2f} GB + {n * L * 20 / 1e9:.2f} GB of fp64 intermediates >> 126 MB L2"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": n * out_bytes / (ms * 1e-3) / 1e9, "unit": "GB/s",
                     "note": "algorithmic bytes = the pass-band output only (payload in is ~0.2 %); the chain keeps three fp64 intermediates in HBM"},
        "e2e": {"value": n / dt_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(n * (fb + 8)), "d2h_bytes_per_step": int(n * out_bytes),
                "api": "mercury_b200_transmit_byte_batch (pinned host output, double-buffered D2H chunks)", "identical_to_device_run": same,
                "pcie_gbs": n * out_bytes / dt_e2e / 1e9, "pageable_frames_per_s": n / dt_pageable},
        "cpu_baseline": cpu}), flush=True)


if __name__ == "__main__":
    main()
