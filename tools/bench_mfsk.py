#!/usr/bin/env python3
"""tools/bench_mfsk.py -- the MFSK row: frames/s through the ROBUST tail (FFT + tone detection + LDPC + CRC) and
buffers/s through the tone-pattern detectors, next to the reference CPU implementation on one host core.

  python tools/bench_mfsk.py [--config 100 --frames 8192 --sigma 40 --steps 5 --warmup 2 --buffers 512 --cpu-frames 64]
Prints ONE JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=100)
    ap.add_argument("--frames", type=int, default=8192)
    ap.add_argument("--sigma", type=float, default=40.0, help="complex noise sigma per base-band sample (tone amplitude 7.07 / sqrt(streams))")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--buffers", type=int, default=512)
    ap.add_argument("--cpu-frames", type=int, default=64)
    a = ap.parse_args()
    import torch

    import mercury_b200 as mb
    from oracle import port, ref
    import mfsk_cases as mc
    dev = torch.device("cuda", 0)
    o, kind = (ref.Ref(a.config, 50), "reference") if ref.available() else (port.Port(a.config, 50), "port")
    ts = mb.TelecomSystemB200(0)
    g = ts.load_configuration(a.config, 50)
    R = mb.ROBUST_MODES[a.config]
    n, S, fb = a.frames, g["Nsymb"], g["frame_bytes"]
    rng = np.random.default_rng(1)
    U = 64  # distinct clean frames from the oracle's TX chain, tiled, independent noise per frame on the device
    pls = rng.integers(0, 256, (U, fb)).astype(np.uint8)
    clean = np.stack([o.tx_baseband(pls[i]) for i in range(U)]).astype(np.complex64).reshape(U, S, 272)
    d_clean = torch.from_numpy(clean).to(dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(3)
    d_x = torch.empty((n, S, 272), dtype=torch.complex64, device=dev)
    for i in range(0, n, U):
        k = min(U, n - i)
        d_x[i:i + k] = d_clean[:k] + torch.view_as_complex(torch.randn((k, S, 272, 2), device=dev, generator=gen) * a.sigma)
    pl_all = np.tile(pls, ((n + U - 1) // U, 1))[:n]
    d_pay = torch.zeros((n, fb), dtype=torch.uint8, device=dev)
    d_st = torch.zeros((n, 32), dtype=torch.uint8, device=dev)
    d_llr = torch.empty((n, mb.HANDOFF_FLOATS), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(a.warmup):
        ts.demod_decode_batch_device(d_x, n, d_pay, d_st, None, stream=stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        ts.demod_decode_batch_device(d_x, n, d_pay, d_st, None, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    ts.demod_batch_device(d_x, n, d_llr, d_st, None, stream=stream)
    ev[1].record()
    ts.ldpc_decode_batch_device(d_llr, n, d_pay, d_st, stream=stream)
    ev[2].record()
    torch.cuda.synchronize()
    t_demod, t_ldpc = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    st = d_st.cpu().numpy().view(mb.STATS_DTYPE).reshape(-1)
    pay = d_pay.cpu().numpy()
    dec = st["message_decoded"] == 1
    mism = int((pay[dec] != pl_all[dec]).any(axis=1).sum())
    its = np.clip(st["iterations_done"], 0, 50).astype(np.float64)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    demod_bytes = S * 272 * 8 + 1600 * 4
    cpu = None
    if a.cpu_frames > 0:
        k = min(a.cpu_frames, n)
        secs, cpay, cdec, cits = o.rx_tail_timed(d_x[:k].cpu().numpy().astype(np.complex128))
        cpu = {"value": k / secs, "unit": "frames/s", "cores": 1, "kind": kind,
               "sample": f"first {k} frames of the same batch, {secs:.1f} s, {int(cdec.sum())} decoded, verdicts equal to the GPU's: "
                         f"{int((cdec.astype(bool) == dec[:k]).sum())}/{k}"}
    # tone-pattern detectors
    nb = a.buffers
    kinds = ["ack", "break", "frame", "noise"]
    base = np.stack([mc.pattern_buffer(o, kinds[i % 4], 1000 + i)[0] for i in range(8)]).astype(np.complex64)
    bufs = np.tile(base, ((nb + 7) // 8, 1))[:nb]
    ts.mfsk_patterns_batch(bufs[:8])
    t0 = time.perf_counter()
    res = ts.mfsk_patterns_batch(bufs)
    dt_pat = time.perf_counter() - t0
    t0 = time.perf_counter()
    for i in range(4):
        o.time_sync_mfsk(base[i].astype(np.complex128)), o.detect_ack_pattern(base[i].astype(np.complex128), False), o.detect_ack_pattern(base[i].astype(np.complex128), True)
    dt_cpu_pat = (time.perf_counter() - t0) / 4
    print(json.dumps({
        "metric": "mfsk_frames_per_s_demod_ldpc_decode", "value": n / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ROBUST_{a.config - 100} ({R['mfsk_M']}-MFSK x{R['nStreams']}, LDPC {R['rate_num']}/16, Nsymb {S}), -I 50, {n} frames, "
                               f"complex noise sigma {a.sigma}", "l2_policy": f"inputs {n * S * 272 * 8 / 1e9:.2f} GB >> 126 MB L2"},
        "roofline": {"kernel": "k_mfsk_demod", "bound": "hbm", "achieved": demod_bytes * n / (t_demod * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": demod_bytes * n / (t_demod * 1e-3) / 1e9 / peak, "traffic": None, "algorithmic_bytes_per_frame": demod_bytes, "kernel_ms": t_demod},
        "ldpc": {"kernel_ms": t_ldpc, "mean_iterations": float(its.mean()), "edge_updates_per_s": float(its.sum()) * R["edges"] / (t_ldpc * 1e-3)},
        "cpu_baseline": cpu,
        "integrity": {"frames_decoded": int(dec.sum()), "frames": n, "payload_mismatches_among_decoded": mism},
        "patterns": {"buffers_per_s": nb / dt_pat, "unit": "buffers/s (64 symbols of base-band each, host complex64 in, H2D included)",
                     "cpu_reference_buffers_per_s": 1.0 / dt_cpu_pat, "api": "mercury_b200_mfsk_patterns_batch",
                     "ack_detected": int((res["ack_metric"] > 8).sum()), "break_detected": int((res["break_metric"] > 8).sum())}}), flush=True)


if __name__ == "__main__":
    main()
