#!/usr/bin/env python3
"""tools/bench_frontend.py -- captures/s through the WHOLE receive_byte(): pass-band capture buffers in,
payloads out, front-end + tail on the GPU; next to the unmodified reference's receive_byte() on one host core.

A capture = Nofdm*buffer_Nsymb*4 real samples at 48 kHz (what cl_telecom_system::receive_byte is handed, telecom_system.cc:646)
holding one frame of the reference's transmit_byte(SINGLE_MESSAGE) at a random delay in white noise.

  python tools/bench_frontend.py [--config 8 --captures 1024 --steps 5 --warmup 2 --noise 0.02 --fmt f32|f64 --cpu-captures 16]
Prints ONE JSON line.  Needs a B200; the CPU baseline leg uses oracle/_ref (the unmodified reference) when it is built, else the C port.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)


def tx_frames(cfg, distinct, rng):
    """-> (frames [k, L] float64, payloads [k, fb]).  Mode 8: the pass-band frame inside the committed reference fixture
    (tests/golden/frontend_mode08_clean.npz, produced by the reference's transmit_byte) -- no oracle involved; other modes need
    oracle/_ref for the reference's TX."""
    fx = os.path.join(ROOT, "tests", "golden", f"frontend_mode{cfg:02d}_clean.npz")
    if os.path.exists(fx):
        g = np.load(fx)
        cap = g["capture"].astype(np.float64)
        nz = np.flatnonzero(cap)
        import mercury_b200 as mb
        m = mb.MODES[cfg]
        L = (m["Nsymb"] + m["preamble_nSymb"]) * 1088
        d = int(g["stats"][5])  # the reference's own sync delay: frame start within a few samples
        start = max(0, min(int(nz[0]), d))
        return cap[start:start + L + 64][None, :], g["tx_payload"].astype(np.uint8)[None, :]
    from oracle import ref
    r = ref.Ref(cfg, 50)
    pls = rng.integers(0, 256, (distinct, r.frame_bytes))
    frames = np.stack([r.transmit_byte(pls[k]) for k in range(distinct)])
    r.close()
    return frames, pls.astype(np.uint8)


def make_captures(cfg, n, noise, dev, buf, pre, seed=11, distinct=16):
    """-> (d_caps [n, buf] float64 on dev holding float32-representable values, payloads [n, fb], delays [n])."""
    import torch
    rng = np.random.default_rng(seed)
    frames, pls = tx_frames(cfg, distinct, rng)
    L = frames.shape[1]
    lo, hi = (pre + 1) * 1088 + 10, buf - L - 2000
    delays = rng.integers(lo, hi, n)
    which = rng.integers(0, frames.shape[0], n)
    d_frames = torch.from_numpy(frames).to(dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    caps = torch.randn((n, buf), device=dev, dtype=torch.float64, generator=gen) * noise
    idx = torch.arange(L, device=dev)
    rows = torch.arange(n, device=dev)[:, None]
    cols = torch.from_numpy(delays).to(dev)[:, None] + idx[None, :]
    caps[rows, cols] += d_frames[torch.from_numpy(which).to(dev)]
    caps = caps.float().double()  # float32-representable values: the f32 and f64 entry points see the same numbers
    return caps, pls[which], delays


def run(config=8, captures=1024, steps=5, warmup=2, noise=0.02, fmt="f32", cpu_captures=16, e2e=True, ts=None):
    """One measurement; returns the dict that main() prints (bench.py embeds it as its "receive_byte" leg)."""
    import torch

    import mercury_b200 as mb
    dev = torch.device("cuda", torch.cuda.current_device())
    own = ts is None
    if own:
        ts = mb.TelecomSystemB200(dev.index)
    g = ts.load_configuration(config, 50)
    n, buf, fb = captures, ts.get_capture_samples(), g["frame_bytes"]
    caps64, pls, delays = make_captures(config, n, noise, dev, buf, g["preamble_nSymb"])
    if fmt == "i16":  # 16-bit PCM capture (audioio.c:907: x / 32768.0); the CPU baseline below sees the same converted values
        import torch as _t
        d_x = _t.clamp(_t.round(caps64 * 32768.0), -32768, 32767).to(_t.int16).contiguous()
        caps64 = d_x.double() / 32768.0
    else:
        d_x = caps64.float().contiguous() if fmt == "f32" else caps64
    sfmt = {"f64": mb.SAMPLES_F64, "f32": mb.SAMPLES_F32, "i16": mb.SAMPLES_I16}[fmt]
    st0 = torch.from_numpy(mb.new_receive_stats(n).view(np.uint8).reshape(n, -1)).to(dev)
    d_st = st0.clone()
    d_pay = torch.zeros((n, fb), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        d_st.copy_(st0)
        ts.receive_byte_batch_device(d_x, sfmt, n, d_pay, d_st, stream=stream)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    l0 = ts.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = (ts.kernel_launches - l0) // steps
    st = d_st.cpu().numpy().view(mb.RECEIVE_STATS_DTYPE).reshape(-1)
    pay = d_pay.cpu().numpy()
    dec = st["message_decoded"] == 1
    mism = int((pay[dec] != pls[dec]).any(axis=1).sum())

    e2e_d = None
    if e2e:
        h_x = torch.empty(d_x.shape, dtype=d_x.dtype, pin_memory=True)
        h_x.copy_(d_x)
        hx = h_x.numpy()
        ts.receive_byte_batch(hx)  # (dtype int16 / float32 / float64 selects the sample format)
        t0 = time.perf_counter()
        for _ in range(steps):
            p2, s2, _ = ts.receive_byte_batch(hx)
        dt = (time.perf_counter() - t0) / steps
        e2e_d = {"value": n / dt, "unit": "captures/s", "h2d_bytes_per_step": int(hx.nbytes + n * 72), "d2h_bytes_per_step": int(n * (fb + 72)),
                 "api": "mercury_b200_receive_byte_batch (pinned host buffers, double-buffered H2D chunks)",
                 "identical_to_device_run": bool(np.array_equal(p2, pay)), "pcie_gbs": hx.nbytes / dt / 1e9}
        del h_x

    single = None
    if e2e:
        # the reference's own call shape, one link: receive_byte(double* data, int* out) on one pageable double capture per call
        k = min(64, n)
        h64 = caps64[:k].cpu().numpy()
        lat, ok = [], 0
        for i in range(-3, k):
            t0 = time.perf_counter()
            out1, st1 = ts.receive_byte(h64[max(i, 0)])
            if i >= 0:
                lat.append(time.perf_counter() - t0)
                ok += int(st1["message_decoded"][0] == dec[i] and (not dec[i] or np.array_equal(out1.astype(np.uint8), pay[i])))
        single = {"median_us": float(np.median(lat) * 1e6), "p99_us": float(np.percentile(lat, 99) * 1e6), "calls": k, "identical_to_batch": ok,
                  "api": "mercury_b200_receive_byte (one capture of doubles in, ints out: H2D + front-end + tail + D2H per call)"}

    cpu = None
    if cpu_captures > 0:
        from oracle import port, ref
        k = min(cpu_captures, n)
        o, kind = (ref.Ref(config, 50), "reference") if ref.available() else (port.Port(config, 50), "port")
        secs, cdec = o.receive_byte_timed(caps64[:k].cpu().numpy())
        same = int(sum(int(cdec[i]) == int(dec[i]) for i in range(k)))
        cpu = {"value": k / secs, "unit": "captures/s", "cores": 1, "kind": kind,
               "sample": f"first {k} captures of the same batch through the {'unmodified ' if kind == 'reference' else 'restated '}receive_byte(), "
                         f"{secs:.1f} s, {int(cdec.sum())} decoded, {same}/{k} verdicts equal to the GPU's"}
    if own:
        ts.close()
    return {"metric": "captures_per_s_receive_byte_passband", "value": n / (ms * 1e-3), "unit": "captures/s", "n_gpus": 1, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "dtype": "f64 front-end (bit-exact sync decisions), f32 tail",
            "data": "synthetic",
            "config": {"workload": f"mode {config}, {n} captures of {buf} pass-band samples ({fmt}), one reference-TX frame per capture at a random "
                                   f"delay, white noise sigma {noise}", "l2_policy": f"captures {d_x.element_size() * n * buf / 1e9:.2f} GB + "
                                   f"{16 * n * buf / 1e9:.2f} GB of fp64 base-band >> 126 MB L2"},
            "gpu_launches": int(launches), "e2e": e2e_d, "single_call": single, "cpu_baseline": cpu,
            "integrity": {"decoded": int(dec.sum()), "captures": n, "payload_mismatches_among_decoded": mism,
                          "delay_error_max": int(np.abs(st["delay"][dec] - delays[dec]).max()) if dec.any() else None,
                          "sync_trials_hist": np.bincount(st["sync_trials"], minlength=4).tolist()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=8)
    ap.add_argument("--captures", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--noise", type=float, default=0.02)
    ap.add_argument("--fmt", default="f32", choices=["f32", "f64", "i16"])
    ap.add_argument("--cpu-captures", type=int, default=16)
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    print(json.dumps(run(a.config, a.captures, a.steps, a.warmup, a.noise, a.fmt, a.cpu_captures, not a.no_e2e)), flush=True)


if __name__ == "__main__":
    main()
