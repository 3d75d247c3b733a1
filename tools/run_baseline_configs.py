#!/usr/bin/env python3
"""Run BASELINE.json's parity-test configurations on one GPU and write profiles/<name>.json.

  #2  mode 8 (and mode 9), -I 50, 65,536 frames
  #3  mode 16, -I 20, 262,144 frames, AWGN sweep, FER/BER next to the reference on a sample of the same frames
  #4  LDPC rate sweep 1..14/16 at fixed symbol count (BPSK geometry CONFIG_0..6 + mode 12 for 14/16), 131,072 frames
  all 17 modes at threshold + 2 dB, 32,768 frames (frames/s of both stages, FER, payload integrity)

  threshold region: modes 0, 8, 13 at the Es/N0 where 10-50 % of the frames fail (both outcomes, near-threshold iteration counts)

Per line: device-resident frames/s (CUDA events), per-kernel times, demod GB/s on the algorithmic bytes of SURVEY.md 8d,
decoder edge-updates/s, FER, payload mismatches among decoded frames, and agreement with the reference CPU path (oracle/_ref if present,
else the C port) on >= 8,192 frames per configuration (the first 4,096 + 4,096 random frames of the batch), decoded by the
reference on all host cores (one forked worker per core, each with its own reference object): decoded flag, payload, iteration count,
with every disagreeing frame listed.  Not a bench line; bench.py keeps the driver's contract.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mercury_b200 as mb  # noqa: E402


_W = {}
_X = None


def _ref_init(cfg, iters):
    from oracle import port, ref
    _W["o"] = ref.Ref(cfg, iters) if ref.available() else port.Port(cfg, iters)


def _ref_job(job):
    lo, hi = job
    _, pay, dec, its = _W["o"].rx_tail_timed(_X[lo:hi])
    return lo, pay, dec, its


def reference_compare(cfg, iters, xs, pay_gpu, st_gpu, idx):
    """Decode xs (complex64 [n, S, 272], frames idx of the batch) with the reference on all cores and compare with the GPU's results."""
    global _X
    import multiprocessing as mp
    from oracle import ref
    n = xs.shape[0]
    _X = xs.astype(np.complex128)
    cores = os.cpu_count() or 1
    step = max(16, (n + 8 * cores - 1) // (8 * cores))
    jobs = [(lo, min(n, lo + step)) for lo in range(0, n, step)]
    pay = np.zeros((n, pay_gpu.shape[1]), np.uint8)
    dec, its = np.zeros(n, np.int32), np.zeros(n, np.int32)
    with mp.get_context("fork").Pool(cores, initializer=_ref_init, initargs=(cfg, iters)) as pool:
        for lo, p, d, i in pool.imap_unordered(_ref_job, jobs):
            pay[lo:lo + len(d)] = p[:, :pay_gpu.shape[1]].astype(np.uint8)
            dec[lo:lo + len(d)], its[lo:lo + len(d)] = d, i
    _X = None
    g_dec, g_its = st_gpu["message_decoded"][idx], st_gpu["iterations_done"][idx]
    both = (dec == 1) & (g_dec == 1)
    bad_dec = np.flatnonzero(dec != g_dec)
    bad_its = np.flatnonzero(its != g_its)
    bad_pay = np.flatnonzero(both & (pay != pay_gpu[idx]).any(axis=1))
    return {"kind": "reference" if ref.available() else "port", "frames": int(n), "cores": cores, "reference_decoded": int((dec == 1).sum()),
            "decoded_flag_agrees": int(n - bad_dec.size), "payload_bit_exact_where_reference_decodes": int(((dec == 1) & (g_dec == 1)).sum() - bad_pay.size),
            "iteration_count_agrees": int(n - bad_its.size), "reference_fer": float(1 - (dec == 1).mean()),
            "reference_mean_iterations": float(np.minimum(its, iters).mean()),
            "disagreements": [{"frame": int(idx[f]), "ref_decoded": int(dec[f]), "gpu_decoded": int(g_dec[f]), "ref_iterations": int(its[f]),
                               "gpu_iterations": int(g_its[f])} for f in sorted(set(bad_dec.tolist()) | set(bad_its.tolist()) | set(bad_pay.tolist()))[:64]]}


def run_one(ts, cfg, B, iters, esn0, steps, sample, decoder="spa"):
    import torch
    from oracle import port, ref
    dev = torch.device("cuda", 0)
    geom = ts.load_configuration(cfg, iters)
    ts.set_decoder(mb.DECODER_SPA if decoder == "spa" else mb.DECODER_MINSUM)
    m = mb.MODES[cfg]
    fb = geom["frame_bytes"]
    d_x, pl = bench.synth_batch_on_device(cfg, B, esn0, dev, seed=1000 + cfg)
    d_pay = torch.zeros((B, fb), dtype=torch.uint8, device=dev)
    d_st = torch.zeros((B, 32), dtype=torch.uint8, device=dev)
    d_llr = torch.empty((B, mb.HANDOFF_FLOATS), dtype=torch.float32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    td = tl = 0.0
    for i in range(steps + 2):
        ev[0].record()
        ts.demod_batch_device(d_x, B, d_llr, d_st, None, stream=s)
        ev[1].record()
        ts.ldpc_decode_batch_device(d_llr, B, d_pay, d_st, stream=s)
        ev[2].record()
        torch.cuda.synchronize()
        if i >= 2:
            td += ev[0].elapsed_time(ev[1]) * 1e-3
            tl += ev[1].elapsed_time(ev[2]) * 1e-3
    td, tl = td / steps, tl / steps
    st = d_st.cpu().numpy().view(mb.STATS_DTYPE).reshape(-1)
    pay = d_pay.cpu().numpy()
    dec = st["message_decoded"] == 1
    its = np.clip(st["iterations_done"], 0, iters)
    bit_err = int(np.unpackbits(pay ^ pl, axis=1).sum())
    out = {
        "config": cfg, "decoder": decoder, "frames": B, "ldpc_iters": iters, "esn0_db": esn0,
        "frames_per_s": B / (td + tl), "demod_ms": td * 1e3, "ldpc_ms": tl * 1e3,
        "demod_gbs_algorithmic": m["demod_bytes"] * B / td / 1e9, "demod_frames_per_s": B / td,
        "ldpc_edge_updates_per_s": float(its.sum()) * m["edges"] / tl, "ldpc_frames_per_s": B / tl,
        "mean_iterations": float(its.mean()), "fer": float(1 - dec.mean()), "ber_payload": bit_err / (pay.size * 8.0),
        "payload_mismatches_among_decoded": int((pay[dec] != pl[dec]).any(axis=1).sum()),
    }
    if sample:
        sample = min(sample, B)
        head = min(sample // 2, B)
        rng = np.random.default_rng(4242 + cfg)
        rest = rng.choice(np.arange(head, B), size=sample - head, replace=False) if B > head else np.zeros(0, np.int64)
        idx = np.concatenate([np.arange(head), np.sort(rest)]).astype(np.int64)
        xs = d_x[torch.from_numpy(idx).to(dev)].cpu().numpy()
        out["reference_sample"] = reference_compare(cfg, iters, xs, pay, st, idx)
    del d_x, d_llr
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_baseline_configs.json"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated subset of: config2,config3,config4,threshold_region,all_modes")
    a = ap.parse_args()
    ts = mb.TelecomSystemB200(0)
    q = 8 if a.quick else 1
    n_ref = 8192 // q
    res = {"config2": [], "config3": [], "config4": [], "threshold_region": [], "all_modes": []}
    only = set(x for x in a.only.split(",") if x)
    want = lambda k: not only or k in only
    for cfg in (8, 9) if want("config2") else ():
        res["config2"].append(run_one(ts, cfg, 65536 // q, 50, mb.THRESH_DB[cfg] + 2.0, 5, n_ref))
        res["config2"].append(run_one(ts, cfg, 65536 // q, 50, mb.THRESH_DB[cfg] + 2.0, 5, 0, decoder="minsum"))
    for esn0 in (18.0, 20.0, 22.0, 25.0, 30.0) if want("config3") else ():
        res["config3"].append(run_one(ts, 16, 262144 // q, 20, esn0, 3, n_ref if esn0 in (18.0, 22.0) else n_ref // 4))
    for cfg in (0, 1, 2, 3, 4, 5, 6, 12) if want("config4") else ():
        res["config4"].append(run_one(ts, cfg, 131072 // q, 50, mb.THRESH_DB[cfg] + 2.0, 3, n_ref))
    # where 10-50 % of the frames fail: the fp32 decoder against the double-precision reference on both outcomes
    for cfg, off in ((0, 0.1), (0, -0.1), (8, -0.65), (8, -0.45), (13, -0.9), (13, -0.7)) if want("threshold_region") else ():
        res["threshold_region"].append(run_one(ts, cfg, 16384 // q, 50, mb.THRESH_DB[cfg] + off, 3, n_ref))
    for cfg in range(17) if want("all_modes") else ():
        esn0 = mb.THRESH_DB[cfg] + (2.0 if cfg < 15 else 14.0)
        res["all_modes"].append(run_one(ts, cfg, 32768 // q, 20 if cfg == 16 else 50, esn0, 3, 256 // q))
    res = {k: v for k, v in res.items() if v}
    json.dump(res, open(a.out, "w"), indent=1)
    for k, v in res.items():
        print(k)
        for r in v:
            rs = r.get("reference_sample", {})
            print(f"  mode {r['config']:2d} {r['decoder']:6s} B={r['frames']:6d} EsN0={r['esn0_db']:5.1f} {r['frames_per_s'] / 1e6:6.2f} Mf/s "
                  f"demod {r['demod_gbs_algorithmic']:6.0f} GB/s ldpc {r['ldpc_edge_updates_per_s'] / 1e9:6.1f} Ge/s it {r['mean_iterations']:5.2f} "
                  f"FER {r['fer']:.4f} mism {r['payload_mismatches_among_decoded']} "
                  f"ref[{rs.get('frames', 0)}]: dec= {rs.get('decoded_flag_agrees', '-')} pay= {rs.get('payload_bit_exact_where_reference_decodes', '-')}/{rs.get('reference_decoded', '-')} it= {rs.get('iteration_count_agrees', '-')} ref-FER {rs.get('reference_fer', float('nan')):.3f}")
            for d in rs.get("disagreements", [])[:8]:
                print("      disagreement:", d)


if __name__ == "__main__":
    main()
