#!/usr/bin/env python3
"""CPU study: how closely do float32 check-node formulations track the reference's double-precision SPA decoder?

For each LDPC rate, frames of BPSK-equivalent channel LLRs (all-zero codeword, L = 2 (1 + sigma n) / sigma^2) are drawn
around the decoding threshold, decoded by the oracle (oracle/mercury_oracle.c: mo_ldpc_decode, the restated
ldpc_decoder_SPA.cc:25-218) and by numpy float32 emulations of candidate GPU check-node arithmetic:

  log : s = -log2 tanh(|q|/2), leave-one-out sum, R = phi(sum)          (the arithmetic of csrc/mb_ldpc.cu DECODER_SPA)
  lin : T = tanh(q/2) signed, P = prod T, R = ln((T + P) / (T - P))      (one product + one division per edge)

Reported: agreement with the reference on converged / not converged and on the iteration count.
Test infrastructure only (uses oracle/); nothing here is on the product path.

usage: python tools/ldpc_numerics.py [frames per point]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import port  # noqa: E402

f32 = np.float32
CLAMP = f32(16.811242831518264)  # 2 atanh(0.9999999)


class Graph:
    def __init__(self, lt):
        Cm = lt["C"]
        chk, var = np.nonzero(Cm >= 0)
        order = np.lexsort((np.arange(len(chk)), chk))
        self.chk, self.var = chk[order], Cm[chk[order], var[order]]
        self.P, self.N = lt["P"], lt["N"]
        self.start = np.searchsorted(self.chk, np.arange(self.P))
        self.deg = np.diff(np.append(self.start, len(self.chk)))


def decode(g, L, max_iters, kind):
    """Flooding schedule of ldpc_decoder_SPA.cc, vectorised over frames. L: [F, N] float32. -> iterations [F] (I+1 = not converged)."""
    F = L.shape[0]
    lam = L.astype(f32).copy()
    R = np.zeros((F, len(g.chk)), f32)
    iters = np.full(F, -1, np.int64)
    active = np.ones(F, bool)
    for p in range(max_iters + 1):
        idx = np.flatnonzero(active)
        if idx.size == 0:
            break
        lv = lam[idx][:, g.var]
        hard = (lv < 0).astype(np.int64)
        synd = np.add.reduceat(hard, g.start, axis=1) & 1
        ok = ~synd.any(axis=1)
        iters[idx[ok]] = p
        active[idx[ok]] = False
        if p == max_iters:
            iters[idx[~ok]] = max_iters + 1
            break
        idx = idx[~ok]
        if idx.size == 0:
            break
        q = (lam[idx][:, g.var] - R[idx]).astype(f32)
        aq = np.abs(q)
        neg = q < 0
        par = np.add.reduceat(neg.astype(np.int64), g.start, axis=1) & 1
        sign = np.where((np.repeat(par, g.deg, axis=1) ^ neg.astype(np.int64)) == 1, f32(-1), f32(1))
        if kind == "log":
            e = np.exp2((-aq * f32(1.4426950408889634)).astype(f32)).astype(f32)
            series = (e * f32(2.885390081777927) * (f32(1) + e * e * (f32(1 / 3) + e * e * f32(0.2)))).astype(f32)
            with np.errstate(divide="ignore"):
                lg = np.log2(((f32(1) + e) / (f32(1) - e)).astype(f32)).astype(f32)
            s = np.where(e < f32(0.1), series, lg).astype(f32)
            s = np.minimum(s, f32(115.0))
            s = np.where(s < f32(5.5511151231257827e-17 * 1.4426950408889634), f32(0), s)
            tot = np.add.reduceat(s.astype(np.float64), g.start, axis=1)
            so = (np.repeat(tot, g.deg, axis=1) - s).astype(f32)  # the kernel's big/rest trick = exact leave-one-out, rounded once
            so = np.maximum(so, f32(0))
            e2 = np.exp2(-so).astype(f32)
            tiny = so < f32(0.015625 * 1.4426950408889634)
            num = np.where(tiny, f32(2 * 1.4426950408889634), f32(1) + e2)
            den = np.where(tiny, so, f32(1) - e2)
            with np.errstate(divide="ignore", invalid="ignore"):
                mag = (f32(0.6931471805599453) * np.log2((num / den).astype(f32))).astype(f32)
            mag = np.where(so > 0, mag, CLAMP)
            mag = np.minimum(mag, CLAMP) if False else mag
            Rn = (sign * mag).astype(f32)
        else:
            e = np.exp2((-aq * f32(1.4426950408889634)).astype(f32)).astype(f32)
            T = ((f32(1) - e) / (f32(1) + e)).astype(f32)
            T = np.maximum(T, f32(1e-20))
            T = np.where(neg, -T, T).astype(f32)
            Pp = np.multiply.reduceat(T, g.start, axis=1).astype(f32)
            Pe = np.repeat(Pp, g.deg, axis=1)
            with np.errstate(divide="ignore", invalid="ignore"):
                ratio = ((T + Pe) / (T - Pe)).astype(f32)
                Rn = (f32(0.6931471805599453) * np.log2(ratio)).astype(f32)
            Rn = np.clip(Rn, -CLAMP, CLAMP).astype(f32)
            Rn = np.where(np.isnan(Rn), f32(0), Rn)
        R[idx] = Rn
        acc = L[idx].astype(f32).copy()
        # posterior = channel + sum of incoming messages (float32 accumulation like the kernel)
        for f_i in range(idx.size):
            acc[f_i] += np.bincount(g.var, weights=Rn[f_i].astype(np.float64), minlength=g.N).astype(f32)
        lam[idx] = acc
    return iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    rng = np.random.default_rng(1234)
    # (config whose rate we borrow, sigma values around the threshold)
    points = {0: (1, [2.45, 2.6]), 5: (6, [1.12, 1.18]), 6: (8, [0.98, 1.03]), 12: (14, [0.50, 0.53])}
    I = 50
    for cfg, (rate, sigmas) in points.items():
        o = port.Port(cfg, I)
        g = Graph(o.ldpc_tables())
        for sigma in sigmas:
            L = (2.0 * (1.0 + sigma * rng.standard_normal((n, g.N))) / sigma**2).astype(f32)
            ref = np.array([o.ldpc_decode(L[f])[0] for f in range(n)])
            line = f"rate {rate:2d}/16 sigma {sigma:.2f}: ref mean it {np.minimum(ref, I).mean():5.1f} fail {np.mean(ref > I):.3f} |"
            for kind in ("log", "lin"):
                it = decode(g, L, I, kind)
                same_conv = np.mean((it > I) == (ref > I))
                same_it = np.mean(it == ref)
                line += f" {kind}: conv-agree {same_conv:.4f} iter-agree {same_it:.4f} |"
            print(line, flush=True)


if __name__ == "__main__":
    main()
