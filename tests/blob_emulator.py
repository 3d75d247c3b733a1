"""numpy walk-through of the table blob, mirroring the index logic of the CUDA kernels (csrc/mb_demod.cu,
csrc/mb_ldpc.cu) step by step.  It exists to pin the HOST-built tables (gather/scatter indices, JDS graph,
CRC matrices) against the oracle on a machine without a GPU; it is test code, not a fallback."""
import numpy as np

MB_NC, MB_N, MAX_CDEG, MAX_VDEG = 50, 1600, 48, 16
ZF_STRIDE, LS_COLS = 27, 18

RATE_DT = np.dtype([(n, "<i4") for n in ("rate_num", "N", "K", "P", "n_edges", "max_cdeg", "max_vdeg", "c_slots", "v_slots", "reserved")] +
                   [(n, "<u4") for n in ("off_cdeg", "off_cgbase", "off_edge_var", "off_vdeg", "off_vgbase", "off_vedge",
                                         "off_var_of_cw", "off_check_of_sorted", "off_vgdeg", "off_edge_varb", "off_vedgeb",
                                         "off_csched", "off_vsched", "off_vtail")] + [("vtail_start", "<i4")])
MODE_DT = np.dtype([(n, "<i4") for n in ("config", "M", "bps", "rate_idx", "rate_num", "Nsymb", "nData", "nPilots", "nBits",
                                         "nReal", "nVirtual", "K", "P", "frame_bytes", "estimator", "phase_only",
                                         "preamble_nSymb", "crc_bytes", "crc_reserved")] +
                   [("crc_init", "<u4"), ("boost", "<f4")] +
                   [(n, "<u4") for n in ("off_pinv", "off_pval", "off_invn", "off_pilot_cell", "off_sym_cell", "off_llr_dst",
                                         "off_llr_dst2", "off_const", "off_bit_var", "off_scr", "off_crcbit", "off_zf_src", "off_pilot_rec",
                                         "off_pilot_f", "off_data_rec", "off_virt")] +
                   [("data_rec_words", "<i4"), ("pinv_mag", "<f4"), ("off_pilot_neg", "<u4")])
HDR_DT = np.dtype([("magic", "<u4"), ("version", "<u4"), ("total_bytes", "<u4"), ("reserved", "<u4"), ("off_twiddle", "<u4"),
                   ("pad", "<u4", (3,)), ("modes", MODE_DT, (17,)), ("rates", RATE_DT, (8,))])


class Blob:
    def __init__(self, buf):
        self.b = np.ascontiguousarray(buf, np.uint8)
        self.hdr = np.frombuffer(self.b[:HDR_DT.itemsize].tobytes(), HDR_DT)[0]
        assert self.hdr["magic"] == 0x42324D42 and self.hdr["total_bytes"] == self.b.size

    def arr(self, off, dtype, n):
        dt = np.dtype(dtype)
        return np.frombuffer(self.b[int(off):int(off) + dt.itemsize * int(n)].tobytes(), dt)

    def mode(self, cfg):
        m = self.hdr["modes"][cfg]
        cells = int(m["Nsymb"]) * MB_NC
        d = {k: (int(m[k]) if k not in ("boost", "pinv_mag") else float(m[k])) for k in MODE_DT.names if not k.startswith("off_")}
        d.update(pinv=self.arr(m["off_pinv"], "<f4", cells), pval=self.arr(m["off_pval"], "<f4", cells),
                 invn=self.arr(m["off_invn"], "<f4", cells), pilot_cell=self.arr(m["off_pilot_cell"], "<u2", m["nPilots"]),
                 sym_cell=self.arr(m["off_sym_cell"], "<u2", m["nData"]), llr_dst=self.arr(m["off_llr_dst"], "<u2", m["nBits"]),
                 llr_dst2=self.arr(m["off_llr_dst2"], "<u2", m["nBits"]),
                 cons=self.arr(m["off_const"], "<f4", 2 * m["M"]).view(np.complex64),
                 bit_var=self.arr(m["off_bit_var"], "<u2", 8 * m["crc_bytes"]), scr=self.arr(m["off_scr"], "u1", MB_N),
                 crcbit=self.arr(m["off_crcbit"], "<u2", 8 * int(m["crc_bytes"])),
                 zf_src=self.arr(m["off_zf_src"], "<u4", int(m["Nsymb"]) * ZF_STRIDE),
                 pilot_rec=self.arr(m["off_pilot_rec"], "<u4", 4 * int(m["nPilots"])).reshape(-1, 4),
                 pilot_f=self.arr(m["off_pilot_f"], "<f4", 2 * int(m["nPilots"])).reshape(-1, 2),
                 data_rec=self.arr(m["off_data_rec"], "<u4", int(m["data_rec_words"]) * int(m["nData"])).reshape(int(m["nData"]), -1),
                 virt=self.arr(m["off_virt"], "<u2", 2 * int(m["nVirtual"])).reshape(-1, 2),
                 pilot_neg=self.arr(m["off_pilot_neg"], "<u8", int(m["Nsymb"])))
        return d

    def rate(self, idx):
        r = self.hdr["rates"][idx]
        d = {k: int(r[k]) for k in RATE_DT.names if not k.startswith("off_")}
        d.update(cdeg=self.arr(r["off_cdeg"], "u1", r["P"]), cgbase=self.arr(r["off_cgbase"], "<u4", 64),
                 edge_var=self.arr(r["off_edge_var"], "<u2", r["c_slots"]), vdeg=self.arr(r["off_vdeg"], "u1", r["N"]),
                 vgbase=self.arr(r["off_vgbase"], "<u4", 64), vedge=self.arr(r["off_vedge"], "<u2", r["v_slots"]),
                 var_of_cw=self.arr(r["off_var_of_cw"], "<u2", r["N"]),
                 check_of_sorted=self.arr(r["off_check_of_sorted"], "<u2", r["P"]), vgdeg=self.arr(r["off_vgdeg"], "u1", 64),
                 edge_varb=self.arr(r["off_edge_varb"], "<u2", r["c_slots"]), vedgeb=self.arr(r["off_vedgeb"], "<u2", r["v_slots"]),
                 csched=self.arr(r["off_csched"], "<u4", 8 * 16).reshape(8, 16), vsched=self.arr(r["off_vsched"], "<u4", 8 * 16).reshape(8, 16),
                 vtail=self.arr(r["off_vtail"], "<u4", r["N"] - r["vtail_start"]))
        return d

    def twiddle(self):
        return self.arr(self.hdr["off_twiddle"], "<f4", 512).view(np.complex64).reshape(16, 16)


def handoff(v):
    """position of internal variable v in the LLR hand-off vector between the two kernels (MB_HANDOFF, mb_tables.h)"""
    v = np.asarray(v, np.int64)
    return (v & ~31) | ((v + (v >> 5)) & 31)


LDPC_DMAX = 7  # MB_LDPC_DMAX (mb_tables.h)


def ldpc_split(d):
    """mb_ldpc_split (mb_tables.h): lanes per check S and edges per lane Dp of a group of 32 sorted checks whose largest degree is d"""
    s = 1
    while (d + s - 1) // s > LDPC_DMAX:
        s *= 2
    return s, (d + s - 1) // s


def cslot(r, k, cs):
    """check-side slot of edge position k of sorted check cs (mb_ldpc_cslot, mb_tables.h: warp-blocked ELL, large checks split over S lanes)"""
    S, Dp = ldpc_split(int(r["cdeg"][cs & ~31]))
    per, ci = 32 // S, cs & 31
    t, cl, j, kk = ci // per, ci % per, k // Dp, k % Dp
    return int(r["cgbase"][cs >> 5]) + (t * Dp + kk) * 32 + j * per + cl


def vslot(r, k, vs):
    return int(r["vgbase"][vs >> 5]) + 32 * k + (vs & 31)


def demod(blob, cfg, x):
    """x: [Nsymb,272] complex64 -> dict(Y,H,Z grids, llr_internal, llr_cw, variance, mean_H, snr) in float32 arithmetic."""
    m = blob.mode(cfg)
    r = blob.rate(m["rate_idx"])
    S = m["Nsymb"]
    f32, c64 = np.float32, np.complex64
    # FFT-256 as 16x16 with the blob's twiddle table (1/256 folded in): X[k1+16k2] = sum_n2 W16^(n2 k2) tw[k1,n2] A[n2,k1]
    tw = blob.twiddle()
    xs = np.asarray(x).reshape(S, 272)[:, 16:].astype(c64).reshape(S, 16, 16)  # [s, n1, n2]
    A = np.fft.fft(xs.astype(np.complex128), axis=1).astype(c64)  # over n1 -> [s, k1, n2]
    B = (A * tw[None, :, :]).astype(c64)
    X = np.fft.fft(B.astype(np.complex128), axis=2).astype(c64)   # over n2 -> [s, k1, k2]
    bins = np.zeros((S, 256), c64)
    for k1 in range(16):
        for k2 in range(16):
            bins[:, k1 + 16 * k2] = X[:, k1, k2]
    Y = np.concatenate([bins[:, 231:256], bins[:, 1:26]], axis=1).reshape(-1)
    pc = m["pilot_cell"].astype(np.int64)
    g = f32(m["boost"]) / f32(np.abs(Y[pc]).astype(f32).sum(dtype=f32) / f32(m["nPilots"]))
    # compact zero-padded pilot rows of Y/p, exactly as the kernel gathers them through zf_src
    zs = m["zf_src"].astype(np.int64)
    valid = ((zs >> 30) & 1).astype(bool)
    pinv = np.where((zs >> 31) & 1, -f32(m["pinv_mag"]), f32(m["pinv_mag"])).astype(f32)
    assert np.array_equal(np.sort((zs[valid] & 0x7FFF) // 8), np.sort(pc)) and valid.sum() == m["nPilots"]
    assert np.array_equal(pinv[valid], m["pinv"][(zs[valid] & 0x7FFF) // 8])
    zf = np.where(valid, Y[(zs & 0x7FFF) // 8] * pinv, 0).astype(c64)
    # the kernel's FFT epilogue builds the same rows from the lattice rule (pilot iff s%3 == c%3) and the per-row sign mask
    zf2 = np.zeros(S * ZF_STRIDE, c64)
    for s_ in range(S):
        for c_ in range(s_ % 3, MB_NC, 3):
            sign = -1.0 if (int(m["pilot_neg"][s_]) >> c_) & 1 else 1.0
            zf2[s_ * ZF_STRIDE + 4 + c_ // 3] = Y[s_ * MB_NC + c_] * f32(sign * m["pinv_mag"])
    assert np.array_equal(zf, zf2)
    prec, pf = m["pilot_rec"].astype(np.int64), m["pilot_f"]
    cellb, zslotb = prec[:, 3] & 0xFFFF, prec[:, 3] >> 16
    assert np.array_equal(cellb // 8, pc) and not (cellb % 8).any() and not (zslotb % 8).any()
    assert np.array_equal(pf[:, 0], m["invn"][pc]) and np.array_equal(pf[:, 1], m["pval"][pc])
    hpil = np.zeros(m["nPilots"], c64)
    if m["estimator"] == 1:
        # window sums (7 consecutive compact entries, 18 windows per row) -> running sums over the rows of each residue
        pm = np.zeros(((S + 1) * LS_COLS,), c64)
        for res in range(3):
            run = np.zeros(LS_COLS, c64)
            for k in range(res, S, 3):
                row = zf[k * ZF_STRIDE:(k + 1) * ZF_STRIDE]
                run = (run + np.array([row[jj:jj + 7].sum() for jj in range(LS_COLS)], c64)).astype(c64)
                pm[k * LS_COLS:(k + 1) * LS_COLS] = run
        acc = np.zeros(m["nPilots"], c64)
        for res in range(3):
            hi, lo = prec[:, res] & 0xFFFF, prec[:, res] >> 16
            assert not (hi % 8).any() and not (lo % 8).any() and hi.max() // 8 < S * LS_COLS and lo.max() // 8 <= S * LS_COLS
            acc = (acc + (pm[hi // 8] - pm[lo // 8])).astype(c64)
        hpil = (acc * (pf[:, 0] * g)).astype(c64)
    else:
        hpil = (zf[zslotb // 8] * g).astype(c64)
    hc = zf.copy()  # the kernel stores the channel at pilots back into the compact rows
    hc[zslotb // 8] = hpil
    H = np.zeros(S * MB_NC, c64)
    H[pc] = hpil
    is_p = np.zeros(S * MB_NC, bool)
    is_p[pc] = True
    drec = m["data_rec"].astype(np.int64)
    dcell, dzs, dt = (drec[:, 0] & 0x7FFF) // 8, ((drec[:, 0] >> 15) & 0x3FFF) // 8, (drec[:, 0] >> 29) - 2
    assert np.array_equal(dcell, np.flatnonzero(~is_p))  # grid (deframer) order
    for cell, z0, t in zip(dcell, dzs, dt):
        s, c = divmod(int(cell), MB_NC)
        r0 = (int(z0) - 4 - c // 3) // ZF_STRIDE
        assert int(z0) == r0 * ZF_STRIDE + 4 + c // 3 and t == s - r0 and is_p[r0 * MB_NC + c] and is_p[(r0 + 3) * MB_NC + c]
        a, b = hc[z0], hc[z0 + 3 * ZF_STRIDE]
        H[cell] = a + (b - a) * f32(t) / f32(3)
    mean_H = f32(np.abs(H[pc]).mean())
    Hraw = H.copy()
    if m["phase_only"]:
        H = np.where(H.real == 0, c64(1j), H / np.abs(H)).astype(c64)
    Yg = (Y * g).astype(c64)
    Z = (Yg / H).astype(c64)
    variance = f32(max((np.abs(Z[pc] - m["pval"][pc]) ** 2).mean(), 1e-30))
    v_rep = f32((np.abs(Yg[pc] / Hraw[pc] - m["pval"][pc]) ** 2).mean()) if m["phase_only"] else variance
    w = Z[dcell]
    D = (np.abs(w[:, None] - m["cons"][None, :]) ** 2).astype(f32)  # [nData, M]
    bps = m["bps"]
    L = np.full(MB_N, np.nan, f32)
    idx = np.arange(m["M"])
    for e in range(bps):  # emitted MSB first: bit mask 1 << (bps-1-e)
        k = bps - 1 - e
        one = ((idx >> k) & 1) == 1
        off = (drec[:, 1 + e // 2] >> (16 * (e & 1))) & 0xFFFF
        assert not (off % 4).any()
        L[off // 4] = (f32(1) / variance) * (D[:, one].min(axis=1) - D[:, ~one].min(axis=1))
    for src, dst in m["virt"].astype(np.int64):
        assert np.isnan(L[dst // 4]) and not np.isnan(L[src // 4])
        L[dst // 4] = L[src // 4]
    assert not np.isnan(L).any()  # every position of the internal-order vector is written exactly once
    # the grid-order records must describe the same scatter as the reference-order tables (used by the TX synthesiser)
    lam = np.zeros(m["nBits"], f32)
    q_of_cell = {int(cc): q for q, cc in enumerate(m["sym_cell"])}
    for D_i, cell in enumerate(dcell):
        q = q_of_cell[int(cell)]
        for e in range(bps):
            off = (int(drec[D_i, 1 + e // 2]) >> (16 * (e & 1))) & 0xFFFF
            assert off // 4 == int(handoff(int(m["llr_dst"][q * bps + e])))
            lam[q * bps + e] = L[off // 4]
    snr = float(10 * np.log10(1.0 / v_rep)) if m["estimator"] == 1 else 0.0
    L_int = L[handoff(np.arange(MB_N))]  # undo the row rotation of the hand-off layout
    return dict(Y=Yg.reshape(S, MB_NC), H=H.reshape(S, MB_NC), Z=Z.reshape(S, MB_NC), llr_demod=lam, llr_handoff=L, llr_internal=L_int,
                llr_cw=L_int[r["var_of_cw"].astype(np.int64)], variance=variance, mean_H=mean_H, snr=snr)


def ldpc_decode(blob, cfg, llr_internal, max_iters):
    """Flooding SPA on the JDS graph exactly as scheduled by mb_ldpc.cu (double precision tanh rule). -> (iterations, posterior)."""
    m = blob.mode(cfg)
    r = blob.rate(m["rate_idx"])
    P, N, E = r["P"], r["N"], r["c_slots"]
    cdeg, ev = r["cdeg"].astype(int), r["edge_var"].astype(int)
    vdeg, ve = r["vdeg"].astype(int), r["vedge"].astype(int)
    lch = llr_internal.astype(np.float64)
    lam = lch.copy()
    R = np.zeros(E)
    rows = [[cslot(r, k, c) for k in range(cdeg[c])] for c in range(P)]
    vrows = [[ve[vslot(r, k, v)] for k in range(vdeg[v])] for v in range(N)]
    p = 0
    while True:
        unsat = False
        newR = R.copy()
        for c in range(P):
            e = rows[c]
            v = ev[e]
            if (lam[v] < 0).sum() & 1:
                unsat = True
            q = lam[v] - R[e]
            t = np.tanh(0.5 * q)
            for i in range(len(e)):
                prod = np.prod(np.delete(t, i))
                if prod == 1:
                    prod = 0.9999999
                if prod == -1:
                    prod = -0.9999999
                newR[e[i]] = 2 * np.arctanh(prod)
        if not unsat:
            return p, lam
        if p == max_iters:
            return max_iters + 1, lam
        R = newR
        for v in range(N):
            acc = lch[v]
            for e in vrows[v]:
                acc += R[e]
            lam[v] = acc
        p += 1


def finish(blob, cfg, posterior):
    """hard decision -> de-scramble -> pack -> table-driven CRC exactly as mb_ldpc.cu. -> (bytes[crc_bytes], crc, all_zeros)."""
    m = blob.mode(cfg)
    nb = m["crc_bytes"]
    bits = (posterior[m["bit_var"].astype(int)] < 0).astype(np.uint8) ^ m["scr"][: 8 * nb]
    by = np.zeros(nb, np.uint8)
    for b in range(8):
        by |= (bits[b::8] << b).astype(np.uint8)
    adv = 0
    for i in np.nonzero(bits)[0]:  # the CRC is linear: XOR of the per-bit table over the set bits
        adv ^= int(m["crcbit"][i])
    all_zeros = int(not by.any())
    crc = 0 if all_zeros else (adv ^ m["crc_init"]) & 0xFFFF
    return by, crc, all_zeros
