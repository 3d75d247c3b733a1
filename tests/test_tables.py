"""Host-built table blob (csrc/mb_tables.cpp) pinned against the oracle on CPU: geometry of all 17 modes, the
composed gather/scatter indices, the JDS Tanner graphs and the CRC matrices, by walking the blob with the same
index logic as the kernels (tests/blob_emulator.py)."""
import numpy as np
import pytest

import mercury_b200 as mb
from oracle import port
from tests import blob_emulator as be

THRESH = mb.THRESH_DB


@pytest.fixture(scope="module")
def blob():
    return be.Blob(mb.build_tables_host())


@pytest.mark.parametrize("cfg", range(17))
def test_geometry_matches_oracle_and_static_mode_list(blob, cfg):
    m, p, s = blob.mode(cfg), port.Port(cfg), mb.MODES[cfg]
    for k_blob, k_port in (("M", "M"), ("Nsymb", "Nsymb"), ("nData", "nData"), ("nPilots", "nPilots"), ("nBits", "nBits"),
                           ("K", "K"), ("P", "P"), ("frame_bytes", "frame_bytes"), ("estimator", "estimator"),
                           ("phase_only", "amp_restore"), ("preamble_nSymb", "preamble_nSymb")):
        assert m[k_blob] == p.geom[k_port], k_blob
    for k in ("M", "Nsymb", "nData", "nPilots", "nBits", "K", "P", "frame_bytes", "nReal", "nVirtual", "estimator", "phase_only", "bps"):
        assert m[k] == s[k], k
    t = p.tables()
    mask = np.zeros(m["Nsymb"] * 50, bool)
    mask[m["pilot_cell"].astype(int)] = True
    assert np.array_equal(mask.reshape(m["Nsymb"], 50), t["carrier_type"] == 1)
    assert np.array_equal(m["pval"][mask].astype(np.float64), t["pilot_seq"].astype(np.float32).astype(np.float64))
    assert np.allclose(m["pinv"][mask] * m["pval"][mask], 1.0, rtol=1e-6) and not m["pinv"][~mask].any()
    assert np.array_equal(m["cons"], t["constellation"].astype(np.complex64))
    assert np.array_equal(m["scr"], t["scrambler"].astype(np.uint8))


@pytest.mark.parametrize("rate_idx", range(8))
def test_jds_graph_is_the_reference_graph(blob, rate_idx):
    r = blob.rate(rate_idx)
    cfg = {1: 0, 2: 1, 3: 2, 4: 3, 5: 4, 6: 5, 8: 6, 14: 12}[r["rate_num"]]
    lt = port.Port(cfg).ldpc_tables()
    cw_of_var = np.zeros(1600, int)
    cw_of_var[r["var_of_cw"].astype(int)] = np.arange(1600)
    assert sorted(r["var_of_cw"].tolist()) == list(range(1600))
    assert (np.diff(r["cdeg"].astype(int)) <= 0).all() and (np.diff(r["vdeg"].astype(int)) <= 0).all()
    slot_of = {}
    for cs in range(r["P"]):
        c = int(r["check_of_sorted"][cs])
        ref_row = [int(v) for v in lt["C"][c] if v != -1]
        row = [int(cw_of_var[r["edge_var"][be.cslot(r, k, cs)]]) for k in range(int(r["cdeg"][cs]))]
        assert sorted(row) == sorted(ref_row)  # the reference's row; its ORDER is the layout's (mercury_b200/data/ldpc_layout.bin: bank-friendly)
        for k, v in enumerate(row):
            slot_of[(c, v)] = be.cslot(r, k, cs)
    assert len(slot_of) == r["n_edges"] == len(set(slot_of.values())) and max(slot_of.values()) < r["c_slots"]
    assert (r["edge_var"] != 0xFFFF).sum() == r["n_edges"] == (r["vedge"] != r["c_slots"]).sum()
    assert all(int(r["vgdeg"][g]) == (int(r["vdeg"][32 * g]) + 1) // 2 * 2 for g in range(50))
    assert r["c_slots"] <= 1.23 * r["n_edges"] and r["v_slots"] <= 1.25 * r["n_edges"]  # padding stays small
    # the decoder kernel's byte-offset copies, neutral padding, degree-<=2 tail and static warp schedules
    real = r["edge_var"] != 0xFFFF
    assert np.array_equal(r["edge_varb"][real], r["edge_var"][real] * 8) and (r["edge_varb"][~real] == 8 * 1600).all()
    assert np.array_equal(r["vedgeb"].astype(int), r["vedge"].astype(int) * 8)
    t0 = r["vtail_start"]
    assert t0 % 32 == 0 and (r["vdeg"][t0:] <= 2).all() and (t0 < 32 or r["vdeg"][t0 - 32] > 2)
    for i, w in enumerate(r["vtail"].astype(int)):
        offs = [int(r["vedge"][be.vslot(r, k, t0 + i)]) * 8 for k in range(int(r["vdeg"][t0 + i]))] + [8 * r["c_slots"]] * 2
        assert (w & 0xFFFF, w >> 16) == (offs[0], offs[1])
    # check tasks (MB_CDESC_*: base | Dp << 16 | log2 S << 20 | task << 22 | (group + 1) << 25): every task of every group exactly once
    ent = [int(e) for row in r["csched"] for e in row[:-1] if e != 0]
    want = []
    for g in range((r["P"] + 31) // 32):
        S, Dp = be.ldpc_split(int(r["cdeg"][32 * g]))
        for t in range(S):
            if g * 32 + t * (32 // S) < r["P"]:
                want.append((int(r["cgbase"][g]) + t * Dp * 32) | Dp << 16 | (S.bit_length() - 1) << 20 | t << 22 | (g + 1) << 25)
    assert sorted(ent) == sorted(want) and all(2 <= (e >> 16) & 0xF <= be.LDPC_DMAX for e in ent)
    body = lambda e: ((e >> 16) & 0xF) + (1 if (e >> 20) & 3 else 0)   # MB_CDESC_BODY: a split task runs the body of one more edge
    for row in r["csched"]:   # sorted by body, 0-terminated, last word = tasks per body (4 bits each from body 2)
        tasks = [int(e) for e in row[:-1] if e != 0]
        assert [body(e) for e in tasks] == sorted(body(e) for e in tasks) and int(row[len(tasks)]) == 0 and len(tasks) <= 14
        assert int(row[-1]) == sum(1 << (4 * (body(e) - 2)) for e in tasks)
    ccost = lambda e: 26 if (e >> 16) & 0xF <= 2 else (25 * ((e >> 16) & 0xF) + 15 if (e >> 20) & 3 == 0 else 37 * ((e >> 16) & 0xF) + 55)
    loads = [sum(ccost(int(e)) for e in row[:-1] if e != 0) for row in r["csched"]]
    assert max(loads) - min(loads) <= max(ccost(e) for e in ent)  # LPT balance, in the builder's cost model
    var_cost = lambda d: 3 * d + 8
    for sched, n_groups, weight, base, cost in ((r["vsched"], t0 // 32, lambda g: int(r["vgdeg"][g]), r["vgbase"], var_cost),):
        ent = [int(e) for row in sched for e in row if e != 0]
        assert sorted((e >> 24) - 1 for e in ent) == list(range(n_groups))     # every group exactly once
        assert all(e & 0xFFFF == int(base[(e >> 24) - 1]) and (e >> 16) & 0xFF == weight((e >> 24) - 1) for e in ent)
        assert all(row[-1] == 0 for row in sched)                              # terminated
        loads = [sum(cost((int(e) >> 16) & 0xFF) for e in row if e != 0) for row in sched]
        assert max(loads) - min(loads) <= max(cost(weight(g)) for g in range(n_groups))  # LPT balance, in the builder's cost model
    for vi in range(1600):
        v = int(cw_of_var[vi])
        ref_row = [int(c) for c in lt["V"][v] if c != -1]
        assert int(r["vdeg"][vi]) == len(ref_row)
        assert sorted(int(r["vedge"][be.vslot(r, k, vi)]) for k in range(len(ref_row))) == sorted(slot_of[(c, v)] for c in ref_row)


@pytest.mark.parametrize("cfg", range(17))
def test_demod_index_tables_reproduce_oracle_llrs(blob, cfg, golden_dir):
    g = np.load(f"{golden_dir}/rx_mode{cfg:02d}.npz")
    e = be.demod(blob, cfg, g["x"])
    for k in ("Y", "H", "Z"):
        ref = g[k]
        assert np.abs(e[k] - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max()), k
    if cfg < 15:
        ref = g["llr_cw"]
        tol = 1e-4 * np.maximum(np.abs(ref), np.median(np.abs(ref)))
        assert (np.abs(e["llr_cw"] - ref) <= tol).all()
        assert abs(e["snr"] - float(g["snr"])) < 1e-3 and abs(e["mean_H"] - float(g["mean_H"])) < 1e-4
    else:  # ZF: the variance is rounding residue, compare what the decoder can see: the signs
        nz = np.abs(g["llr_cw"]) > 0
        assert np.array_equal(np.signbit(e["llr_cw"][nz]), np.signbit(g["llr_cw"][nz]))


@pytest.mark.parametrize("cfg", [0, 5, 8, 10, 13, 14, 16])
def test_decoder_schedule_and_crc_tables(blob, cfg, golden_dir):
    g = np.load(f"{golden_dir}/rx_mode{cfg:02d}.npz")
    r = blob.rate(blob.mode(cfg)["rate_idx"])
    L = np.zeros(1600, np.float32)
    L[r["var_of_cw"].astype(int)] = g["llr_cw"]
    its, post = be.ldpc_decode(blob, cfg, L, int(g["ldpc_iters"]))
    assert its == int(g["iterations"])
    by, crc, az = be.finish(blob, cfg, post)
    assert np.array_equal(by, g["bytes"]) and crc == int(g["crc"]) == 0 and az == int(g["all_zeros"])
    # a corrupted byte must produce the reference's CRC value through the warp-parallel matrices too
    m = blob.mode(cfg)
    post2 = post.copy()
    post2[m["bit_var"][13]] *= -1
    by2, crc2, _ = be.finish(blob, cfg, post2)
    assert crc2 == port.port_crc16(by2.tolist()) != 0


def _gather_wavefronts(r):
    """Shared-memory wavefronts per warp gather of the decoder, from the blob: lane i of a check task (variable group) reads the float2
    (pair of frames) posterior[edge_var] (message[vedge]) at step k.  A 64-bit access is served half-warp by half-warp over 16 eight-byte
    banks: distinct words in one bank serialise, the same word is a broadcast; the minimum is 2.  -> (check side, variable side)."""
    def rows_cost(tab, starts):
        tot = 0
        for st in starts:
            row = [1600 if int(w) == 0xFFFF else int(w) for w in tab[st:st + 32]]
            for half in (row[:16], row[16:]):
                if half:
                    tot += max(np.bincount([w % 16 for w in set(half)], minlength=16))
        return tot / len(starts)
    csteps = []
    for g in range((r["P"] + 31) // 32):
        S, Dp = be.ldpc_split(int(r["cdeg"][32 * g]))
        csteps += [int(r["cgbase"][g]) + 32 * i for i in range(S * Dp)]
    vsteps = [int(r["vgbase"][g]) + 32 * k for g in range((r["N"] + 31) // 32) for k in range(int(r["vgdeg"][g]))]
    return [rows_cost(r["edge_var"].astype(int), csteps), rows_cost(r["vedge"].astype(int), vsteps)]


def test_layout_file_only_permutes_and_lowers_bank_conflicts(tmp_path):
    """mercury_b200/data/ldpc_layout.bin (tools/ldpc_layout_opt.cpp): the same graphs (test_jds_graph_is_the_reference_graph runs on it), fewer
    shared-memory bank conflicts in the decoder's gathers than the reference order; a file that is not a permutation is rejected."""
    import os
    import shutil
    from mercury_b200 import _lib
    d = str(tmp_path)
    shutil.copy(_lib.LDPC_TABLES, os.path.join(d, "ldpc_tables.bin"))
    plain = be.Blob(mb.build_tables_host(os.path.join(d, "ldpc_tables.bin")))   # no layout file next to it: reference order
    tuned = be.Blob(mb.build_tables_host())
    for idx in (0, 5, 7):
        (c0, v0), (c1, v1) = _gather_wavefronts(plain.rate(idx)), _gather_wavefronts(tuned.rate(idx))
        assert c1 < 0.6 * c0 and v1 < 0.75 * v0 and c1 < 2.6 and v1 < 3.9, (idx, c0, v0, c1, v1)
        assert plain.rate(idx)["c_slots"] == tuned.rate(idx)["c_slots"] and plain.rate(idx)["v_slots"] == tuned.rate(idx)["v_slots"]
    lay = bytearray(open(os.path.join(os.path.dirname(_lib.LDPC_TABLES), "ldpc_layout.bin"), "rb").read())
    lay[12 + 12 + 2 * 1600 + 2 * 1500 + 10] ^= 1   # an edge of a check of the first rate now names another variable
    open(os.path.join(d, "ldpc_layout.bin"), "wb").write(bytes(lay))
    with pytest.raises(mb.MercuryB200Error):
        mb.build_tables_host(os.path.join(d, "ldpc_tables.bin"))


def test_qam_demap_decompositions_equal_the_brute_force(blob):
    """mb_demod.cu demap_scatter_qam16 / demap_scatter_qam32: the max-log LLR numerators (Dmin1 - Dmin0, psk.cc:278-326) from per-axis minima
    over the rectangles the 16QAM / 32QAM index mapping is made of, against the minimum over all constellation points of the table."""
    rng = np.random.default_rng(5)
    z = (rng.standard_normal(40000) + 1j * rng.standard_normal(40000)) * 0.9
    mn = np.minimum

    def brute(c, nbits):
        D = np.abs(z[:, None] - c[None, :].astype(np.complex128)) ** 2
        out = np.zeros((z.size, nbits))
        for k in range(nbits):
            one = ((np.arange(c.size) >> k) & 1) == 1
            out[:, k] = D[:, one].min(1) - D[:, ~one].min(1)
        return out

    c = blob.mode(13)["cons"].astype(np.complex128)          # 16QAM: index = a * 4 + b, in-phase level from c[4 a], quadrature from c[b]
    dI = [(z.real - c[4 * a].real) ** 2 for a in range(4)]
    dQ = [(z.imag - c[b].imag) ** 2 for b in range(4)]
    fast = np.stack([mn(dQ[1], dQ[3]) - mn(dQ[0], dQ[2]), mn(dQ[2], dQ[3]) - mn(dQ[0], dQ[1]),
                     mn(dI[1], dI[3]) - mn(dI[0], dI[2]), mn(dI[2], dI[3]) - mn(dI[0], dI[1])], axis=1)
    assert np.abs(fast - brute(c, 4)).max() < 1e-12

    c = blob.mode(16)["cons"].astype(np.complex128)          # 32QAM: index 9 is (-1, +1) u
    u = c[9].imag
    assert u > 0 and c[9].real == -u
    X = {l: (z.real - l * u) ** 2 for l in (-5, -3, -1, 1, 3, 5)}
    Y = {l: (z.imag - l * u) ** 2 for l in (-5, -3, -1, 1, 3, 5)}
    Am, Ap = mn(X[-3], X[-1]), mn(X[3], X[1])
    Bm, Bp = mn(Am, X[-5]), mn(Ap, X[5])
    A, B, C = mn(Am, Ap), mn(Bm, Bp), mn(mn(X[-5], X[-3]), mn(X[5], X[3]))
    X3, X1, X5 = mn(X[-3], X[3]), mn(X[-1], X[1]), mn(X[-5], X[5])
    Yb1_0, Yb1_1, Yb0_0, Yb0_1 = mn(Y[3], Y[1]), mn(Y[-3], Y[-1]), mn(Y[3], Y[-3]), mn(Y[1], Y[-1])
    Yc, Ycap = mn(Yb1_0, Yb1_1), mn(Y[5], Y[-5])
    fast = np.stack([mn(X1 + Ycap, B + Yb0_1) - mn(X3 + Ycap, B + Yb0_0), mn(A + Y[-5], B + Yb1_1) - mn(A + Y[5], B + Yb1_0),
                     (C + Yc) - mn(A + Ycap, X1 + Yc), (A + Yc) - mn(A + Ycap, X5 + Yc), mn(Ap + Ycap, Bp + Yc) - mn(Am + Ycap, Bm + Yc)], axis=1)
    assert np.abs(fast - brute(c, 5)).max() < 1e-6   # the table's levels are float32 roundings of 3 u and 5 u, the kernel multiplies u
