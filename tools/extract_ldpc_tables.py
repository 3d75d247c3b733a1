#!/usr/bin/env python3
"""Derive mercury_b200/data/ldpc_tables.bin from the reference's LDPC table sources.

The Mercury LDPC codes (N=1600 IRA codes, rates {1,2,3,4,5,6,8,14}/16) are *defined* by static
integer tables in /root/reference/source/physical_layer/mercury_normal_<r>_16.cc (bound to K in
ldpc.cc:135-263).  They are interoperability data (like the parity-check matrices of a standard),
not an algorithm, and there is no generator for them in the reference.  This script reads the four
tables of every rate, checks that three of them are redundant given the check-node table
  * QCmatrixEnc[i]  == QCmatrixC[i] without the check's own parity bit K+i      (ldpc.cc:111-132)
  * QCmatrixd       == run-length encoding of the variable degrees              (ldpc_decoder_SPA.cc:106-122)
and that QCmatrixV[v] is a permutation of the checks that contain v (its order is kept: it fixes the
summation order of the posterior in ldpc_decoder_SPA.cc:162-170), and stores only what is needed to
rebuild all four tables bit-exactly:

  header : 'MLDP' u32 version(1) u32 n_rates
  rate   : u16 rate_num, N, K, P, Cwidth, Vwidth ; u32 n_edges ;
           u16 check_deg[P] ; u16 edge_var[n_edges]   (check-major, ascending inside a check)
           u16 var_deg[N]   ; u16 var_check[n_edges]  (variable-major, in the reference's V-row order)

Run here (needs /root/reference); the .bin is committed so that nothing reads the reference at run time.
"""
import os
import re
import struct
import sys

import numpy as np

REF = os.environ.get("MERCURY_REF", "/root/reference")
RATES = (1, 2, 3, 4, 5, 6, 8, 14)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mercury_b200", "data", "ldpc_tables.bin")


def parse_tables(rate):
    src = open(os.path.join(REF, "source", "physical_layer", f"mercury_normal_{rate}_16.cc")).read()
    out = {}
    for name in ("Cwidth", "Vwidth", "dwidth"):
        out[name] = int(re.search(rf"mercury_normal_{name}_{rate}_16\s*=\s*(\d+)", src).group(1))
    for name in ("QCmatrixC", "QCmatrixV", "QCmatrixd", "QCmatrixEnc"):
        m = re.search(rf"mercury_normal_{name}_{rate}_16((?:\[\d+\])+)\s*=\s*\{{(.*?)\}}\s*;", src, re.S)
        dims = [int(x) for x in re.findall(r"\[(\d+)\]", m.group(1))]
        vals = np.array([int(x) for x in re.findall(r"-?\d+", m.group(2))], dtype=np.int32)
        out[name] = vals.reshape(dims)
    return out


def pack_rate(rate):
    t = parse_tables(rate)
    N, K = 1600, 100 * rate
    P = N - K
    Cm, Vm, d, E = t["QCmatrixC"], t["QCmatrixV"], t["QCmatrixd"], t["QCmatrixEnc"]
    assert Cm.shape == (P, t["Cwidth"]) and Vm.shape == (N, t["Vwidth"]) and E.shape == (P, t["Cwidth"] - 1)
    deg, edges, rows = [], [], [[] for _ in range(N)]
    for c in range(P):
        row = [int(v) for v in Cm[c] if v != -1]
        assert row == sorted(set(row)) and len(row) >= 1 and list(Cm[c, : len(row)]) == row, "C row not ascending/unique/-1-tailed"
        assert [int(v) for v in E[c] if v != -1] == [v for v in row if v != K + c], "Enc not derivable"
        deg.append(len(row))
        edges += row
        for v in row:
            rows[v].append(c)
    var_check = []
    for v in range(N):
        vr = [int(c) for c in Vm[v] if c != -1]
        assert sorted(vr) == rows[v] and list(Vm[v, : len(vr)]) == vr, "V row is not a permutation of the checks of v"
        var_check += vr
    vdeg = [len(r) for r in rows]
    rl, i = [], 0
    while i < N:
        j = i
        while j < N and vdeg[j] == vdeg[i]:
            j += 1
        rl += [j - i, vdeg[i]]
        i = j
    assert rl == [int(x) for x in d], "d not derivable"
    blob = struct.pack("<6HI", rate, N, K, P, t["Cwidth"], t["Vwidth"], len(edges))
    blob += np.asarray(deg, "<u2").tobytes() + np.asarray(edges, "<u2").tobytes()
    blob += np.asarray(vdeg, "<u2").tobytes() + np.asarray(var_check, "<u2").tobytes()
    return blob, len(edges)


def main():
    body = b""
    for r in RATES:
        b, ne = pack_rate(r)
        print(f"rate {r}/16: {ne} edges, {len(b)} bytes")
        body += b
    with open(OUT, "wb") as f:
        f.write(b"MLDP" + struct.pack("<II", 1, len(RATES)) + body)
    print("wrote", os.path.normpath(OUT), 12 + len(body), "bytes")


if __name__ == "__main__":
    sys.exit(main())
