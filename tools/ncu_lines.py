#!/usr/bin/env python3
"""Per-source-line dynamic instruction counts and stall samples from an .ncu-rep captured with --import-source on.

usage: python tools/ncu_lines.py gpurun_out/prof.ncu-rep mercury_b200/libmercury_b200.so <kernel substring> [top N] [mangled substring]

The ncu source page lists SASS instructions in cubin order; nvdisasm -g gives the source line of each SASS instruction
of the same cubin.  The two are joined by instruction index inside the kernel.  Needs ncu, cuobjdump, nvdisasm (no GPU).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, kern):
    """[(sass text, file, line)] for the first kernel whose mangled name contains `kern`."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True, check=True)
    out = []
    for cub in sorted(os.listdir(tmp)):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
        cur, inside, f, l = None, False, "?", 0
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
            if m:
                inside = kern in m.group(1) and not out
                cur = m.group(1)
                continue
            if not inside:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                f, l = os.path.basename(m.group(1)), int(m.group(2))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out.append((m.group(2).strip(), f, l))
        if out:
            break
    return out


def main():
    rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # split per kernel
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    blk = next(b for b in blocks if kern in b["name"])
    hdr = blk["hdr"]
    i_exec, i_samp, i_src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    sl = sass_lines(so, sys.argv[5] if len(sys.argv) > 5 else kern.split("<")[0].split("::")[-1])
    n = min(len(sl), len(blk["rows"]))
    if len(sl) != len(blk["rows"]):
        print(f"warning: {len(sl)} SASS instructions in the cubin vs {len(blk['rows'])} in the report")
    per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot_i = tot_s = 0
    for k in range(n):
        r = blk["rows"][k]
        key = (sl[k][1], sl[k][2])
        ex, sm = int(r[i_exec] or 0), int(r[i_samp] or 0)
        per[key][0] += ex
        per[key][1] += sm
        for c in stall_cols:
            v = int(r[c] or 0)
            if v:
                per[key][2][hdr[c]] += v
        tot_i += ex
        tot_s += sm
    print(f"kernel {blk['name']}: {tot_i} warp instructions, {tot_s} samples")
    print("by source line (top by instructions):")
    for key, (ex, sm, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        s3 = ", ".join(f"{a[6:]}={b}" for a, b in st.most_common(3))
        print(f"  {key[0]}:{key[1]:<5d} inst {ex:>11d} {100.0 * ex / tot_i:5.1f}%   samples {sm:>7d} {100.0 * sm / max(tot_s, 1):5.1f}%   {s3}")
    print("by source line, in line order:")
    for key in sorted(per):
        ex, sm, st = per[key]
        print(f"  {key[0]}:{key[1]:<5d} inst {100.0 * ex / tot_i:5.1f}%  samples {100.0 * sm / max(tot_s, 1):5.1f}%")


if __name__ == "__main__":
    main()
