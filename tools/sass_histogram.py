#!/usr/bin/env python3
"""SASS instruction histogram of the hot kernels of mercury_b200/libmercury_b200.so (cuobjdump, no GPU needed).

usage: python tools/sass_histogram.py [out.txt]      (default profiles/r2_sass_histogram.txt)

Per kernel: instruction count, code bytes, and the opcode histogram (top 24 + every packed-fp32 / MUFU / bulk-copy / tensor opcode), i.e.
the evidence for what the kernels are made of: FADD2 / FMUL2 / FFMA2 (packed fp32), MUFU.*, LDS / STS widths, UBLKCP (bulk store),
and the absence of UTC*MMA / HMMA (no dense contraction on this path)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
SO = os.path.join(ROOT, "mercury_b200", "libmercury_b200.so")
HOT = ("mb_ldpc_kernel", "mb_demod_kernel", "k_mfsk_demod", "k_fe_p2b_full", "k_tx_baseband")
KEEP = re.compile(r"^(FADD2|FMUL2|FFMA2|MUFU|UBLKCP|UTMA|UTC|HMMA|IMMA|LDS|STS|LDG|STG|SHFL|BAR|ATOM|RED|DADD|DMUL|DFMA)")


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_histogram.txt")
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
    kernels, cur = [], None
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = [names[len(kernels)], collections.Counter()]
            kernels.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m and cur is not None:
            cur[1][m.group(1)] += 1
    lines = [f"SASS instruction histogram of {os.path.relpath(SO, ROOT)} (cuobjdump -sass; sm_100a); hot kernels only", ""]
    for name, hist in kernels:
        if not any(h in name for h in HOT):
            continue
        n = sum(hist.values())
        base = collections.Counter()
        for op, c in hist.items():
            base[op.split(".")[0]] += c
        lines.append(f"== {name}")
        lines.append(f"   {n} instructions, {n * 16 / 1024:.1f} KB of code")
        lines.append("   by opcode: " + ", ".join(f"{op} {c}" for op, c in base.most_common(24)))
        special = sorted((op, c) for op, c in hist.items() if KEEP.match(op))
        lines.append("   packed fp32 / MUFU / memory / sync variants: " + ", ".join(f"{op} {c}" for op, c in special))
        tensor = [op for op in hist if op.startswith(("UTC", "HMMA", "IMMA", "UTMALDG"))]
        lines.append("   tensor-core / TMA-load opcodes: " + (", ".join(tensor) if tensor else "none (no dense contraction on this path; inputs stream through LDG, the LLR vector leaves through UBLKCP)"))
        lines.append("")
    open(out_path, "w").write("\n".join(lines))
    print("\n".join(lines[:12]))
    print(f"... written to {out_path}")


if __name__ == "__main__":
    main()
