#!/usr/bin/env python3
"""Instruction histogram of the kernels of libjwave_cuda.so (cuobjdump -sass), per kernel.

    python tools/sass_hist.py [--so PATH] [--match REGEX] [--top N] [--excerpt MNEMONIC[,MNEMONIC..]]

Static counts (every SASS instruction once, loops not weighted): what they show is which
instruction families a kernel contains at all (UTMALDG / UTMASTG = TMA load / store, LDGSTS =
cp.async, STG.E.ENL2.256 = 256-bit stores, DFMA with UR / c[0x0] operands = taps from the constant
bank) and the DFMA share of the unrolled bodies.  --excerpt prints the first lines that carry the
given mnemonics, as evidence for profiles/.
"""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--so", default=os.path.join(ROOT, "jwave_b200", "libjwave_cuda.so"))
    ap.add_argument("--match", default=".")
    ap.add_argument("--top", type=int, default=12)
    ap.add_argument("--excerpt", default="")
    args = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", args.so], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur:
            kernels[cur].append(m.group(1).strip())
    names = demangle(list(kernels))
    want = [w for w in args.excerpt.split(",") if w]
    for k, ins in kernels.items():
        nice = re.sub(r"\(.*", "", names.get(k, k)).replace("void ", "")
        if not re.search(args.match, nice):
            continue
        hist = collections.Counter()
        for i in ins:
            t = i.split()
            op = t[1] if t[0].startswith("@") else t[0]
            hist[op] += 1
        total = sum(hist.values())
        dfma = sum(v for o, v in hist.items() if o.startswith("DFMA") or o.startswith("DMUL") or o.startswith("DADD"))
        print(f"{nice}: {total} instructions, FP64 {dfma} ({100.0 * dfma / max(total, 1):.1f} %)")
        print("   " + ", ".join(f"{o} {v}" for o, v in hist.most_common(args.top)))
        for w in want:
            hits = [i for i in ins if w in i][:3]
            for h in hits:
                print(f"      [{w}] {h}")


if __name__ == "__main__":
    main()
