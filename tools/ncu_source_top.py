#!/usr/bin/env python3
"""Top stall sites of an `ncu --page source --csv` export: per kernel, the SASS instructions with the most
warp-stall samples and the split of all samples by instruction class and by stall reason.

    python tools/ncu_source_top.py gpurun_out/<tag>.source.csv [N]
"""
import collections
import csv
import sys


def main():
    path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 15
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
    names = rows[hdr]
    col = {n: i for i, n in enumerate(names)}
    stalls = [n for n in names if n.startswith("stall_")]
    blocks, cur = [], []
    for r in rows[hdr + 1:]:
        if len(r) < len(names) or r == names:
            if cur:
                blocks.append(cur)
                cur = []
            continue
        cur.append(r)
    if cur:
        blocks.append(cur)
    for bi, blk in enumerate(blocks):
        def f(r, n):
            try:
                return float(r[col[n]])
            except Exception:
                return 0.0
        total = sum(f(r, "# Samples") for r in blk) or 1.0
        by_op, by_reason = collections.Counter(), collections.Counter()
        for r in blk:
            t = r[col["Source"]].split()
            if not t:
                continue
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            by_op[op] += f(r, "# Samples")
            for s in stalls:
                by_reason[s] += f(r, s)
        print(f"== block {bi}: {len(blk)} instructions, {total:.0f} samples")
        print("   by opcode : " + ", ".join(f"{o} {100 * v / total:.1f}%" for o, v in by_op.most_common(8)))
        rs = sum(by_reason.values()) or 1.0
        print("   by reason : " + ", ".join(f"{o[6:]} {100 * v / rs:.1f}%" for o, v in by_reason.most_common(8)))
        for r in sorted(blk, key=lambda r: -f(r, "# Samples"))[:top]:
            why = sorted(((f(r, s), s[6:]) for s in stalls), reverse=True)[:2]
            print(f"   {100 * f(r, '# Samples') / total:5.2f}%  {r[col['Address']][-5:]}  {r[col['Source']][:70]:70s} "
                  + ", ".join(f"{n} {v:.0f}" for v, n in why if v > 0))


if __name__ == "__main__":
    main()
