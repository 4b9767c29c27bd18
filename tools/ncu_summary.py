#!/usr/bin/env python3
"""Markdown table of the rows of an `ncu --page raw --csv` export (one row per profiled launch).

    python tools/ncu_summary.py gpurun_out/prof_<tag>.raw.csv [--json out.json]

Columns: duration, DRAM bytes read / written, FP64 pipe % (sm__pipe_fp64_cycles_active of peak sustained
active), issue slots used, registers, dynamic shared memory, CTA / grid size, warps active %, shared-memory
wavefronts and bank conflicts (and their ratio), the four largest warp-stall reasons per issued instruction.
--json also writes {kernel: dram bytes per launch} for profiles/traffic.json.
"""
import csv
import json
import re
import sys


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return float("nan")


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    col = {n: i for i, n in enumerate(names)}
    stall = [n for n in names if re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", n)]

    def get(r, n):
        return num(r[col[n]]) if n in col and col[n] < len(r) else float("nan")

    def scaled(r, n, to):
        """value of a byte / time metric in `to` units"""
        v, u = get(r, n), units[col[n]] if n in col else ""
        f = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
             "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "second": 1.0}
        return v * f.get(u, 1.0) / f[to]

    print("| kernel | time ms | dram rd GB | dram wr GB | FP64 pipe % | issue % | regs | dyn smem KB | block | grid | "
          "warps active % | smem wavefronts | bank conflicts | conflicts / wavefronts | top stalls (warps per issue) |")
    print("|" + "---|" * 15)
    traffic = {}
    for r in data:
        if len(r) <= col["Kernel Name"]:
            continue
        k = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("jwc::", "")
        st = sorted(((get(r, n), re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue", n).group(1)) for n in stall),
                    reverse=True)
        st = [(v, n) for v, n in st if n != "selected"][:4]
        wf = get(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        bc = get(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
        rd, wr = scaled(r, "dram__bytes_read.sum", "Gbyte"), scaled(r, "dram__bytes_write.sum", "Gbyte")
        traffic.setdefault(k, []).append((rd + wr) * 1e9)
        print(f"| `{k}` | {scaled(r, 'gpu__time_duration.sum', 'ms'):.4g} | {rd:.4g} | {wr:.4g} | "
              f"{get(r, 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{get(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{get(r, 'launch__registers_per_thread'):.0f} | {scaled(r, 'launch__shared_mem_per_block_dynamic', 'Kbyte'):.1f} | "
              f"{get(r, 'launch__block_size'):.0f} | {get(r, 'launch__grid_size'):.0f} | "
              f"{get(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {wf:.4g} | {bc:.4g} | "
              f"{(bc / wf if wf else float('nan')):.3f} | " + ", ".join(f"{n} {v:.2f}" for v, n in st) + " |")
    if "--json" in sys.argv:
        out = sys.argv[sys.argv.index("--json") + 1]
        json.dump({k: max(v) for k, v in traffic.items()}, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
