#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per CUDA source line:
instructions executed, stall samples and the dominant stall reasons.  usage: ncu_lines.py export.csv [top]"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    csv.field_size_limit(1 << 30)
    cur_file = "?"
    hdr = None
    agg = defaultdict(lambda: defaultdict(float))
    text = {}
    line_no = None
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 2:
            continue
        if r[0].strip():
            line_no = int(r[0])
            text[(cur_file, line_no)] = r[1].strip()[:110]
        key = (cur_file, line_no)
        for name, v in zip(hdr[4:], r[4:]):
            try:
                agg[key][name] += float(v)
            except ValueError:
                pass
    tot_s = sum(a["# Samples"] for a in agg.values())
    tot_i = sum(a["Instructions Executed"] for a in agg.values())
    print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
    stall_names = [n for n in hdr[4:] if n.startswith("stall_") and "Not Issued" not in n]
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
        st = sorted(((a[n], n[6:]) for n in stall_names), reverse=True)[:3]
        print(f"{100 * a['# Samples'] / tot_s:5.1f}% smp {100 * a['Instructions Executed'] / max(tot_i, 1):5.1f}% ins  {key[0]}:{key[1]:<4} "
              f"{' '.join(f'{n}={100 * v / max(a['# Samples'], 1):.0f}%' for v, n in st if v > 0):40s} | {text.get(key, '')}")


if __name__ == "__main__":
    main()
