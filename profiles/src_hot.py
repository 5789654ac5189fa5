#!/usr/bin/env python
"""Summarise an `ncu --page source --print-source cuda,sass --csv` dump: hottest CUDA lines by executed
warp instructions and by stall samples.  usage: src_hot.py dump.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No" and "Instructions Executed" in r)
hdr = rows[hi]
ci, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
items = []
fname = ""
for r in rows[hi + 1:]:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
    if len(r) > ci and r[0].isdigit() and r[ci].isdigit():
        items.append((int(r[ci]), int(r[si]) if r[si].isdigit() else 0, int(r[ti]) if r[ti].isdigit() else 0, r[0], r[1]))
tot = sum(i[0] for i in items) or 1
tots = sum(i[1] for i in items) or 1
print("total warp instructions %d, samples %d" % (tot, tots))
print("--- by instructions")
for c, s, t, ln, text in sorted(items, reverse=True)[:top]:
    print("%10d %5.1f%%  samp %5.1f%%  lanes %4.1f  L%-4s %s" % (c, 100 * c / tot, 100 * s / tots, t / max(c, 1), ln, text.strip()[:100]))
print("--- by stall samples")
for c, s, t, ln, text in sorted(items, key=lambda x: -x[1])[:top]:
    print("%10d %5.1f%%  samp %5.1f%%  lanes %4.1f  L%-4s %s" % (c, 100 * c / tot, 100 * s / tots, t / max(c, 1), ln, text.strip()[:100]))
