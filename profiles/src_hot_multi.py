#!/usr/bin/env python
"""Hottest CUDA source lines per kernel of a multi-kernel `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`
dump (one group of per-file sections per profiled launch, in launch order).
usage: src_hot_multi.py dump.csv name1:first_section name2:first_section ... [top=N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Line No":
        cur = {"hdr": r, "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
groups = [a.split(":") for a in sys.argv[2:] if ":" in a]
top = next((int(a[4:]) for a in sys.argv[2:] if a.startswith("top=")), 22)
bounds = [int(g[1]) for g in groups] + [len(secs)]
for (name, _), lo, hi in zip(groups, bounds, bounds[1:]):
    items = []
    for s in secs[lo:hi]:
        hdr = s["hdr"]
        if "Instructions Executed" not in hdr:
            continue
        ci, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        for r in s["rows"]:
            if len(r) > ci and r[0].isdigit() and r[ci].isdigit():
                items.append((int(r[ci]), int(r[si]) if r[si].isdigit() else 0, int(r[ti]) if r[ti].isdigit() else 0, r[0], r[1]))
    tot = sum(i[0] for i in items) or 1
    tots = sum(i[1] for i in items) or 1
    print("== %s hottest lines" % name)
    print("total warp instructions %d, samples %d" % (tot, tots))
    print("--- by instructions")
    for c, s, t, ln, text in sorted(items, reverse=True)[:top]:
        print("%10d %5.1f%%  samp %5.1f%%  lanes %4.1f  L%-4s %s" % (c, 100 * c / tot, 100 * s / tots, t / max(c, 1), ln, text.strip()[:100]))
    print("--- by stall samples")
    for c, s, t, ln, text in sorted(items, key=lambda x: -x[1])[:top]:
        print("%10d %5.1f%%  samp %5.1f%%  lanes %4.1f  L%-4s %s" % (c, 100 * c / tot, 100 * s / tots, t / max(c, 1), ln, text.strip()[:100]))
    print()
