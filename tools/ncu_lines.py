#!/usr/bin/env python
"""Aggregate an ncu report's warp-stall samples by CUDA source line.
Usage: python tools/ncu_lines.py gpurun_out/prof_X.ncu-rep [topN]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; agg = collections.OrderedDict(); tot = 0; toti = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] in ('Line No', 'Function Name'): continue
    if r[0] != '':
        try: line = int(r[0]); samples = int(r[6]); inst = int(r[7])
        except ValueError: continue
        a = agg.setdefault((cur, line), [0, 0, r[1][:100]])
        a[0] += samples; a[1] += inst; tot += samples; toti += inst
print('total samples', tot, 'total warp instructions', toti)
for (f, l), (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print('%5.1f%% %5.1f%%i  %s:%d  %s' % (100.0 * s / tot, 100.0 * i / max(toti, 1), f, l, src))
