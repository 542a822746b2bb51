#!/usr/bin/env python
"""Aggregate an ncu report's stall samples / executed instructions by device function of lm_kernel.cuh."""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
src = open('lsqfit_b200/csrc/lm_kernel.cuh').read().split('\n')
funcs = []
for n, l in enumerate(src, 1):
    m = re.search(r'^(?:__device__|__global__)[^(]*?\b(\w+)\s*\(', l)
    if m: funcs.append((n, m.group(1)))
def fn(line):
    name = '?'
    for n, f in funcs:
        if n <= line: name = f
    return name
cur = None; agg = collections.Counter(); aggi = collections.Counter(); tot = 0; toti = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] in ('Line No', 'Function Name') or r[0] == '': continue
    try: line = int(r[0]); s = int(r[6]); i = int(r[7])
    except ValueError: continue
    key = fn(line) if cur == 'lm_kernel.cuh' else cur
    agg[key] += s; aggi[key] += i; tot += s; toti += i
print('total warp instructions %.3e' % toti)
for k, v in agg.most_common(16): print('%-28s %5.1f%% samples %5.1f%% instr' % (k, 100 * v / tot, 100 * aggi[k] / toti))
