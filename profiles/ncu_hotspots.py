#!/usr/bin/env python
"""Per-instruction hot spots from a .ncu-rep captured with --import-source on:  ncu_hotspots.py rep kernel-regex [min%]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
k = 0
while k < len(rows):
    if rows[k] and rows[k][0].startswith('Kernel Name'):
        name = rows[k][1][:100]; hdr = rows[k + 1]; data = []
        k += 2
        while k < len(rows) and rows[k] and not rows[k][0].startswith('Kernel Name'):
            if len(rows[k]) == len(hdr): data.append(rows[k])
            k += 1
        iS, iI, iT, isrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('Source')
        tot = sum(int(r[iS]) for r in data) or 1; toti = sum(int(r[iI]) for r in data) or 1
        print('==', name, '| SASS instrs', len(data), 'samples', tot, 'warp-inst', toti)
        for lo in range(0, len(data), 32):
            seg = data[lo:lo + 32]; s = sum(int(r[iS]) for r in seg); i = sum(int(r[iI]) for r in seg)
            if s > tot * 0.02 or i > toti * 0.02:
                print(f'  [{lo:4d}..] samples {100 * s / tot:5.1f}%  inst {100 * i / toti:5.1f}%')
        for n, r in enumerate(data):
            s = int(r[iS])
            if s > tot * thr / 100:
                print(f'  {n:4d} {r[isrc].strip()[:58]:58s} samp {100 * s / tot:4.1f}% exec {r[iI]:>9s} thr/inst {int(r[iT]) / max(1, int(r[iI])):4.1f}')
        break
    k += 1
