#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export per basic block: share of issued warp instructions, average active
lanes, and the split of the stall samples by reason.   python profiles/ncu_blocks.py src.csv [min_share_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
kern = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'rows': []}; kern.append(cur); continue
    if r and r[0] == 'Address': cur['hdr'] = r; continue
    if cur is not None and r: cur['rows'].append(r)
for K in kern:
    h = {n: i for i, n in enumerate(K['hdr'])}
    stall_cols = [n for n in K['hdr'] if n.startswith('stall_') and 'Not Issued' not in n]
    data = []
    for r in K['rows']:
        data.append(dict(src=r[h['Source']].strip(), ie=int(r[h['Instructions Executed']]), te=int(r[h['Thread Instructions Executed']]),
                         smp=int(r[h['# Samples']]), st={c: int(r[h[c]] or 0) for c in stall_cols}))
    tot = sum(d['ie'] for d in data); smp = sum(d['smp'] for d in data); ttot = sum(d['te'] for d in data)
    print('=====', K['name'][:60])
    print('instructions %d, total warp-inst %.1fM thread-inst %.1fM avg lanes %.2f, samples %d' % (len(data), tot / 1e6, ttot / 1e6, ttot / max(tot, 1), smp))
    allst = {c: sum(d['st'][c] for d in data) for c in stall_cols}
    print('stall samples overall:', {k.replace('stall_', ''): round(100 * v / max(smp, 1), 1) for k, v in sorted(allst.items(), key=lambda x: -x[1]) if v > 0.01 * smp})
    segs = []; start = 0
    for i in range(1, len(data) + 1):
        if i == len(data) or abs(data[i]['ie'] - data[start]['ie']) > 0.03 * max(data[start]['ie'], 1):
            segs.append((start, i)); start = i
    for s, e in segs:
        ie = sum(d['ie'] for d in data[s:e]); te = sum(d['te'] for d in data[s:e]); sm = sum(d['smp'] for d in data[s:e])
        if 100.0 * ie / tot >= thr or 100.0 * sm / max(smp, 1) >= thr:
            ops = {}
            for d in data[s:e]:
                t = d['src'].split()
                op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
                ops[op] = ops.get(op, 0) + 1
            top = sorted(ops.items(), key=lambda x: -x[1])[:4]
            st = {}
            for d in data[s:e]:
                for c, v in d['st'].items(): st[c] = st.get(c, 0) + v
            tops = [(k.replace('stall_', ''), round(100 * v / max(sm, 1))) for k, v in sorted(st.items(), key=lambda x: -x[1])[:3] if v]
            print(f'[{s:4d},{e:4d}) n={e - s:3d} exec/inst={data[s]["ie"] / 1e6:7.3f}M share={100 * ie / tot:5.1f}% lanes={te / max(ie, 1):5.1f} samples={100 * sm / max(smp, 1):5.1f}% {tops} {top}')
