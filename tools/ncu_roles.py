#!/usr/bin/env python
"""Stall samples of a conv3x3_chain_kernel ncu capture (--page source --csv --print-source cuda,sass) by warp role and
by call site (inlined waits are told apart by the chain.cuh line that precedes them in address order)."""
import csv, sys
path = sys.argv[1]
lo_prod, lo_iss, lo_epi = (int(v) for v in sys.argv[2:5])   # first source line of producer / issuer / epilogue code
rows = list(csv.reader(open(path)))
cur = hdr = line = None
recs = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 3 and r[0] == 'Line No': hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr and len(r) == len(hdr):
        if r[2] == '-': line = (cur, int(r[0])); continue
        try: addr = int(r[2], 16)
        except ValueError: continue
        try: smp = float(r[ix['# Samples']])
        except ValueError: smp = 0
        recs.append((addr, line, smp, r[3]))
recs.sort()
role_of = lambda ln: 'prologue' if ln < lo_prod else ('producer' if ln < lo_iss else ('issuer' if ln < lo_epi else 'epilogue'))
last = None
agg, roles = {}, {}
for addr, line, smp, sass in recs:
    if line[0] == 'conv3x3_chain.cuh' and line[1] >= 200: last = line[1]
    role = role_of(last) if last else 'prologue'
    roles[role] = roles.get(role, 0) + smp
    key = (role, last, line)
    agg[key] = agg.get(key, 0) + smp
tot = sum(roles.values())
print({k: round(100 * v / tot, 1) for k, v in roles.items()}, 'total', tot)
for role in ('issuer', 'producer'):
    print('==', role, roles.get(role))
    for (rl, ctx, line), s in sorted(agg.items(), key=lambda kv: -kv[1]):
        if rl == role and s > 0.01 * roles[role]:
            print('   ctx %4d  %-18s %4d %6.0f  %4.1f%%' % (ctx, line[0][:18], line[1], s, 100 * s / roles[role]))
