"""Chronological view of a tools/trace_retrieval.py dump:  python tools/trace_events.py FILE SECTION LO HI"""
import re
import sys
fn, sec_want, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
sec = role = None
data = {}
for ln in open(fn).read().split('\n'):
    m = re.match(r'==== (\w+):', ln)
    if m: sec = m.group(1); continue
    m = re.match(r'-- (\w+)', ln)
    if m: role = m.group(1); continue
    m = re.match(r'\s+tile\s+(\d+):(.*)', ln)
    if m and sec:
        data.setdefault(sec, {}).setdefault(role, {})[int(m.group(1))] = [int(x) if x != '-' else None for x in m.group(2).split()]
d = data[sec_want]; ev = []
for t, v in d['MMA'].items():
    if sec_want == 'fwd':
        ev.append((v[0], f'S({t}) issuer: wait tile')); ev.append((v[1], f'S({t}) issuer: tile landed (waited {v[1]-v[0]})'))
        ev.append((v[2], f'S({t}) issue start (S buffer wait {v[2]-v[1]})')); ev.append((v[3], f'S({t}) issue end ({v[3]-v[2]})'))
    else:
        ev.append((v[0], f'S({t}) issue start'))
        ev.append((v[2], f'   dX({t}) issue start (waited {v[2]-v[1]})'))
        ev.append((v[3], f'   dX({t}) issue end ({v[3]-v[2]})'))
for g in ['WG0', 'WG1']:
    for t, v in d[g].items():
        ev.append((v[1], f'      {g} S({t}) ready (waited {v[1]-v[0]})'))
        ev.append((v[2], f'      {g} {"loaded" if sec_want == "fwd" else "exps"}({t}) done ({v[2]-v[1]})'))
        ev.append((v[3], f'      {g} {"tile" if sec_want == "fwd" else "dS"}({t}) done ({v[3]-v[2]})'))
for t, v in d['TMA'].items():
    ev.append((v[1], f'            TMA({t}) issued (waited {v[1]-v[0]})'))
for c, e in sorted(ev):
    if lo < c < hi: print(c, e)
