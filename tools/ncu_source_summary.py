"""Summarise an `ncu --page source --csv` export of one kernel: warp-stall samples and
executed instructions by opcode, stall reasons, and the hottest instructions."""
import csv, re, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = 0; by_op = Counter(); ex = Counter(); st = Counter(); data = []
def num(x):
    try: return int(float(x))
    except Exception: return 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    m = re.match(r"\s*(@!?U?P\d\s+)?([A-Z0-9_.]+)", r[idx['Source']])
    if not m: continue
    op = m.group(2).split('.')[0]
    ns = num(r[idx['# Samples']]); ie = num(r[idx['Instructions Executed']])
    tot += ns; by_op[op] += ns; ex[op] += ie
    for s in stalls: st[s] += num(r[idx[s]])
    data.append((r[idx['Address']], r[idx['Source']], ns, ie))
print(rows[0][1] if rows[0] else '')
print("total samples", tot, " warp instructions executed", sum(ex.values()))
print("samples by opcode:", [(k, round(100*v/tot, 1)) for k, v in by_op.most_common(14)])
print("executed by opcode (%):", [(k, round(100*v/sum(ex.values()), 1)) for k, v in ex.most_common(24)])
print("stall reasons (%):", [(k[6:], round(100*v/max(1, sum(st.values())), 1)) for k, v in st.most_common(12)])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
data.sort(key=lambda x: -x[2])
for d in data[:n]: print(d[0][-5:], "%5.2f%%" % (100*d[2]/tot), d[3], d[1][:100])
