"""Summarise an `ncu --page source --csv` dump: instruction mix, hot regions and the instructions that collect the
stall samples. Usage: ncu_src.py file.csv n_rays [kernel index] [n top instructions]"""
import csv, collections, math, sys
rows = list(csv.reader(open(sys.argv[1])))
nr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ntop = int(sys.argv[4]) if len(sys.argv) > 4 else 0
kern = []
for i, r in enumerate(rows):
    if r and r[0] == 'Kernel Name':
        kern.append(dict(name=r[1], hdr=rows[i + 1], data=[]))
    elif kern and len(r) == len(kern[-1]['hdr']) and r[0].startswith('0x'):
        kern[-1]['data'].append(r)
print('kernels:', [k['name'][:50] for k in kern])
k = kern[which]; hdr, data = k['hdr'], k['data']
ia = hdr.index('Source'); ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
tot = sum(int(r[ie]) for r in data); ts = sum(int(r[isamp]) for r in data)
print(k['name'][:60], 'SASS instrs', len(data), 'total warp instr', tot, 'per ray', tot / nr, 'samples', ts)
def opof(s):
    t = s.split()
    op = t[1] if t[0].startswith('@') else t[0]
    return op.split('.')[0]
ops = collections.Counter(); samp = collections.Counter()
for r in data:
    ops[opof(r[ia])] += int(r[ie]); samp[opof(r[ia])] += int(r[isamp])
for kk, v in ops.most_common(24): print(f"{kk:10s} {v:12d} {100*v/tot:5.1f}%  stall-samples {100*samp[kk]/ts:5.1f}%")
print('--- hot regions (contiguous runs of similar execution count)')
seg = []; cur = None
for i, r in enumerate(data):
    e = int(r[ie]); b = 0 if e == 0 else int(math.log2(e) * 2)
    if cur is None or abs(cur[0] - b) > 1: cur = [b, i, i, 0, 0]; seg.append(cur)
    cur[2] = i; cur[3] += e; cur[4] += int(r[isamp])
for s in seg:
    if s[3] > tot * 0.02: print(f"instr {s[1]:5d}-{s[2]:5d} ({s[2]-s[1]+1:4d}) exec/instr {s[3]/(s[2]-s[1]+1):10.0f}  {100*s[3]/tot:5.1f}% of instr, {100*s[4]/ts:5.1f}% of samples")
if ntop:
    print('--- instructions with the most stall samples')
    reasons = [c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
    top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:ntop]
    for i in sorted(top):
        r = data[i]
        rs = sorted(((int(r[hdr.index(c)] or 0), c) for c in reasons), reverse=True)[:2]
        print(f"{i:5d} {100*int(r[isamp])/ts:5.1f}% exec {int(r[ie]):9d}  {r[ia].strip()[:70]:70s} {rs[0][1]}:{rs[0][0]} {rs[1][1]}:{rs[1][0]}")
