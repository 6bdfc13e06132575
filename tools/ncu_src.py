"""Summarise an `ncu --page source --csv` dump: instruction mix and hot regions. Usage: ncu_src.py file.csv n_rays"""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
nr=float(sys.argv[2]) if len(sys.argv)>2 else 1.0
hdr=rows[1]; which=int(sys.argv[3]) if len(sys.argv)>3 else 0
kern=[]; 
for r in rows:
    if r and r[0]=='Kernel Name': kern.append([r[1]]); continue
    if len(r)==len(hdr) and r[0].startswith('0x'): kern[-1].append(r)
print('kernels:', [k[0][:60] for k in kern]); data=kern[which][1:]
ia=hdr.index('Source'); ie=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples')
tot=sum(int(r[ie]) for r in data); print('SASS instrs', len(data), 'total warp instr', tot, 'per ray', tot/nr)
ops=collections.Counter(); samp=collections.Counter()
def opof(s):
    t=s.split()
    op=t[1] if t[0].startswith('@') else t[0]
    return op.split('.')[0]
for r in data:
    ops[opof(r[ia])]+=int(r[ie]); samp[opof(r[ia])]+=int(r[isamp])
ts=sum(samp.values())
for k,v in ops.most_common(28): print(f"{k:10s} {v:12d} {100*v/tot:5.1f}%  stall-samples {100*samp[k]/ts:5.1f}%")
print('--- hot regions (contiguous runs of similar execution count)')
seg=[];cur=None
for i,r in enumerate(data):
    e=int(r[ie]); b = 0 if e==0 else int(__import__('math').log2(e)*2)
    if cur is None or abs(cur[0]-b)>1: cur=[b,i,i,0,0]; seg.append(cur)
    cur[2]=i; cur[3]+=e; cur[4]+=int(r[isamp])
for s in seg:
    if s[3]>tot*0.02: print(f"instr {s[1]:5d}-{s[2]:5d} ({s[2]-s[1]+1:4d}) exec/instr {s[3]/(s[2]-s[1]+1):10.0f}  {100*s[3]/tot:5.1f}% of instr, {100*s[4]/ts:5.1f}% of samples")
