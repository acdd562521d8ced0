#!/usr/bin/env python
"""Stall samples / instructions / shared-memory wavefronts of an ncu SASS page (csv of --page source --print-source sass),
cut into the phases of the kernel at its barriers, branches and tensor-copy issues: tools/ncu_phases.py sass.csv"""
import csv,sys,re
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; ix={k:i for i,k in enumerate(hdr)}
segs=[]; cur=dict(n=0,samp=0,inst=0,wf=0,wfi=0,start=None,ops={},stalls={})
tot=0
def flush(mark):
    global cur
    if cur['n']: cur['mark']=mark; segs.append(cur)
    cur=dict(n=0,samp=0,inst=0,wf=0,wfi=0,start=None,ops={},stalls={})
stk=[k for k in hdr if k.startswith('stall_') and 'Not' not in k]
for r in rows[2:]:
    if len(r)<len(hdr)-2: continue
    src=r[ix['Source']].strip(); op=src.split()[0] if not src.startswith('@') else src.split()[1]
    s=int(r[ix['# Samples']]); e=int(r[ix['Instructions Executed']])
    wf=int(r[ix['L1 Wavefronts Shared']] or 0); wfi=int(r[ix['L1 Wavefronts Shared Ideal']] or 0)
    if cur['start'] is None: cur['start']=r[0][-5:]
    cur['n']+=1; cur['samp']+=s; cur['inst']+=e; cur['wf']+=wf; cur['wfi']+=wfi
    o=op.split('.')[0]; cur['ops'][o]=cur['ops'].get(o,0)+e
    for k in stk:
        v=int(r[ix[k]] or 0)
        if v: cur['stalls'][k[6:]]=cur['stalls'].get(k[6:],0)+v
    tot+=s
    if re.match(r'(BAR|SYNCS|WARPSYNC|BRA|EXIT|USETMAXREG|UTMALDG|UBLKCP)',op): flush(src[:60])
flush('end')
print('total samples',tot)
for g in segs:
    if g['samp']<tot*0.004 and g['inst']<1e5: continue
    ops=sorted(g['ops'].items(),key=lambda kv:-kv[1])[:6]
    st=sorted(g['stalls'].items(),key=lambda kv:-kv[1])[:4]
    print(f"{g['start']} n={g['n']:4d} samp={100*g['samp']/tot:5.1f}% inst={g['inst']/1e6:7.2f}M wf={g['wf']/1e6:6.2f}M ideal={g['wfi']/1e6:6.2f}M | "+' '.join(f'{k}:{v/1e6:.1f}' for k,v in ops)+' | '+' '.join(f'{k}={v}' for k,v in st)+' | '+g['mark'])
