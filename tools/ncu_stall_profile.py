"""Condense `ncu -i rep --page source --csv --kernel-name regex:K > file.csv` (warp-stall sampling per SASS instruction) into the stall mix of the
kernel's hot loop body (the instructions executed most often), the busiest opcodes and a timeline.  usage: ncu_stall_profile.py file.csv"""
import csv,sys
from collections import Counter
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
data=[]
for r in rows[2:]:
    if r and r[0]=='Kernel Name': break          # first kernel of the file only
    if len(r)==len(hdr) and r!=hdr: data.append(r)
ix={h:i for i,h in enumerate(hdr)}
stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
cnt=Counter(int(r[ix['Instructions Executed']]) for r in data if 'FFMA2' in r[ix['Source']])
modal=cnt.most_common(1)[0][0]
body=[(k,r) for k,r in enumerate(data) if int(r[ix['Instructions Executed']])==modal]
print('body instrs',len(body),'first',body[0][0],'last',body[-1][0], 'exec each', modal)
tot=sum(int(r[ix['# Samples']]) for k,r in body)
print('samples in body',tot,'of',sum(int(r[ix['# Samples']]) for r in data))
agg={h:sum(int(r[ix[h]]) for k,r in body) for h in stall_cols}
for h,v in sorted(agg.items(), key=lambda x:-x[1])[:8]: print(' ',h,v,'%.1f%%'%(100*v/tot))
c=Counter(); ce=Counter()
def opof(r):
    toks=r[ix['Source']].split(); return (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
for k,r in body:
    op=opof(r); c[op]+=int(r[ix['# Samples']]); ce[op]+=1
for op,v in c.most_common(14): print(' ',op,v,'%.1f%%'%(100*v/tot),'count',ce[op])
print('timeline (bucket of 40 instrs): samples')
for b in range(0,len(body),40):
    seg=body[b:b+40]
    ops=Counter(opof(r) for k,r in seg)
    st=Counter()
    for k,r in seg:
        for h in stall_cols: st[h]+=int(r[ix[h]])
    print(b, sum(int(r[ix['# Samples']]) for k,r in seg), dict(ops.most_common(4)), dict(st.most_common(3)))
# outside body hot
print('outside body, top instrs:')
out=[(int(r[ix['# Samples']]),k,r[ix['Source']].strip()[:60]) for k,r in enumerate(data) if int(r[ix['Instructions Executed']])!=modal]
for n,k,sx in sorted(out,reverse=True)[:12]: print(n,k,sx)
print('--- regions by execution count')
groups=Counter(); gs=Counter(); first={}
for k,r in enumerate(data):
    e=int(r[ix['Instructions Executed']]); groups[e]+=1; gs[e]+=int(r[ix['# Samples']]); first.setdefault(e,k)
for e,n in sorted(groups.items(), key=lambda x:-gs[x[0]])[:14]: print('exec',e,'instrs',n,'samples',gs[e],'first idx',first[e])
