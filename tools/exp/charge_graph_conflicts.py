import sys, itertools, collections
sys.path.insert(0,'/root/repo/tests')
from dem_util import read_dem
import re
name=sys.argv[1] if len(sys.argv)>1 else 'midout488_color_code_d9_r33_p1000'
text=read_dem(name)
# flatten with repeat/shift
lines=text.split('\n')
def parse(lines,pos):
    out=[]
    while pos<len(lines):
        l=lines[pos].strip(); pos+=1
        if not l: continue
        if l=='}': return out,pos
        m=re.match(r'repeat (\d+) \{',l)
        if m:
            body,pos=parse(lines,pos); out.append(('repeat',int(m.group(1)),body)); continue
        out.append(l)
    return out,pos
tree,_=parse(lines,0)
errs=[]; dets={}
off=[0]; coff=[0,0,0,0]
def run(block):
    for ins in block:
        if isinstance(ins,tuple):
            for _ in range(ins[1]): run(ins[2])
        elif ins.startswith('error'):
            p=float(ins[ins.index('(')+1:ins.index(')')]); t=ins[ins.index(')')+1:].split()
            d=[int(x[1:])+off[0] for x in t if x[0]=='D']; o=[int(x[1:]) for x in t if x[0]=='L']
            errs.append((p,tuple(sorted(d)),sum(1<<k for k in o)))
        elif ins.startswith('detector'):
            c=[float(x) for x in ins[ins.index('(')+1:ins.index(')')].split(',')]
            t=ins[ins.index(')')+1:].split()
            for x in t:
                dets[int(x[1:])+off[0]]=[c[i]+(coff[i] if i<len(coff) else 0) for i in range(len(c))]
        elif ins.startswith('shift_detectors'):
            if '(' in ins:
                c=[float(x) for x in ins[ins.index('(')+1:ins.index(')')].split(',')]
                for i,v in enumerate(c): coff[i]+=v
                t=ins[ins.index(')')+1:].split()
            else: t=ins.split()[1:]
            if t: off[0]+=int(t[0])
run(tree)
color={d:int(c[3])%3+1 for d,c in dets.items()}; basis={d:int(c[3])//3 for d,c in dets.items()}
B=2**32-1
atoms={}
for p,d,o in errs:
    if len(d)==1: atoms[(d[0],B,B)]=o
    elif len(d)==2 and basis[d[0]]==basis[d[1]]: atoms[(d[0],d[1],B)]=o
    elif len(d)==3 and basis[d[0]]==basis[d[1]]==basis[d[2]] and color[d[0]]^color[d[1]]^color[d[2]]==0: atoms[(d[0],d[1],d[2])]=o
print('atoms',len(atoms))
gen=collections.defaultdict(list)
for k in sorted(atoms):
    if k[2]==B: gen[(k[0],k[1])].append(('direct',k,atoms[k]))
n2e=collections.defaultdict(list)
for k in sorted(atoms):
    for n in k:
        if n!=B: n2e[n].append(k)
for n in sorted(n2e):
    L=n2e[n]
    for i in range(len(L)):
        for j in range(i+1,len(L)):
            e1,e2=L[i],L[j]
            w1=sum(x!=B for x in e1); w2=sum(x!=B for x in e2)
            if w1<3 and w2<3: continue
            x=[]
            for v in sorted(e1+e2):
                if x and x[-1]==v: x.pop()
                else: x.append(v)
            if len(x)==1: a,b=x[0],B
            elif len(x)==2 or (len(x)==3 and x[-1]==B): a,b=x[0],x[1]
            else: continue
            gen[(a,b)].append((n,e1,e2,atoms[e1]^atoms[e2]))
conf={k:v for k,v in gen.items() if len(set(g[-1] for g in v))>1}
print('edges',len(gen),'conflicting',len(conf))
for k in sorted(conf)[:12]:
    print(k,[ (dets[x][:4] if x!=B else 'B') for x in k])
    for g in conf[k]: print('    ',g)
import pickle; pickle.dump((dets,atoms,gen,conf),open('/tmp/exp/cg_%s.pkl'%name[:12],'wb'))
