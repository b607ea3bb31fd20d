import sys
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__ if '__file__' in dir() else 'tools/x'))); sys.path.insert(0,os.path.join(R,'tests')); sys.path.insert(0,R)
import streamdump, numpy as np, refharness
G=dict(streamdump.load_golden())
key=sys.argv[1] if len(sys.argv)>1 else '8-bit/data/00000658.ivf#0'
s=G[key]
ref=refharness.load()
cur=refharness.RefFrame(ref,s,1); cur.load_filter_meta(); cur.set_planes(s.pre); cur.filter(2)
exp=streamdump.visible(s,cur.get_planes()); cur.close()
bd=s.bpc; bdmin8=bd-8; bdmax=(1<<bd)-1
m=s.masks; lut=s.lut
w4=(s.w+3)//4; h4=(s.h+3)//4
bw=(s.w+7)//8*2; sb128w=(bw+31)//32
lv=s.levels.reshape(-1,4); b4=(bw+31)&~31
def idx(dir,x4,y4):
    M=m[(y4>>5)*sb128w+(x4>>5)]
    xi=x4&31; yi=y4&31
    a,b=(xi,yi) if dir==0 else (yi,xi)
    fy=M['filter_y'][dir][a]
    for k in (2,1,0):
        if (int(fy[k][b>>4])>>(b&15))&1: return k
    return -1
def clip(v,lo,hi): return max(lo,min(hi,v))
def smooth(x,N,N2,LOG2):
    LO=8-(N+1); HI=8+N
    cl=lambda k: x[clip(k,LO,HI)]
    out=[]
    for i in range(-N,N):
        c=8+i
        win=sum(cl(c+j) for j in range(-N,N+1))
        centre=x[c]+((cl(c-1)+cl(c+1)) if N2 else 0)
        out.append((win+centre+(1<<(LOG2-1)))>>LOG2)
    for i,v in enumerate(out): x[8-N+i]=v
def edge(wd,x,E,I,H):
    p0,q0,p1,q1=x[7],x[8],x[6],x[9]
    step01=max(abs(p1-p0),abs(q1-q0)); step=step01
    if wd>4: step=max(step,abs(x[5]-p1),abs(x[10]-q1))
    if wd>6: step=max(step,abs(x[4]-x[5]),abs(x[11]-x[10]))
    if step>I or abs(p0-q0)*2+(abs(p1-q1)>>1)>E: return 0
    if wd>=6:
        F=1<<bdmin8
        dev=max(step01,abs(x[5]-p0),abs(x[10]-q0))
        if wd>=8: dev=max(dev,abs(x[4]-p0),abs(x[11]-q0))
        if dev<=F:
            if wd==16:
                far=max(abs(x[k]-p0) for k in (1,2,3)); far=max(far,max(abs(x[k]-q0) for k in (12,13,14)))
                if far<=F: smooth(x,6,1,4); return 6
            if wd>=8: smooth(x,3,0,3); return 3
            smooth(x,2,1,3); return 2
    lo=-(128<<bdmin8); hi=(128<<bdmin8)-1
    hev=step01>H
    f=clip(p1-q1,lo,hi) if hev else 0
    f=clip(f+3*(q0-p0),lo,hi)
    f1=min(f+4,hi)>>3; f2=min(f+3,hi)>>3
    x[8]=clip(q0-f1,0,bdmax); x[7]=clip(p0+f2,0,bdmax)
    if hev: return 1
    f3=(f1+1)>>1
    x[9]=clip(q1-f3,0,bdmax); x[6]=clip(p1+f3,0,bdmax)
    return 2
pic=s.pre[0].astype(np.int64).copy()
for d in (0,1):
    for y4 in range(h4):
        for x4 in range(w4):
            if (x4 if d==0 else y4)==0: continue
            i=idx(d,x4,y4)
            if i<0: continue
            L=int(lv[y4*b4+x4][d])
            if not L: L=int(lv[y4*b4+x4-(1 if d==0 else b4)][d])
            if not L: continue
            H=(L>>4)<<bdmin8; E=int(lut.e[L])<<bdmin8; I=int(lut.i[L])<<bdmin8
            wd=4<<i
            for k in range(4):
                if d==0:
                    yy=y4*4+k; x0=x4*4
                    x=[int(pic[yy,x0-8+j]) if 0<=x0-8+j<pic.shape[1] else 0 for j in range(16)]
                    n=edge(wd,x,E,I,H)
                    for j in range(8-n,8+n): pic[yy,x0-8+j]=x[j]
                else:
                    xx=x4*4+k; y0=y4*4
                    x=[int(pic[y0-8+j,xx]) if 0<=y0-8+j<pic.shape[0] else 0 for j in range(16)]
                    n=edge(wd,x,E,I,H)
                    for j in range(8-n,8+n): pic[y0-8+j,xx]=x[j]
vis=pic[:s.h,:s.w]
bad=np.argwhere(vis!=exp[0])
print(key,'python port vs oracle: ',len(bad),'px differ', bad[:10])
