#!/usr/bin/env python
"""fp32 emulation of the two sub-pose schemes (plane rotation vs Reinsch cosine
recurrence) for the ground test, against the fp64 truth.  Used to choose the
kernel formulation; see DESIGN.md."""
import sys
import numpy as np
f32 = np.float32
def fma(a,b,c): return (np.asarray(a,np.float64)*np.asarray(b,np.float64)+np.asarray(c,np.float64)).astype(f32)
def mul(a,b): return (np.asarray(a,f32)*np.asarray(b,f32)).astype(f32)
def add(a,b): return (np.asarray(a,f32)+np.asarray(b,f32)).astype(f32)

rng = np.random.RandomState(0)
N = 400000
wide = len(sys.argv) > 1
g = rng.randint(-180,180,size=(N,4)).astype(np.float64)
a = rng.randint(-180,180,size=(N,4)).astype(np.float64)
if wide:
    a = rng.uniform(-720,720,size=(N,4)).astype(f32).astype(np.float64); g = rng.uniform(-720,720,size=(N,4)).astype(f32).astype(np.float64)
k = np.arange(25)[:,None,None]
th = np.radians(g[None] + k*(a-g)[None]/24)
c1,s1,c2,c3,s3 = np.cos(th[...,1]),np.sin(th[...,1]),np.cos(th[...,2]),np.cos(th[...,3]),np.sin(th[...,3])
ze = 4.3+24.3*c1; zt = ze+27*(c3*c1-s3*s1*c2)
ztrue = np.minimum(ze,zt)
neg_true = (ztrue<0).any(0); margin = np.abs(np.concatenate([ze,zt])).min(0)

def sc32(deg):
    r = np.radians(np.asarray(deg,np.float64)); return np.sin(r).astype(f32), np.cos(r).astype(f32)
A32, G32 = a.astype(f32), g.astype(f32)
d = mul((A32-G32).astype(f32), f32(1/24))

s,c = sc32(A32); sd,cd = sc32(d)
C1,S1,C2,S2,C3,S3 = c[:,1].copy(),s[:,1].copy(),c[:,2].copy(),s[:,2].copy(),c[:,3].copy(),s[:,3].copy()
zr = np.empty((25,N),f32)
def zpair(C1,S1,C2,C3,S3):
    ze = fma(f32(24.3),C1,f32(4.3))
    zt = fma(f32(27),fma(C3,C1,-mul(mul(S3,S1),C2)),ze)
    return np.minimum(ze,zt)
zr[24]=zpair(C1,S1,C2,C3,S3)
def rot(c,s,cd,sd):
    return fma(c,cd,mul(s,sd)), fma(s,cd,-mul(c,sd))
for kk in range(23,-1,-1):
    C1,S1=rot(C1,S1,cd[:,1],sd[:,1]); C2,S2=rot(C2,S2,cd[:,2],sd[:,2]); C3,S3=rot(C3,S3,cd[:,3],sd[:,3])
    zr[kk]=zpair(C1,S1,C2,C3,S3)
err_r = np.abs(zr.astype(np.float64)-ztrue).max(0)

def reinsch_init(amp, cth, sth, half_deg):
    sh,ch = sc32(half_deg)
    alpha = mul(mul(f32(4),sh),sh)
    sind = mul(mul(f32(2),sh),ch)
    x0 = mul(amp,cth)
    d0 = mul(amp, fma(sth,sind,-mul(mul(f32(0.5),alpha),cth)))
    return x0,d0,alpha
c1f,s1f,c2f,s2f,c3f,s3f = c[:,1],s[:,1],c[:,2],s[:,2],c[:,3],s[:,3]
cA = fma(c1f,c3f,-mul(s1f,s3f)); sA = fma(s1f,c3f,mul(c1f,s3f))
cB = fma(c1f,c3f,mul(s1f,s3f));  sB = fma(s1f,c3f,-mul(c1f,s3f))
h = mul(d,f32(0.5))
x1,d1,a1 = reinsch_init(f32(24.3),c1f,s1f,h[:,1])
x2,d2,a2 = reinsch_init(f32(1.0),c2f,s2f,h[:,2])
xA,dA,aA = reinsch_init(f32(13.5),cA,sA,add(h[:,1],h[:,3]))
xB,dB,aB = reinsch_init(f32(13.5),cB,sB,add(h[:,1],-h[:,3]))
zq = np.empty((25,N),f32)
def zz(x1,x2,xA,xB):
    p=add(xA,xB); m=add(xA,-xB); w=add(x1,p); t=fma(x2,m,w)
    return add(np.minimum(x1,t),f32(4.3))
zq[24]=zz(x1,x2,xA,xB)
for kk in range(23,-1,-1):
    x1=add(x1,d1); d1=fma(-a1,x1,d1)
    x2=add(x2,d2); d2=fma(-a2,x2,d2)
    xA=add(xA,dA); dA=fma(-aA,xA,dA)
    xB=add(xB,dB); dB=fma(-aB,xB,dB)
    zq[kk]=zz(x1,x2,xA,xB)
err_q = np.abs(zq.astype(np.float64)-ztrue).max(0)
for name,z,err in (("rotation",zr,err_r),("reinsch",zq,err_q)):
    neg=(z<0).any(0)
    mism = neg!=neg_true
    print(f"{name}: max|dz| {err.max():.2e}  p99.9 {np.quantile(err,0.999):.2e}  mean {err.mean():.2e}  "
          f"flag mismatches {mism.sum()} / {N}  (max margin among mismatches {margin[mism].max() if mism.any() else 0:.2e})")

# ---- round 2: two sub-poses per packed iteration (lane .x = odd k, lane .y = even k, both stepping 2*delta) ----
sd1,cd1 = sc32(d[:,1]); sd2,cd2 = sc32(d[:,2]); sd3,cd3 = sc32(d[:,3])
cdA = fma(cd1,cd3,-mul(sd1,sd3)); sdA = fma(sd1,cd3,mul(cd1,sd3))
cdB = fma(cd1,cd3,mul(sd1,sd3));  sdB = fma(sd1,cd3,-mul(cd1,sd3))
def init2(amp, cth, sth, sd, cd):
    na = -mul(mul(f32(4),sd),sd)                       # -4 sin^2(delta): the step of both lanes is 2*delta
    ax, ay = mul(amp,cth), mul(amp,sth)
    x_m1 = fma(ax,cd,-mul(ay,sd))                      # amp cos(theta + delta)   (k = -1)
    x_0 = ax                                           # k = 0
    d_x = mul(mul(f32(2),ay),sd)                       # x(1) - x(-1) = 2 amp sin(theta) sin(delta)
    d_y = fma(mul(ay,mul(f32(2),sd)),cd, mul(mul(f32(0.5),na),ax))   # x(2) - x(0) = amp (sin th sin 2d - 2 sin^2 d cos th)
    return [x_m1,x_0],[d_x,d_y],na
X1,D1,n1 = init2(f32(24.3),c1f,s1f,sd1,cd1)
X2,D2,n2 = init2(f32(1.0),c2f,s2f,sd2,cd2)
XA,DA,nA = init2(f32(13.5),cA,sA,sdA,cdA)
XB,DB,nB = init2(f32(13.5),cB,sB,sdB,cdB)
zp = np.empty((25,N),f32); zp[24]=zz(mul(f32(24.3),c1f), c2f, mul(f32(13.5),cA), mul(f32(13.5),cB))
for it in range(12):
    for lane in (0,1):
        for X,D,n in ((X1,D1,n1),(X2,D2,n2),(XA,DA,nA),(XB,DB,nB)):
            X[lane]=add(X[lane],D[lane]); D[lane]=fma(n,X[lane],D[lane])
        k = 2*it+1+lane
        zp[24-k]=zz(X1[lane],X2[lane],XA[lane],XB[lane])
err_p = np.abs(zp.astype(np.float64)-ztrue).max(0)
neg=(zp<0).any(0); mism = neg!=neg_true
print(f"paired : max|dz| {err_p.max():.2e}  p99.9 {np.quantile(err_p,0.999):.2e}  mean {err_p.mean():.2e}  "
      f"flag mismatches {mism.sum()} / {N}  (max margin among mismatches {margin[mism].max() if mism.any() else 0:.2e})")
