import numpy as np, scipy.linalg as sl
EPS=2.220446049250313e-16
def pwk_trace(d,e):
    n=len(d); d=d.copy(); e=e.copy()
    tn=max(np.abs(d).max(),np.abs(e).max())
    eps2=EPS*EPS; abstol2=0.25*eps2*tn*tn
    e[:n-1]=e[:n-1]**2; e[n-1]=0
    trace=[]  # list of (l, trip)
    m=-1
    for l in range(n):
        it=0
        while True:
            if m<l:
                m=l
                while m<n-1:
                    em=e[m]
                    if em<=abstol2 or em<=eps2*abs(d[m]*d[m+1]): break
                    m+=1
            else:
                el=e[l]
                if el<=abstol2 or el<=eps2*abs(d[l]*d[l+1]): m=l
            if m==l: break
            it+=1
            rte=np.sqrt(e[l]); p=d[l]
            sigma=(d[l+1]-p)*0.5/rte
            r0=np.sqrt(sigma*sigma+1)
            sigma=p-rte/(sigma+np.copysign(r0,sigma))
            c=1.;sn=0.;gamma=d[m]-sigma;p=gamma*gamma
            msplit=m;dnext=0.;enew=0.
            trace.append((l,m-l))
            for i in range(m-1,l-1,-1):
                bb=e[i]; r=p+bb
                if i!=m-1:
                    enew=sn*r; e[i+1]=enew
                oldc=c
                c=p/r; sn=bb/r
                oldgam=gamma; alpha=d[i]
                gamma=c*(alpha-sigma)-sn*oldgam
                dn=oldgam+(alpha-gamma)
                d[i+1]=dn
                if i!=m-1 and (enew<=abstol2 or enew<=eps2*abs(dn*dnext)): msplit=i+1
                dnext=dn
                p=gamma*gamma*(r/p) if c!=0 else oldc*bb
            e[l]=sn*p; d[l]=sigma+gamma; m=msplit
    return np.sort(d),trace
rng=np.random.default_rng(1)
N=64
lanes=[]
for z in range(32):
    mloc=int(rng.integers(150,250))
    A=rng.normal(size=(mloc,N))*0.5/np.sqrt(N-1); A-=A.mean(1,keepdims=True)
    dist=np.sqrt(rng.uniform(0,64,size=mloc)); w=np.exp(-(dist/4)**2)
    rm=0.05*(1+0.5*rng.uniform(size=mloc))
    G=(A*( (w**2/rm**2)[:,None])).T@A
    H,Q=sl.hessenberg(G,calc_q=True)
    d=np.diag(H).copy(); e=np.append(np.diag(H,-1),0.)
    ev,tr=pwk_trace(d,e)
    assert np.allclose(ev,np.linalg.eigvalsh(G),atol=1e-9*abs(ev).max())
    lanes.append(tr)
rot=[sum(t for _,t in tr) for tr in lanes]; sw=[len(tr) for tr in lanes]
print("per-lane rotations mean",np.mean(rot),"sweeps",np.mean(sw))
# current structure
cur_rot=0;cur_sw=0
for l in range(N):
    per=[[t for ll,t in tr if ll==l] for tr in lanes]
    K=max(len(p) for p in per)
    for k in range(K):
        cur_rot+=max((p[k] for p in per if len(p)>k)); cur_sw+=1
print("current: warp rot trips",cur_rot,"sweep setups",cur_sw)
S=max(sw); fl=0
for s in range(S):
    fl+=max(tr[s][1] for tr in lanes if len(tr)>s)
print("flat: warp rot trips",fl,"sweeps",S)
