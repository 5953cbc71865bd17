import numpy as np, sys
sys.path.insert(0,'/root/repo')
import oracle
from mcos_b200.synthetic import block_diagonal_truth

def make_obj(lam, q, h=0.25, pts=1000):
    lam = np.asarray(lam)
    N = lam.size
    c = 1.0/(N*h*np.sqrt(2*np.pi))
    def fg(v):
        lo = v*(1-(1./q)**.5)**2; hi = v*(1+(1./q)**.5)**2
        x = np.linspace(lo, hi, pts)
        pdf = q/(2*np.pi*v*x)*np.sqrt(np.maximum((hi-x)*(x-lo),0))
        dif = x[:,None]-lam[None,:]
        E = np.exp(-dif**2/(2*h*h))
        kde = c*E.sum(1)
        dkde = c*(E*(-dif/(h*h))).sum(1)*(x/v)
        dpdf = -pdf/v
        r = kde-pdf
        return float((r*r).sum()), float((2*r*(dkde-dpdf)).sum())
    return fg

def dcstep(stx,fx,dx,sty,fy,dy,stp,fp,dp,brackt,stpmin,stpmax):
    sgnd = dp*(dx/abs(dx))
    if fp > fx:
        theta = 3*(fx-fp)/(stp-stx)+dx+dp
        s = max(abs(theta),abs(dx),abs(dp))
        gamma = s*np.sqrt((theta/s)**2-(dx/s)*(dp/s))
        if stp < stx: gamma=-gamma
        p=(gamma-dx)+theta; q=((gamma-dx)+gamma)+dp; r=p/q
        stpc = stx+r*(stp-stx)
        stpq = stx+((dx/((fx-fp)/(stp-stx)+dx))/2)*(stp-stx)
        stpf = stpc if abs(stpc-stx)<abs(stpq-stx) else stpc+(stpq-stpc)/2
        brackt=True
    elif sgnd < 0:
        theta = 3*(fx-fp)/(stp-stx)+dx+dp
        s = max(abs(theta),abs(dx),abs(dp))
        gamma = s*np.sqrt((theta/s)**2-(dx/s)*(dp/s))
        if stp > stx: gamma=-gamma
        p=(gamma-dp)+theta; q=((gamma-dp)+gamma)+dx; r=p/q
        stpc = stp+r*(stx-stp); stpq = stp+(dp/(dp-dx))*(stx-stp)
        stpf = stpc if abs(stpc-stp)>abs(stpq-stp) else stpq
        brackt=True
    elif abs(dp) < abs(dx):
        theta = 3*(fx-fp)/(stp-stx)+dx+dp
        s = max(abs(theta),abs(dx),abs(dp))
        gamma = s*np.sqrt(max(0.,(theta/s)**2-(dx/s)*(dp/s)))
        if stp > stx: gamma=-gamma
        p=(gamma-dp)+theta; q=(gamma+(dx-dp))+gamma; r=p/q
        if r<0 and gamma!=0: stpc = stp+r*(stx-stp)
        elif stp>stx: stpc=stpmax
        else: stpc=stpmin
        stpq = stp+(dp/(dp-dx))*(stx-stp)
        if brackt:
            stpf = stpc if abs(stpc-stp)<abs(stpq-stp) else stpq
            if stp>stx: stpf=min(stp+.66*(sty-stp),stpf)
            else: stpf=max(stp+.66*(sty-stp),stpf)
        else:
            stpf = stpc if abs(stpc-stp)>abs(stpq-stp) else stpq
            stpf=min(stpmax,stpf); stpf=max(stpmin,stpf)
    else:
        if brackt:
            theta=3*(fp-fy)/(sty-stp)+dy+dp
            s=max(abs(theta),abs(dy),abs(dp))
            gamma=s*np.sqrt((theta/s)**2-(dy/s)*(dp/s))
            if stp>sty: gamma=-gamma
            p=(gamma-dp)+theta; q=((gamma-dp)+gamma)+dy; r=p/q
            stpc=stp+r*(sty-stp); stpf=stpc
        elif stp>stx: stpf=stpmax
        else: stpf=stpmin
    if fp>fx:
        sty,fy,dy=stp,fp,dp
    else:
        if sgnd<0: sty,fy,dy=stx,fx,dx
        stx,fx,dx=stp,fp,dp
    return stx,fx,dx,sty,fy,dy,stpf,brackt

def lbfgsb_1d(fg, x0=0.5, l=1e-5, u=1-1e-5, pgtol=1e-5, factr=1e7, maxiter=15000, verbose=False):
    eps=np.finfo(float).eps
    ftol,gtol,xtol=1e-3,0.9,0.1
    x=min(max(x0,l),u)
    f,g=fg(x); nfev=1
    def projg(x,g):
        return max(x-u,g) if g<0 else min(x-l,g)
    if abs(projg(x,g))<=pgtol: return x,True,0,nfev
    theta=1.0; it=0
    while True:
        # cauchy + subspace min in 1-D
        z=min(max(x-g/theta,l),u)
        d=z-x
        # line search
        if it==0: stpmx=1.0
        else:
            stpmx=1e10
            if d<0: stpmx=(l-x)/d
            elif d>0: stpmx=(u-x)/d
        stp=1.0
        xold,fold,gold=x,f,g
        gd=g*d; gdold=gd
        if gd>=0:
            # ascent direction: refresh memory & restart
            if theta==1.0: return x,False,it,nfev
            theta=1.0; continue
        # dcsrch start
        brackt=False; stage=1; finit=f; ginit=gd; gtest=ftol*ginit; width=stpmx-0.0; width1=2*width
        stx=0.;fx=finit;gx=ginit;sty=0.;fy=finit;gy=ginit;stmin=0.;stmax=stp+4.0*stp
        nls=0; ok=True
        while True:
            x = z if stp==1.0 else xold+stp*d
            f,g=fg(x); nfev+=1; nls+=1
            gd=g*d
            ftest=finit+stp*gtest
            if stage==1 and f<=ftest and gd>=0: stage=2
            task=None
            if brackt and (stp<=stmin or stp>=stmax): task='W'
            if brackt and stmax-stmin<=xtol*stmax: task='W'
            if stp==stpmx and f<=ftest and gd<=gtest: task='W'
            if stp==0.0 and (f>ftest or gd>=gtest): task='W'
            if f<=ftest and abs(gd)<=gtol*(-ginit): task='C'
            if task is not None: break
            if nls>=20: ok=False; break
            if stage==1 and f<=fx and f>ftest:
                fm=f-stp*gtest; fxm=fx-stx*gtest; fym=fy-sty*gtest; gm=gd-gtest; gxm=gx-gtest; gym=gy-gtest
                stx,fxm,gxm,sty,fym,gym,stp,brackt=dcstep(stx,fxm,gxm,sty,fym,gym,stp,fm,gm,brackt,stmin,stmax)
                fx=fxm+stx*gtest; fy=fym+sty*gtest; gx=gxm+gtest; gy=gym+gtest
            else:
                stx,fx,gx,sty,fy,gy,stp,brackt=dcstep(stx,fx,gx,sty,fy,gy,stp,f,gd,brackt,stmin,stmax)
            if brackt:
                if abs(sty-stx)>=.66*width1: stp=stx+.5*(sty-stx)
                width1=width; width=abs(sty-stx)
            if brackt: stmin=min(stx,sty); stmax=max(stx,sty)
            else: stmin=stp+1.1*(stp-stx); stmax=stp+4.0*(stp-stx)
            stp=max(stp,0.0); stp=min(stp,stpmx)
            if (brackt and (stp<=stmin or stp>=stmax)) or (brackt and stmax-stmin<=xtol*stmax): stp=stx
        if not ok or task=='W' and False:
            pass
        if not ok:
            # line search failed: restore and restart from scratch once
            x,f,g=xold,fold,gold
            if theta==1.0: return x,False,it,nfev
            theta=1.0; continue
        it+=1
        if verbose: print(it,x,f,g,stp,nfev)
        if abs(projg(x,g))<=pgtol: return x,True,it,nfev
        ddum=max(abs(fold),abs(f),1.0)
        if fold-f<=eps*factr*ddum: return x,True,it,nfev
        if it>=maxiter: return x,False,it,nfev
        r=g-gold
        if stp==1.0: dr=gd-gdold; ddum=-gdold
        else: dr=(gd-gdold)*stp; d=d*stp; ddum=-gdold*stp
        rr=r*r
        if dr<=eps*ddum:
            pass
        else:
            theta=rr/dr

if __name__=='__main__':
    from scipy.optimize import minimize
    import time
    for (N,T,seeds) in [(100,200,range(6)),(20,50,range(6)),(60,30,range(4)),(500,1000,range(3))]:
        if N==20:
            st=np.load('/root/repo/tests/golden/stock.npz'); mu,cov=st['mu'],st['cov']
        else:
            mu,cov=block_diagonal_truth(N, n_blocks=10 if N!=60 else 6)
        for sd in seeds:
            np.random.seed(sd)
            x=np.random.multivariate_normal(mu,cov,size=T)
            c=np.cov(x,rowvar=False) if T>N else oracle.moments_ledoit_wolf(x)[1]
            lam=np.sort(np.linalg.eigvalsh(oracle.cov_to_corr(c)))[::-1]
            q=T/N
            t0=time.time()
            out=minimize(lambda v: oracle.mp_fit_objective(v,lam,q),.5,bounds=((1e-5,1-1e-5),))
            t1=time.time()
            fg=make_obj(lam,q)
            xm,ok,it,nf=lbfgsb_1d(fg)
            lp=lambda v: v*(1+(1/q)**.5)**2
            nf_s=N-int(lam[::-1].searchsorted(lp(out.x[0] if out.success else 1.0)))
            nf_m=N-int(lam[::-1].searchsorted(lp(xm if ok else 1.0)))
            print(N,T,sd,'scipy',out.x[0],out.success,out.nit,out.nfev,'| mine',xm,ok,it,nf,'| nfacts',nf_s,nf_m, out.message)
