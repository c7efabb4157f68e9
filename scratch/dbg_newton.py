import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,ROOT+'/oracle')
import numpy as np, scipy.sparse as sp, torch, math
import mgb_b200, mgb_oracle as O
from mgb_b200 import solver, capi, amg
geom = mgb_b200.fem2d(3); p=1.5
# oracle trace
trace_o=[]
orig_newton = O.newton
def newton_trace(F0,F1,F2,x,maxit=50,alpha=0.1,beta=0.25,solve_fn=O.solve):
    y=F0(x); g=F1(x); k=0; rec=[]
    while k<maxit:
        H=F2(x); n=solve_fn(H,g); inc=float(g@n)
        rec.append(("inc",inc,"y",y))
        if not math.isfinite(inc) or inc <= O.NEWTON_NOISE*O.EPS*max(1.0,abs(y)): break
        k+=1; s=1.0; ok=False
        while s>1e-12:
            xn=x-s*n; yn=F0(xn)
            if math.isfinite(yn) and yn<=y-alpha*s*inc: ok=True; break
            s*=beta
        rec.append(("step",s,yn))
        if not ok: k-=1; break
        x,y=xn,yn; g=F1(x)
    trace_o.append(rec)
    return dict(x=x,y=y,k=k,converged=True)
O.newton=newton_trace
sol_o=O.amgb(geom,p=p)
for r in trace_o[:3]: print("ORACLE", r)
# gpu trace
orig=solver.newton_device
recs=[]
import time
def nd(prob,J,z,c,t,maxit,alpha=0.1,beta=0.25,solve_fn=solver.solve):
    lv=prob.level(J); plan=lv.plan; Dz0=prob.apply_D(z); lv.s.zero_()
    F0,FG,FH=1,2,4
    plan.assemble(lv.s,Dz0,c,t,7,lv.scal,lv.grad,lv.hval); sc=lv.scal.cpu(); y=float(sc[0]); k=0; rec=[]
    while k<maxit:
        H=sp.csr_matrix((lv.hval.cpu().numpy()[:plan.nnzH],lv.colidx,lv.rowptr),shape=(plan.m,plan.m)); g=lv.grad.cpu().numpy()
        n=solve_fn(H,g); inc=float(g@n); rec.append(("inc",inc,"y",y))
        if not math.isfinite(inc) or inc<=solver.NEWTON_NOISE*solver.EPS*max(1.0,abs(y)): break
        k+=1; lv.step.copy_(torch.from_numpy(n)); s=1.0; ok=False
        while s>1e-12:
            torch.add(lv.s,lv.step,alpha=-s,out=lv.trial); plan.assemble(lv.trial,Dz0,c,t,1,lv.scal); sc=lv.scal.cpu(); yn=float(sc[0])
            if sc[1]==1.0 and math.isfinite(yn) and yn<=y-alpha*s*inc: ok=True; break
            s*=beta
        rec.append(("step",s,yn))
        if not ok: k-=1; break
        lv.s.copy_(lv.trial); y=yn; plan.assemble(lv.s,Dz0,c,t,6,lv.scal,lv.grad,lv.hval)
    recs.append(rec)
    lv.R.mv(lv.s,z,beta=1.0,y0_dev=z)
    return dict(k=k,converged=True,y=y)
solver.newton_device=nd
sol_g=solver.amgb(geom,p=p)
for r in recs[:3]: print("GPU   ", r)
