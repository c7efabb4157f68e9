"""config C5: parabolic_solve on fem2d level L with h = 0.02 (the 50-step run of docs/src/guide.md:360-367), first
`steps` time steps: symbolic phase once, then per step assembly seconds vs solve-seam seconds"""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mgb_b200
from mgb_b200 import solver
L = int(sys.argv[1]); steps = int(sys.argv[2])
geom = mgb_b200.fem2d(L)
t0 = time.time()
sol = solver.parabolic_solve(geom, h=0.02, t1=1.0, p=1.0, max_steps=steps)
out = dict(config=f"fem2d L={L} parabolic, h=0.02 (50 steps to t1=1), first {steps} steps", n=int(geom.x.shape[0]), wall_s=time.time() - t0,
           plan_s_total=sol.stats["plan_s_total"], steps=sol.stats["steps"],
           u_range=[[float(u[:, 0].min()), float(u[:, 0].max())] for u in sol.u])
print(json.dumps(out))
