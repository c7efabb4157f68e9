"""fem3d whole solve (config C4 beyond the finest-level assembly): wall / assembly / solve-seam seconds"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mgb_b200
from mgb_b200 import solver
L = int(sys.argv[1])
geom = mgb_b200.fem3d(L)
t0 = time.time()
sol = solver.amgb(geom, p=1.0)
print(json.dumps(dict(mesh=f"fem3d L={L}", n=int(geom.x.shape[0]), wall_s=time.time() - t0, newton_its=int(sol.SOL_main["its"].sum()),
                      its_per_level=sol.SOL_main["its"].sum(axis=1).tolist(), stats=sol.stats)))
