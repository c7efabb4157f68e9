"""1-GPU scratch: N virtual ranks (split mode, one stream) -> per-rank kernel timeline without NVLink."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MGB_DIST_DEBUG"] = "1"
import bench
from mgb_b200 import capi, dist as mdist
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
ctx = capi.Context(0, st.cuda_stream)
pr = bench.build_problem(8, 1.0); geom = pr["geom"]
n, m = geom.x.shape[0], pr["R"].shape[1]
flush_buf = torch.zeros(256 << 17, dtype=torch.float64, device=dev)
for N in (int(a) for a in sys.argv[1:] or ["8"]):
    rp, op = mdist.peer_partitions(n, m, geom.block, N)
    plans = [capi.DistPlan(ctx, pr["D"], pr["R"], geom.x, geom.w, pr["idx"], 1.0, r, N, rp, op) for r in range(N)]
    wins = [p.window()[0] for p in plans]
    for p in plans: p.attach_local(wins)
    s_d = torch.from_numpy(pr["s"]).to(dev)
    cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
    ins = [(cm(pr["Dz0"][rp[r]:rp[r+1]]), cm(pr["c"][rp[r]:rp[r+1]])) for r in range(N)]
    reps = 20
    for i in range(reps):
        for r, p in enumerate(plans):
            flush_buf.sum()
            p.begin(s_d, ins[r][0], ins[r][1], 1.0, 7)
        for p in plans: p.end(1.0, 7)
    torch.cuda.synchronize()
    for r, p in enumerate(plans):
        tl = p.debug_timeline().astype(np.int64)[1:reps + 1][5:]
        d0 = tl[:, 6]
        print(N, r, {k: round(float(np.mean(tl[:, c] - d0)) / 1e3, 2) for k, c in (("push_start", 0), ("stores_issued", 1), ("last_ticket", 2), ("flags_pub", 3))},
              "n_local_h", p.dinfo["n_local_h"])
    for p in plans: p.close()
