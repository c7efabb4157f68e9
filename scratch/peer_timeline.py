"""torchrun scratch: per-kernel timeline of the peer exchange on N GPUs (events between begin/end)."""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mgb_b200 import capi, dist as mdist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
L = int(os.environ.get("BL", "8"))
pr = bench.build_problem(L, 1.0); geom = pr["geom"]
_st = torch.cuda.Stream(dev); torch.cuda.set_stream(_st)
ctx = capi.Context(lr, _st.cuda_stream)
plan = mdist.create_peer_plan(ctx, pr["D"], pr["R"], geom.x, geom.w, pr["idx"], 1.0, geom.block, rank, world)
d = plan.dinfo; r0, r1 = d["row0"], d["row1"]
s_d = torch.from_numpy(pr["s"]).to(dev)
cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
Dz0_d, c_d = cm(pr["Dz0"][r0:r1]), cm(pr["c"][r0:r1])
flush_buf = torch.zeros(256 << 17, dtype=torch.float64, device=dev)
res = {}
ep = 0
for mode in ("flush", "noflush"):
    for flags in (7, 1):
        n = 60
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
        dist.barrier(); torch.cuda.synchronize()
        for i in range(n):
            if mode == "flush": flush_buf.sum()
            evs[i][0].record(); plan.dist_assemble(s_d, Dz0_d, c_d, 1.0, flags)
            evs[i][1].record()
            evs[i][2].record()
        torch.cuda.synchronize()
        b = np.array([e[0].elapsed_time(e[1]) for e in evs[10:]]) * 1e3
        f = np.array([e[1].elapsed_time(e[2]) for e in evs[10:]]) * 1e3
        tl = plan.debug_timeline().astype(np.int64) if os.environ.get("MGB_DIST_DEBUG") else None
        if tl is not None:
            rows = np.array([tl[(ep + 1 + i) % 512] for i in range(10, n)])
            ep += n
            d0 = rows[:, 6]
            res[f"{mode}_flags{flags}_tl_us"] = {k: float(np.mean(rows[:, c] - d0)) / 1e3 for k, c in
                (("push_start", 0), ("stores_issued", 1), ("last_ticket", 2), ("flags_pub", 3), ("flags_seen", 4), ("finish_done", 5))}
        res[f"{mode}_flags{flags}"] = dict(begin_us=float(b.mean()), begin_min=float(b.min()), finish_us=float(f.mean()), finish_min=float(f.min()))
out = [None] * world
dist.all_gather_object(out, dict(rank=rank, dinfo=d, res=res))
if rank == 0:
    print(json.dumps(out, indent=1))
mdist.destroy_peer_plan(plan)
dist.destroy_process_group()
