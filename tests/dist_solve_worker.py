"""torchrun worker: BASELINE.json configs[0] - `fem2d_mpi_solve(Float64; L=3, p=1.0)` on 2 ranks (the reference's
quick example runs it with `mpiexec -n 2`, docs/src/guide.md:248) - through the API mirror with one GPU per rank:
the finest level plan is sharded (owner-computes), the solve seam sees the replicated system.
Checked on every rank against the CPU oracle: identical t-schedule and Newton iteration counts per level,
solution within 1e-9 relative (north_star bar).  Launched by tests/test_dist_peer_gpu.py with >= 2 GPUs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import mgb_b200  # noqa: E402
from mgb_b200 import api  # noqa: E402
import mgb_oracle as O  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
for name, kw, geom in (("fem2d", dict(L=3, p=1.0), mgb_b200.fem2d(3)), ("fem1d", dict(L=4, p=2.0), mgb_b200.fem1d(4)),
                       ("fem2d", dict(L=4, p=1.5), mgb_b200.fem2d(4)),
                       ("fem3d", dict(L=2, p=1.0), mgb_b200.fem3d(2))):   # Q3 elements: CSR path, assembled redundantly
    sol = getattr(api, name + "_mpi_solve")(**kw)          # backend: this job's ranks, one GPU each
    assert sol.stats["nranks"] == world
    native = api.mpi_to_native(sol)                         # collective gather of the row-partitioned solution
    ref = O.amgb(geom, p=kw["p"])
    assert np.array_equal(native.SOL_main["ts"], ref.SOL_main["ts"])
    assert np.array_equal(native.SOL_main["its"], ref.SOL_main["its"]), (native.SOL_main["its"], ref.SOL_main["its"])
    rel = np.linalg.norm(native.z - ref.z) / np.linalg.norm(ref.z)
    assert rel < 1e-9, (name, kw, rel)
    # the row-partitioned result holds this rank's block only
    part = sol.z.row_partition
    assert sol.z.A.shape[1] == int(part[rank + 1] - part[rank])
    if rank == 0:
        print(f"{name} {kw}: its total {int(native.SOL_main['its'].sum())}, rel {rel:.2e}, assemblies {sol.stats['assemblies']}, "
              f"f0 {sol.stats['f0_evals']}", flush=True)
dist.barrier()
if rank == 0:
    print("SOLVE_OK")
dist.destroy_process_group()
