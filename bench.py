#!/usr/bin/env python
"""bench.py - ms per Newton-step assembly (gradient + restricted Hessian + objective) on fem2d L=8.

One "step" = one pass of the hot path at a fixed seeded feasible iterate on the finest level:
apply_D -> barrier F/F1/F2 -> gradient -> Hessian numeric phase -> R'HR values (BASELINE.json metric).

  value     device-resident inputs, CUDA events on the launching stream, L2 flushed between steps
  e2e       the same step through the C ABI with HOST buffers (mgb_assemble_host): H2D of the
            Newton unknown, D2H of gradient + Hessian values + scalars inside the timed region
  roofline  dominant (= longest) kernel: algorithmic bytes / event time vs measured HBM peak, next to the
            fraction by ncu DRAM traffic and by the bytes this two-kernel design has to move
  cpu_baseline / --impl reference: the CPU oracle restatement (the Julia reference cannot run here:
            no julia/mpiexec in the image) timed on the host cores.

N > 1 (torchrun): owner-computes sharding - the rows of R'HR and the gradient entries are split over the ranks (rank r
owns block r of every variable: the unknowns are renumbered rank-major at the boundary, `--ownership contiguous` keeps
the reference's stacked numbering instead), a rank evaluates every element touching its rows and completes them
locally; only the objective scalars cross NVLink (peer-memory words, written and awaited inside the gather kernel, all inside the
timed region).  The fixed L=8 problem is split, so scaling = "strong"; `sub_records` adds L=9 (and L=10 at N >= 8) on the same GPUs.
Every line carries `parity`: the buffers that were just timed against the CPU oracle (max over ranks).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ms per Newton-step assembly (fem2d L=8)"


def build_problem(L: int, p: float, seed: int = 20261018):
    import mgb_b200
    from mgb_b200 import amg as amg_mod
    geom = mgb_b200.fem2d(L)
    M, _ = amg_mod.amg(geom)
    n = geom.x.shape[0]
    g, f = amg_mod.DEFAULT_G[2], amg_mod.DEFAULT_F[2]
    z0 = np.array([g(geom.x[i]) for i in range(n)], dtype=float).reshape(-1, order="F")
    c = np.array([f(geom.x[i]) for i in range(n)], dtype=float)
    R = M.R_fine[-1]
    rng = np.random.default_rng(seed)
    s = 1e-3 * rng.uniform(-1.0, 1.0, size=R.shape[1])
    Dz0 = np.stack([Dk @ z0 for Dk in M.D], axis=1)
    return dict(geom=geom, M=M, R=R, D=M.D, z0=z0, c=c, s=s, Dz0=Dz0, idx=[1, 2, 3], p=p)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._th = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([v.strip() for v in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()

    def stop(self):
        self._stop.set()
        if self._th:
            self._th.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows for k in range(4) if len(r) >= 6 and r[2 + k] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel: str, L: int):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture of this bench command at N = 1
    (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        with open(path) as fh:
            d = json.load(fh)
        return d.get(f"L{L}", {}).get(kernel)
    except Exception:
        return None


def cpu_assembly_sample(pr, t, reps):
    """CPU oracle restatement of one assembly (f1 + f2 + f0), `reps` times, one core; ms per assembly."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mgb_oracle as O
    Q = O.EuclidianPower(idx=pr["idx"], p=pr["p"])
    args = (pr["s"], pr["geom"].x, pr["geom"].w, t * pr["c"], pr["R"], pr["D"], pr["z0"], Q)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        O.f0(*args)
        O.f1(*args)
        O.f2(*args)
        times.append((time.perf_counter() - t0) * 1e3)
    return float(np.mean(times)) if times else float("nan")


_W = {}


def available_cores() -> int:
    """host cores this process may really use: the smallest of cpu_count, the affinity mask and the cgroup CPU quota
    (a container can see 16 CPUs and be allowed 2: forking 16 ranks there measures the scheduler, not the code)"""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        with open("/sys/fs/cgroup/cpu.max") as fh:
            quota, period = fh.read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        pass
    try:  # cgroup v1
        with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as fh:
            quota = int(fh.read())
        with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as fh:
            period = int(fh.read())
        if quota > 0 and period > 0:
            n = min(n, max(1, quota // period))
    except Exception:
        pass
    return max(1, n)


def _single_thread():
    """one thread per rank, like an MPI rank of the reference (BLAS threads via env, docs/src/guide.md:221-230)"""
    try:
        from threadpoolctl import threadpool_limits
        _W["tp"] = threadpool_limits(limits=1)
    except Exception:
        pass


def _worker_init(pr, t, blocks):
    _single_thread()
    _W.update(pr=pr, t=t, blocks=blocks)


def _worker_local(k):
    """this rank's local blocks, built once (setup, like native_to_mpi): quadrature rows [r0,r1) of every operator
    with the columns compressed to the ones the rows touch (HPCSparseMatrix stores its local block the same way:
    col_indices / ncols_compressed, reference src/MultiGridBarrierMPI.jl:216-221), and R restricted likewise."""
    cache = _W.setdefault("cache", {})
    if k not in cache:
        import scipy.sparse as sp
        pr = _W["pr"]
        r0, r1 = _W["blocks"][k]
        Dl = [sp.csr_matrix(d[r0:r1]) for d in pr["D"]]
        cols = np.unique(np.concatenate([d.indices for d in Dl]))
        Dc = [sp.csr_matrix(d[:, cols]) for d in Dl]
        Rc = sp.csr_matrix(pr["R"].tocsr()[cols, :])
        dofs = np.unique(Rc.indices)
        cache[k] = dict(D=Dc, R=sp.csr_matrix(Rc[:, dofs]), z0=pr["z0"][cols], s=pr["s"][dofs],
                        x=pr["geom"].x[r0:r1], w=pr["geom"].w[r0:r1], c=pr["c"][r0:r1])
    return cache[k]


def _worker_run(k):
    """one 'rank' of the CPU path: the oracle on its block of quadrature rows (the partial Hessian
    stays on the rank, as in the reference's row-partitioned MPI layout)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mgb_oracle as O
    pr, t = _W["pr"], _W["t"]
    loc = _worker_local(k)
    Q = O.EuclidianPower(idx=pr["idx"], p=pr["p"])
    args = (loc["s"], loc["x"], loc["w"], t * loc["c"], loc["R"], loc["D"], loc["z0"], Q)
    t0 = time.perf_counter()
    f0 = O.f0(*args)
    g = O.f1(*args)
    H = O.f2(*args)
    dt = time.perf_counter() - t0
    return f0, float(np.abs(g).sum()), float(abs(H).sum()), dt


def cpu_assembly_parallel(pr, t, reps, nproc):
    """The same restatement sharded over `nproc` processes by quadrature-row blocks (what
    `mpiexec -n nproc` does in the reference); ms per assembly = wall time of the slowest rank + fork/join."""
    import multiprocessing as mp
    from mgb_b200.hpc import uniform_partition
    n, B = pr["geom"].x.shape[0], pr["geom"].block
    part = uniform_partition(n, nproc, B)
    blocks = [(int(part[k] - 1), int(part[k + 1] - 1)) for k in range(nproc)]
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(nproc, initializer=_worker_init, initargs=(pr, t, blocks)) as pool:
        # one task per worker (chunksize 1); two warm-up rounds build every rank's local blocks in whichever
        # worker gets it (setup, not timed)
        pool.map(_worker_run, range(nproc), chunksize=1)
        pool.map(_worker_run, range(nproc), chunksize=1)
        for _ in range(reps):
            t0 = time.perf_counter()
            res = pool.map(_worker_run, range(nproc), chunksize=1)
            times.append((time.perf_counter() - t0) * 1e3)
    return float(np.mean(times)), float(max(r[3] for r in res) * 1e3)


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU path cannot run (no julia / mpiexec in the image); its
    restatement (oracle/mgb_oracle.py) is timed on the host cores, sharded by row blocks like MPI ranks."""
    if rank != 0:
        return
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(var, "1")
    _single_thread()
    pr = build_problem(args.L, args.p)
    avail = available_cores()
    cores = max(1, min(avail, args.cpu_procs if args.cpu_procs > 0 else avail))
    steps = max(1, args.steps)
    if cores > 1:
        ms, ms_slowest = cpu_assembly_parallel(pr, args.t, steps, cores)
    else:
        ms = ms_slowest = cpu_assembly_sample(pr, args.t, steps)
    ms_1 = cpu_assembly_sample(pr, args.t, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.L, args.p, pr['geom'].x.shape[0]),
                   "note": "CPU restatement (oracle/mgb_oracle.py, scipy CSC), NOT the Julia reference: julia/mpiexec are "
                           "absent from this image; sharded over processes by row blocks like `mpiexec -n cores`",
                   "one_core_ms": ms_1, "slowest_rank_compute_ms": ms_slowest,
                   "host_cpus_visible": os.cpu_count(), "host_cpus_usable": avail},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "port",
                         "sample": f"{steps} full assemblies (f0+f1+f2) at L={args.L}"},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_name(L, p, n):
    return f"fem2d L={L} p={p} finest-level assembly (n={n})"


def oracle_outputs(pr, t):
    """CPU oracle (checker, never timed here): objective, gradient, R'HR of the bench problem"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mgb_oracle as O
    Q = O.EuclidianPower(idx=pr["idx"], p=pr["p"])
    args = (pr["s"], pr["geom"].x, pr["geom"].w, t * pr["c"], pr["R"], pr["D"], pr["z0"], Q)
    return O.f0(*args), O.f1(*args), O.f2(*args).tocsr()


def parity_block(pr, t, f0, grad_own, hval_own, rowptr, colidx, lo, hi, perm=None):
    """relative errors of the buffers that were just timed against the oracle's rows [lo, hi) (``perm``: the plan's
    renumbering of the unknowns, new -> old: the oracle's outputs are renumbered the same way first)"""
    import scipy.sparse as sp
    f0_o, g_o, H_o = oracle_outputs(pr, t)
    if perm is not None:
        g_o, H_o = g_o[perm], H_o[perm][:, perm].tocsr()
    m = H_o.shape[0]
    Hc = sp.csr_matrix((hval_own, colidx.astype(np.int64), rowptr.astype(np.int64)), shape=(hi - lo, m))
    hn, gn = abs(H_o).max(), np.abs(g_o).max()
    return {"f0_rel": float(abs(f0 - f0_o) / abs(f0_o)), "grad_rel": float(np.abs(grad_own - g_o[lo:hi]).max() / gn),
            "hess_rel": float(abs(Hc - H_o[lo:hi]).max() / hn)}


def run_config(args, L, rank, world, local_rank, dev, ctx, stream, steps, warmup, want_cpu=False):
    """one timed configuration (fem2d level L on `world` GPUs); returns the record (rank 0) or None"""
    import torch
    import torch.distributed as dist
    from mgb_b200 import capi
    from mgb_b200 import dist as mdist

    pr = build_problem(L, args.p)
    geom = pr["geom"]
    n = geom.x.shape[0]
    B = geom.block
    sharded = world > 1
    t_plan = time.perf_counter()
    if sharded:
        plan = mdist.create_peer_plan(ctx, pr["D"], pr["R"], geom.x, geom.w, pr["idx"], pr["p"], B, rank, world,
                                      colocate=args.ownership == "colocated")
        rows = plan.rows
    else:
        plan = capi.Plan(ctx, pr["D"], pr["R"], geom.x, geom.w, pr["idx"], pr["p"])
        rows = np.arange(n)
    t_plan = time.perf_counter() - t_plan
    flags = capi.WANT_F0 | capi.WANT_GRAD | capi.WANT_HESS
    f64 = torch.float64
    perm = plan.perm if sharded else None    # rank-major renumbering of the unknowns (dist.colocated_partition)
    s_d = torch.from_numpy(pr["s"] if perm is None else pr["s"][perm]).to(dev)
    Dz0_d = torch.from_numpy(np.ascontiguousarray(pr["Dz0"][rows].T)).to(dev)   # (nD, n_local) = column-major n_local x nD
    c_d = torch.from_numpy(np.ascontiguousarray(pr["c"][rows].T)).to(dev)
    n_h = plan.dinfo["n_own_h"] if sharded else plan.nnzH
    n_g = plan.dinfo["n_own_g"] if sharded else plan.m
    scal_d = torch.zeros(4, dtype=f64, device=dev)
    grad_d = torch.zeros(max(n_g, 1), dtype=f64, device=dev)
    hval_d = torch.zeros(max(n_h, 1), dtype=f64, device=dev)
    flush = 0 if args.no_flush else (2 if args.flush_mode == "read" else 1)
    flush_buf = torch.zeros(256 << 17, dtype=f64, device=dev) if sharded else None  # 256 MiB

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    views = {}

    def window_views(ptrs):
        """zero-copy tensors over the owned results (library-owned device memory)"""
        if ptrs not in views:
            views[ptrs] = tuple(torch.as_tensor(capi.DeviceView(p_, cnt), device=dev)
                                for p_, cnt in zip(ptrs, (max(n_h, 1), max(n_g, 1), 4)))
        return views[ptrs]

    s_own_al = s_d[plan.dinfo["own0"]: plan.dinfo["own1"]].clone() if sharded else None

    def step_multi(nsteps, align=False):
        """per-step CUDA events on the launching stream, the cross-rank sum of the scalars inside the timed bracket
        (the gather kernel of every rank waits for its peers' words).  All steps are enqueued before the host waits,
        so the ranks are paced by their GPUs and not by host launch jitter.  ``align`` (side figure only): a device-side
        rendezvous (the peer-memory all-gather of s: publish + wait, untimed) lines the ranks up before each start
        event, which separates skew between the ranks (flush-time differences) from the assembly itself."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
        # the barrier in front of the timed region, on the DEVICE: the hosts leave dist.barrier() tens to hundreds of
        # microseconds apart, and since every assembly waits for every peer, the first step of the early ranks would
        # absorb that start-up skew (measured at N=8, 6 steps: 76.9 us per step with it, 36.8 without)
        plan.s_publish(s_own_al)
        plan.s_wait()
        for r in range(nsteps):
            if flush == 2:
                flush_buf.sum()  # read-evict: leaves L2 full of clean lines
            elif flush:
                flush_buf.fill_(float(r))
            if align:
                plan.s_publish(s_own_al)
                plan.s_wait()
            evs[r][0].record()
            plan.dist_assemble(s_d, Dz0_d, c_d, args.t, flags)
            evs[r][1].record()
        torch.cuda.synchronize(dev)
        if plan.dist_info()["err"]:
            raise SystemExit("bench.py: a peer's objective partials timed out (scalar exchange protocol error)")
        return sum(a.elapsed_time(b) for a, b in evs) / nsteps

    # ---- warm-up
    if not sharded:
        plan.time_assemble(s_d, Dz0_d, c_d, args.t, flags, scal_d, grad_d, hval_d, warmup, flush, split=False)
    else:
        step_multi(warmup)
    launches0 = capi.launch_count()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    if not sharded:
        ms_total, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, args.t, flags, scal_d, grad_d, hval_d, steps, flush, split=False)
    else:
        ms_total = step_multi(steps)
    launches = capi.launch_count() - launches0 - (2 if sharded else 0)   # minus the rendezvous in front of the loop
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    ms_aligned = None   # side figure: the same steps with the ranks lined up on the device before each start event
    if sharded:
        barrier()
        ms_aligned = step_multi(steps, align=True)
    ms_sdist = None
    if sharded:   # the same step starting from a row-distributed unknown (the reference's HPCVector): + all-gather of s
        s_own = s_d[plan.dinfo["own0"]: plan.dinfo["own1"]].clone()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10 + 3)]
        for a, b in evs:
            if flush:
                flush_buf.sum()
            a.record()
            plan.dist_assemble_s(s_own, Dz0_d, c_d, args.t, flags)
            b.record()
        torch.cuda.synchronize(dev)
        ms_sdist = sum(a.elapsed_time(b) for a, b in evs[3:]) / 10
    # ---- parity of the buffers that were just timed (every rank: its owned rows against the oracle's)
    if sharded:
        ho, go, so = window_views(plan.dist_assemble(s_d, Dz0_d, c_d, args.t, flags))
        torch.cuda.synchronize(dev)
        lo, hi = plan.dinfo["own0"], plan.dinfo["own1"]
        scal_now, g_now, h_now = so.cpu().numpy(), go[:n_g].cpu().numpy(), ho[:n_h].cpu().numpy()
    else:
        lo, hi = 0, plan.m
        scal_now, g_now, h_now = scal_d.cpu().numpy(), grad_d.cpu().numpy(), hval_d[:n_h].cpu().numpy()
    rp, ci = plan.pattern()
    parity = parity_block(pr, args.t, scal_now[0], g_now, h_now, rp, ci, lo, hi, perm) if not args.no_parity else None
    # per-kernel split (separate pass, not part of `value`): this rank's element / gather kernels
    # (N > 1: the rank's own two kernels without the cross-rank scalar exchange)
    _, ms_elem, ms_gather = plan.time_assemble(s_d, Dz0_d, c_d, args.t, flags, scal_d, grad_d, hval_d,
                                               max(10, steps // 2), flush, split=True)
    ms_f0 = None
    if not sharded:   # line-search point: objective only (reported as a separate line, ms per f0)
        plan.time_assemble(s_d, Dz0_d, c_d, args.t, capi.WANT_F0, scal_d, grad_d, hval_d, 3, flush, split=False)
        ms_f0, _, _ = plan.time_assemble(s_d, Dz0_d, c_d, args.t, capi.WANT_F0, scal_d, grad_d, hval_d,
                                         max(10, steps // 2), flush, split=False)
    # ---- e2e: the same step through HOST buffers.  N = 1: the C-ABI host call (mgb_assemble_host) on the caller's
    # page-locked arrays - H2D of the Newton unknown, D2H of gradient + Hessian values + scalars inside the call.
    # N > 1: pinned torch buffers around mgb_dist_assemble (the owned blocks come back).
    e2e_steps = max(5, min(steps, 20))
    e2e_extra = {}
    if not sharded:
        plan.assemble_host(pr["s"], np.asfortranarray(pr["Dz0"]), np.asfortranarray(pr["c"]), args.t, flags, True)
        t0 = time.perf_counter()
        for _ in range(5):
            plan.assemble_host(pr["s"], None, None, args.t, flags, upload_inputs=False)
        e2e_extra["c_abi_pageable_ms"] = (time.perf_counter() - t0) * 1e3 / 5
        s_reg = pr["s"].copy()
        outb = plan.assemble_host(s_reg, None, None, args.t, flags, upload_inputs=False)
        regs = (s_reg, outb["hval"], outb["grad"], outb["scal"])
        for a in regs:
            capi.host_register(a)
        for _ in range(3):
            plan.assemble_host(s_reg, None, None, args.t, flags, upload_inputs=False, out=outb)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            plan.assemble_host(s_reg, None, None, args.t, flags, upload_inputs=False, out=outb)
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        for a in regs:
            capi.host_unregister(a)
        e2e_extra["path"] = "mgb_assemble_host on arrays page-locked once with mgb_host_register"
    else:
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        s_h = pin(pr["s"] if perm is None else pr["s"][perm])
        hval_h = torch.empty(max(n_h, 1), dtype=f64).pin_memory()
        grad_h = torch.empty(max(n_g, 1), dtype=f64).pin_memory()
        scal_h = torch.empty(4, dtype=f64).pin_memory()

        def e2e_step():
            s_d.copy_(s_h, non_blocking=True)
            ho, go, so = window_views(plan.dist_assemble(s_d, Dz0_d, c_d, args.t, flags))
            hval_h[:n_h].copy_(ho[:n_h], non_blocking=True)
            grad_h[:n_g].copy_(go[:n_g], non_blocking=True)
            scal_h.copy_(so, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        barrier()   # the ranks leave the (CPU) parity check seconds apart; the peer wait inside an assembly times out after 30 s
        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        e2e_extra["path"] = "pinned host buffers around mgb_dist_assemble (owned blocks)"

    par = [parity[k] if parity else 0.0 for k in ("f0_rel", "grad_rel", "hess_rel")]
    vals = torch.tensor([ms_total, ms_elem, ms_gather, e2e_ms] + par + [float(rows.size), float(n_h), ms_sdist or 0.0, ms_aligned or 0.0],
                        dtype=f64, device=dev)
    mx = vals.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    ms_total, ms_elem, ms_gather, e2e_ms = (float(v) for v in mx[:4].cpu())
    rec = None
    if rank == 0:
        peak, peak_src = peak_hbm()
        info = plan.info
        nD, dim, nu = info["nD"], 2, info["nu"]
        nnzD = int(sum(d.nnz for d in pr["D"]))
        nnzR = int(pr["R"].nnz)
        if sharded:   # algorithmic bytes are a property of the whole level: take them from a symbolic global plan
            ginfo = capi.Plan(None, pr["D"], pr["R"], geom.x, geom.w, pr["idx"], pr["p"]).info
        else:
            ginfo = info
        alg, nnzH, N, E = ginfo["alg_bytes"], ginfo["nnzH"], ginfo["N"], ginfo["elements"]
        # B = apply_D + barrier + y1 + y2 + gradient + Hessian numeric + restriction; the fine-space
        # pattern size nnz(sum_jk D_j' D_k) is recovered from the library's total (it enters twice, 8 bytes each)
        y2u = nD * (nD + 1) // 2
        other = (N * 8 + nnzD * 12 + nD * (n + 1) * 4 + n * nD * 8) + n * (nD + dim + 1 + nD) * 8 + n * nD * 8 + n * y2u * 8 \
            + (nnzD * 12 + n * nD * 8 + N * 8) + (n * y2u * 8 + nnzD * 8) + (nnzR * 24 + nnzH * 8)
        nnzS = (alg - other) // 16
        restr = nnzS * 8 + nnzR * 24 + nnzH * 8
        shares = {"element_kernel": alg - restr, "gather_kernel": restr}
        # bytes the two-kernel design itself has to move (compulsory traffic of THIS implementation, for the honest
        # fraction next to the SURVEY formula): element kernel reads records / ids / c / Dz0 / s and writes the slot
        # and gradient records, the gather kernel reads them back with the index lists and writes R'HR and g
        RW = (dim * B + 1 + nu + 1 + 1) // 2 * 2
        NS = ginfo["slots_per_element"]
        design = {"element_kernel": n * RW * 8 + E * nu * 8 * 4 + 2 * n * nD * 8 + ginfo["m"] * 8 + E * NS * 8 + E * nu * 8 * 8,
                  "gather_kernel": nnzH * 8 + E * NS * 8 + nnzH * 8 + ginfo["grad_contribs"] * 12 + ginfo["m"] * 8}
        kms = {"element_kernel": ms_elem, "gather_kernel": ms_gather}
        dom = max(kms, key=kms.get) if not sharded else None
        traffic = ncu_traffic(dom, L) if dom else None
        ach_all = alg / (ms_total * 1e-3) / 1e9
        roof = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                "assembly": {"algorithmic_bytes": int(alg), "ms": ms_total, "achieved": ach_all, "frac": ach_all / peak,
                             "design_bytes": int(sum(design.values())),
                             "design_frac": sum(design.values()) / (ms_total * 1e-3) / 1e9 / peak,
                             "element_ms": ms_elem, "gather_ms": ms_gather, "f0_ms": ms_f0}}
        if dom:
            ach = shares[dom] / (kms[dom] * 1e-3) / 1e9
            roof.update({"kernel": dom, "kernel_ms": kms[dom], "algorithmic_bytes": int(shares[dom]), "achieved": ach,
                         "frac": ach / peak, "traffic": traffic,
                         "traffic_frac": (traffic / (kms[dom] * 1e-3) / 1e9 / peak) if traffic else None,
                         "design_bytes": int(design[dom]), "design_frac": design[dom] / (kms[dom] * 1e-3) / 1e9 / peak})
        else:   # N > 1: per-kernel split and ncu traffic are single-GPU measurements; the whole assembly stands in
            roof.update({"kernel": "element_kernel+gather_kernel (whole assembly, max over ranks)", "achieved": ach_all,
                         "frac": ach_all / peak, "traffic": None})
        rec = {
            "value": ms_total, "ms_per_step": ms_total,
            "config": {"workload": workload_name(L, args.p, n), "m": ginfo["m"], "nnzH": nnzH,
                       "l2": ("flushed between steps (256 MiB %s)" % args.flush_mode) if flush else "not flushed",
                       "iterate": "boundary lift g(x)=[x1^2+x2^2,100] + 1e-3*U(-1,1), seed 20261018",
                       "path": "element" if info["path"] == 1 else "csr", "plan_seconds": round(t_plan, 3),
                       "rows_evaluated_max_rank": int(mx[7].cpu()), "rows_total": n, "owned_hessian_entries_max_rank": int(mx[8].cpu()),
                       "ms_from_row_distributed_s": (float(mx[9].cpu()) if sharded else None),
                       "ownership": None if not sharded else args.ownership,
                       "ms_rank_aligned": (float(mx[10].cpu()) if sharded else None),
                       "multi_gpu": None if not sharded else (
                           "owner-computes: a rank evaluates every element touching its output rows "
                           + ("(unknowns renumbered rank-major - rank r owns block r of u AND of s - so that is E/P elements "
                              "plus a thin halo) " if perm is not None else
                              "(contiguous blocks of the stacked [u | s] unknowns: each element on about two ranks) ") +
                           "and completes its rows of R'HR / block of g locally; only the 3 objective scalars cross "
                           "NVLink, as epoch-tagged peer-memory words written and awaited inside the gather kernel; no NCCL on "
                           "the data path")},
            "clocks": clocks,
            "e2e": dict({"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(ginfo["m"] * 8),
                         "d2h_bytes_per_step": int((n_g + n_h + 4) * 8), "host_memory": "pinned / page-locked"}, **e2e_extra),
            "gpu_launches": int(launches), "pdl_active": capi.pdl_active(), "cuda_graph": (plan.graph_stats() if not sharded else None),
            "parity": None if parity is None else {"f0_rel": float(mx[4].cpu()), "grad_rel": float(mx[5].cpu()), "hess_rel": float(mx[6].cpu()),
                                                    "against": "CPU oracle on the same inputs; max over ranks of each rank's owned rows",
                                                    "tolerance": 1e-12},
            "roofline": roof,
            "wall_s_timed_region": wall,
        }
    views.clear()
    if sharded:
        mdist.destroy_peer_plan(plan)
    else:
        plan.close()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--L", type=int, default=8)
    ap.add_argument("--p", type=float, default=1.0)
    ap.add_argument("--t", type=float, default=1.0)
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--flush-mode", default="read", choices=["read", "write"], help="evict L2 by reading (clean lines) or writing (dirty lines) 256 MiB")
    ap.add_argument("--cpu-reps", type=int, default=3)
    ap.add_argument("--cpu-procs", type=int, default=0, help="processes for the CPU restatement (0 = all cores)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the timed buffers")
    ap.add_argument("--ownership", default="colocated", choices=["colocated", "contiguous"],
                    help="N > 1: rank r owns block r of every variable (unknowns renumbered rank-major) / a contiguous block of the stacked unknowns")
    ap.add_argument("--sub", default=None, help="comma-separated extra levels timed as sub-records (default: 9 when N > 1, 9,10 when N >= 8; \"\" for none)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from mgb_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(dev)      # a real (non-legacy) stream: programmatic dependent launch needs one
    torch.cuda.set_stream(stream)
    ctx = capi.Context(local_rank, stream.cuda_stream)

    rec = run_config(args, args.L, rank, world, local_rank, dev, ctx, stream, args.steps, args.warmup)
    subs = []
    # default sub-records: L=9 on any sharded run (with L=8 on one GPU and L=9 on four: the weak-scaling pair), plus the
    # 3.67 M-point L=10 mesh on a full box (strong scaling against profiles/r2_bench_L10_n1.json; adds about a minute)
    default_sub = [] if world == 1 else (["9", "10"] if world >= 8 else ["9"])
    sub_levels = [int(v) for v in (args.sub.split(",") if args.sub else (default_sub if args.sub is None else [])) if v]
    for Ls in sub_levels:   # larger meshes on the same GPUs: where sharding has work to split
        r = run_config(args, Ls, rank, world, local_rank, dev, ctx, stream, max(5, args.steps // 5), 3)
        if rank == 0:
            subs.append({"workload": r["config"]["workload"], "ms_per_step": r["value"], "parity": r["parity"],
                         "rows_evaluated_max_rank": r["config"]["rows_evaluated_max_rank"], "rows_total": r["config"]["rows_total"],
                         "ms_rank_aligned": r["config"]["ms_rank_aligned"], "ms_from_row_distributed_s": r["config"]["ms_from_row_distributed_s"],
                         "element_ms": r["roofline"]["assembly"]["element_ms"], "gather_ms": r["roofline"]["assembly"]["gather_ms"],
                         "assembly_frac_of_peak": r["roofline"]["assembly"]["frac"]})
    if rank == 0:
        line = {"metric": METRIC, "value": rec["value"], "unit": "ms", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
        line.update({k: rec[k] for k in ("config", "clocks", "e2e", "gpu_launches", "pdl_active", "cuda_graph", "parity", "roofline", "wall_s_timed_region")})
        if subs:
            line["sub_records"] = subs
        if world == 1 and args.cpu_reps > 0:
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--L", str(args.L),
                                      "--p", str(args.p), "--t", str(args.t), "--steps", str(args.cpu_reps), "--warmup", "0",
                                      "--cpu-procs", str(args.cpu_procs)], capture_output=True, text=True, timeout=900)
                ref = json.loads(out.stdout.strip().splitlines()[-1])
                line["cpu_baseline"] = ref["cpu_baseline"]
                line["cpu_baseline"]["one_core_ms"] = ref["config"]["one_core_ms"]
                line["cpu_baseline"]["note"] = "restatement written for this project, not the Julia reference (cannot run here)"
            except Exception as exc:  # pragma: no cover
                line["cpu_baseline_error"] = str(exc)
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
