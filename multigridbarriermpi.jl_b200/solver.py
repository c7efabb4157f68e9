"""Host driver for the barrier solve with device-resident state: ``amgb`` / ``parabolic_solve``.

The outer algorithm (central path, multigrid sweep, damped Newton) is upstream
``MultiGridBarrier.amgb`` - outside the reference repository, reached through
``fem{1,2,3}d_mpi_solve`` (reference src/MultiGridBarrierMPI.jl:594-600, 661-667, 735-745).  It is
restated here so the GPU assembly can be driven end to end; every f0/f1/f2 evaluation goes through
the C ABI (``capi.Plan.assemble``), and the only work left on the host per Newton step is the sparse
direct solve ``H \\ g`` - the solve seam the north-star leaves outside the graft
(reference test/test_newton_matrix_compare.jl:33-51), timed separately in ``SOL_main``.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

from . import capi
from .amg import AMG, DEFAULT_D, DEFAULT_F, DEFAULT_G, DEFAULT_STATE, amg_helper
from .geometry import Geometry

EPS = float(np.finfo(np.float64).eps)
# Newton stops once the decrement is within the rounding noise of the objective; the margin keeps the
# accept/stop decisions identical between implementations whose f0 agree to ~1e-13 relative
NEWTON_NOISE = 1024.0


def solve(H: sp.spmatrix, g: np.ndarray) -> np.ndarray:
    """MultiGridBarrier.solve(A, b) = A \\ b.  Host sparse LU stands in for MUMPS (outside the graft).
    The barrier Hessian is symmetric positive definite: minimum-degree ordering on A'+A with diagonal pivots
    halves SuperLU's fill against the default COLAMD (fem2d L=6: 1.28 M vs 2.37 M factor entries)."""
    A = sp.csc_matrix(H)
    try:
        return spla.splu(A, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True)).solve(g)
    except RuntimeError:   # exactly singular pivot without partial pivoting: fall back to the general ordering
        return spla.splu(A).solve(g)


class LevelState:
    """Plan + preallocated device buffers of one multigrid level.  ``assemble`` leaves the objective scalars in
    ``scal``, the gradient in ``grad`` and the CSR values of R'HR in ``hval`` - complete on every rank."""

    def __init__(self, prob: "DeviceProblem", J: int):
        M = prob.M
        dev = prob.device
        self.sharded = False
        # the level's unknowns in the numbering this object works in: a sharded level renumbers them rank-major
        # (dist.colocated_partition; R_level = R[:, perm]) - Newton on s -> f(z + R_level s) does not care, and z itself
        # lives in the fine space
        R_level = M.R_fine[J]
        if prob.nranks > 1:
            from . import dist as mdist
            try:
                self.plan = mdist.create_peer_plan(prob.ctx, M.D, M.R_fine[J], M.x, M.w, prob.idx, prob.p, prob.block,
                                                   prob.rank, prob.nranks, group=prob.group, slack=prob.slack,
                                                   idx2=prob.idx2, p2=prob.p2)
                self.sharded = True
                R_level = self.plan.R
            except capi.MgbError as exc:
                # only the levels on the element path with thread-per-entry gather (the fine ones, where the work
                # is) are sharded; any other level - coarse levels, fem3d's Q3 elements and every operator table
                # on the CSR path - is assembled redundantly by every rank instead: same inputs, same kernels,
                # hence bit-identical results on all ranks and no communication.  Every rank takes this branch
                # together: the refusal depends on the (replicated) operators only.
                if "sharded plans need the element path" not in str(exc):
                    raise
        if self.sharded:
            # the replicated symbolic plan gives the global pattern the solve seam needs (owned row blocks are
            # contiguous, so the global value array is the concatenation of the ranks' owned values)
            sym = capi.Plan(None, M.D, R_level, M.x, M.w, prob.idx, prob.p, slack=prob.slack, idx2=prob.idx2, p2=prob.p2)
            rp, ci = sym.pattern()
            m, nnz = sym.m, sym.nnzH
            d = self.plan.dinfo
            self.own_h = (int(rp[d["own0"]]), d["n_own_h"])
            self.own_g = (d["own0"], d["n_own_g"])
            assert self.own_h[0] + self.own_h[1] == int(rp[d["own1"]])
            self._views = {}
            # the plan evaluates its own list of quadrature rows (primary elements + halo): Dz0 / c blocks over them
            self.rows_idx = torch.from_numpy(self.plan.rows).to(dev)
            self.Dz0_loc = torch.zeros((M.nD, self.plan.rows.size), dtype=torch.float64, device=dev)
            self._c_cache = None
        else:
            self.plan = capi.Plan(prob.ctx, M.D, M.R_fine[J], M.x, M.w, prob.idx, prob.p,
                                  slack=prob.slack, idx2=prob.idx2, p2=prob.p2)
            m, nnz = self.plan.m, self.plan.nnzH
            rp, ci = self.plan.pattern()
        # no back-reference to `prob`: a reference cycle would hand destruction order (plans before their context)
        # to the cyclic garbage collector
        self.nranks, self.group, self.device = prob.nranks, prob.group, dev
        self.m, self.nnzH = m, nnz
        f64 = torch.float64
        self.s = torch.zeros(m, dtype=f64, device=dev)
        self.trial = torch.zeros(m, dtype=f64, device=dev)
        self.step = torch.zeros(m, dtype=f64, device=dev)
        # one buffer [hval | grad]: on several ranks a single all-reduce completes both
        self.hg = torch.zeros(max(nnz, 1) + m, dtype=f64, device=dev)
        self.hval, self.grad = self.hg[: max(nnz, 1)], self.hg[max(nnz, 1):]
        self.scal = torch.zeros(4, dtype=f64, device=dev)
        self.R = capi.SpMat(prob.ctx, R_level)
        self.rowptr, self.colidx = rp.astype(np.int64), ci.astype(np.int64)
        # pinned host mirrors for the solve seam
        self.h_hval = torch.zeros(max(nnz, 1), dtype=f64).pin_memory()
        self.h_grad = torch.zeros(m, dtype=f64).pin_memory()
        self.h_step = torch.zeros(m, dtype=f64).pin_memory()

    def assemble(self, s: torch.Tensor, Dz0: torch.Tensor, c: torch.Tensor, t: float, flags: int):
        """``Dz0`` / ``c``: column-major (nD, n) blocks over ALL quadrature rows (the unknown is replicated)"""
        if not self.sharded:
            self.plan.assemble(s, Dz0, c, t, flags, self.scal, self.grad, self.hval)
            return
        # sharded: every rank evaluates the elements that touch its output rows and completes its own rows of R'HR /
        # block of the gradient; the objective scalars arrive globally summed (mgb_dist_assemble).  The host solve
        # below needs the whole system, so the owned blocks are then replicated (solve seam only).
        import torch.distributed as dist
        torch.index_select(Dz0, 1, self.rows_idx, out=self.Dz0_loc)
        if self._c_cache is None or self._c_cache[0] is not c:
            self._c_cache = (c, torch.index_select(c, 1, self.rows_idx).contiguous())
        ptrs = self.plan.dist_assemble(s, self.Dz0_loc, self._c_cache[1], t, flags)
        if ptrs not in self._views:
            cnt = (max(self.own_h[1], 1), max(self.own_g[1], 1), 4)
            self._views[ptrs] = tuple(torch.as_tensor(capi.DeviceView(p_, n_), device=self.device) for p_, n_ in zip(ptrs, cnt))
        vh, vg, vs = self._views[ptrs]
        self.scal.copy_(vs)
        if flags & (capi.WANT_GRAD | capi.WANT_HESS):
            self.hg.zero_()
            if flags & capi.WANT_HESS:
                self.hval[self.own_h[0]: self.own_h[0] + self.own_h[1]] = vh[: self.own_h[1]]
            if flags & capi.WANT_GRAD:
                self.grad[self.own_g[0]: self.own_g[0] + self.own_g[1]] = vg[: self.own_g[1]]
            dist.all_reduce(self.hg, group=self.group)   # every entry has exactly one non-zero contributor: exact

    def read_scal(self) -> torch.Tensor:
        """scalars on the host; a sharded level whose peer never delivered its partial sums raises on every rank
        that noticed (the library poisons the scalars instead of returning a partial sum)"""
        sc = self.scal.cpu()
        if self.sharded and float(sc[3]) < 0.0 and self.plan.dist_info()["err"]:
            raise RuntimeError("sharded assembly: a peer's objective partials did not arrive within MGB_DIST_TIMEOUT_S")
        return sc

    def close(self):
        if self.sharded:
            from . import dist as mdist
            self._views.clear()
            mdist.destroy_peer_plan(self.plan, group=self.group)
        else:
            self.plan.close()


class DeviceProblem:
    """One AMG hierarchy resident on one GPU - or, with ``nranks > 1`` (one process per GPU, torch.distributed
    initialised), this rank's shard of it: quadrature rows split on element boundaries as HPCSparseArrays
    splits them (hpc.uniform_partition), the unknown ``z`` replicated, every level plan a DistPlan."""

    def __init__(self, M: AMG, idx: Sequence[int], p: float, slack: bool = False, device: int = 0,
                 ctx: Optional[capi.Context] = None, idx2: Optional[Sequence[int]] = None, p2: float = 2.0,
                 rank: int = 0, nranks: int = 1, block: int = 1, group=None):
        self.M, self.idx, self.p, self.slack = M, list(idx), float(p), bool(slack)
        self.idx2, self.p2 = (list(idx2) if idx2 else None), float(p2)
        self.rank, self.nranks, self.block, self.group = int(rank), int(nranks), int(block), group
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.current_stream(self.device)
        self.ctx = ctx or capi.Context(device, self.stream.cuda_stream)
        self.n = M.x.shape[0]
        self.N = M.nu * self.n
        self.levels: Dict[int, LevelState] = {}
        # operator-only plan with R = I: Dz0 = D z for any fine-space z (all quadrature rows: the unknown is
        # replicated, and a sharded level picks the rows of its own elements out of it)
        self.op_plan = capi.Plan(self.ctx, M.D, sp.identity(self.N, format="csr"), M.x, M.w, self.idx, self.p,
                                 slack=self.slack, force_path=capi.PLAN_NO_HESSIAN, idx2=self.idx2, p2=self.p2)
        self.Dz0 = torch.zeros((M.nD, self.n), dtype=torch.float64, device=self.device)  # column-major n x nD
        self.stats = dict(assemblies=0, f0_evals=0, solve_s=0.0, assemble_s=0.0)

    def level(self, J: int) -> LevelState:
        if J not in self.levels:
            t0 = time.perf_counter()
            self.levels[J] = LevelState(self, J)
            self.stats["plan_s"] = self.stats.get("plan_s", 0.0) + time.perf_counter() - t0   # symbolic phase: once per level
        return self.levels[J]

    def apply_D(self, z_dev: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Dz0 = D z on every quadrature row"""
        out = self.Dz0 if out is None else out
        self.op_plan.apply_D(z_dev, None, out)
        return out

    def close(self):
        """collective on several ranks: nobody unmaps an exchange window a peer may still store into"""
        for J in sorted(self.levels):
            self.levels[J].close()
        self.levels.clear()


def newton_device(prob: DeviceProblem, J: int, z: torch.Tensor, c: torch.Tensor, t: float, maxit: int,
                  alpha: float = 0.1, beta: float = 0.25, solve_fn: Callable = solve):
    """Damped Newton on level J for s -> f(z + R_J s); same decisions as the oracle's ``newton``."""
    lv = prob.level(J)
    Dz0 = prob.apply_D(z)
    lv.s.zero_()
    F0, FG, FH = capi.WANT_F0, capi.WANT_GRAD, capi.WANT_HESS
    t0 = time.perf_counter()
    lv.assemble(lv.s, Dz0, c, t, F0 | FG | FH)
    sc = lv.read_scal()
    prob.stats["assemblies"] += 1
    y = float(sc[0])
    if not (sc[1] == 1.0 and math.isfinite(y)):
        raise RuntimeError("newton: infeasible start")
    k, converged, stalled = 0, False, False
    while k < maxit:
        lv.h_hval.copy_(lv.hval, non_blocking=True)
        lv.h_grad.copy_(lv.grad, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        prob.stats["assemble_s"] += time.perf_counter() - t0
        ts = time.perf_counter()
        H = sp.csr_matrix((lv.h_hval.numpy()[: lv.nnzH], lv.colidx, lv.rowptr), shape=(lv.m, lv.m))
        g = lv.h_grad.numpy()
        nstep = solve_fn(H, g)
        inc = float(np.dot(g, nstep))
        prob.stats["solve_s"] += time.perf_counter() - ts
        t0 = time.perf_counter()
        if not math.isfinite(inc) or inc <= NEWTON_NOISE * EPS * max(1.0, abs(y)):
            converged = True
            break
        k += 1
        lv.h_step.copy_(torch.from_numpy(nstep))
        lv.step.copy_(lv.h_step, non_blocking=True)
        sstep, ok = 1.0, False
        while sstep > 1e-12:
            torch.add(lv.s, lv.step, alpha=-sstep, out=lv.trial)
            lv.assemble(lv.trial, Dz0, c, t, F0)
            sc = lv.read_scal()
            prob.stats["f0_evals"] += 1
            yn = float(sc[0])
            if sc[1] == 1.0 and math.isfinite(yn) and yn <= y - alpha * sstep * inc:
                ok = True
                break
            sstep *= beta
        if not ok:   # no step length decreases the objective: stagnation at rounding level (reported, not hidden)
            converged = True
            stalled = True
            k -= 1
            break
        lv.s.copy_(lv.trial)
        y = yn
        lv.assemble(lv.s, Dz0, c, t, FG | FH)
        prob.stats["assemblies"] += 1
    prob.stats["assemble_s"] += time.perf_counter() - t0
    # z <- z + R s
    lv.R.mv(lv.s, z, beta=1.0, y0_dev=z)
    return dict(k=k, converged=converged, stalled=stalled, y=y)


@dataclass
class AMGBSOL:
    """Mirror of upstream AMGBSOL (reference src/MultiGridBarrierMPI.jl:467-473): z, SOL_feasibility,
    SOL_main (ts, its, c_dot_Dz, t_elapsed), log, geometry."""
    z: np.ndarray
    SOL_feasibility: Optional[dict]
    SOL_main: dict
    log: str
    geometry: Geometry
    stats: dict = field(default_factory=dict)


@dataclass
class ParabolicSOL:
    """Mirror of upstream ParabolicSOL (reference src/MultiGridBarrierMPI.jl:512-516): geometry, ts, u.
    ``stats``: per time step wall / assembly / solve-seam seconds, Newton iterations and assembly counts, plus the
    symbolic-phase seconds spent once for all steps (the "pattern reuse" of config C5)."""
    geometry: Geometry
    ts: np.ndarray
    u: List[np.ndarray]
    stats: Optional[dict] = None


def amgb_core(prob: DeviceProblem, z: torch.Tensor, c: torch.Tensor, tol, t0, kappa, maxit, max_newton_fine,
              verbose=False, solve_fn=solve, logfile=None):
    L = len(prob.M.R_fine)
    t = t0
    ts, its, cdots = [], [], []
    unconverged, stalls = [], []   # t values whose finest-level Newton hit maxit / whose line search stagnated
    t_begin = time.time()
    kk = 0
    while t <= 1.0 / tol:
        kk += 1
        ts.append(t)
        row = [0] * L

        def level(J, mi):
            sol = newton_device(prob, J, z, c, t, mi, solve_fn=solve_fn)
            row[J] += sol["k"]
            if J == L - 1:
                fine_state["converged"], fine_state["stalled"] = sol["converged"], sol["stalled"]
            return sol["converged"]

        ok = False
        fine_state = dict(converged=False, stalled=False)
        if kk > 1:
            ok = level(L - 1, max_newton_fine)
        if not ok:
            for J in range(L):
                ok = level(J, maxit)
        its.append(row)
        if not fine_state["converged"]:
            unconverged.append(t)
        if fine_state["stalled"]:
            stalls.append(t)
        # <c, Dz>_w through the objective kernel on the finest plan (s = 0)
        lv = prob.level(L - 1)
        Dz0 = prob.apply_D(z)
        lv.s.zero_()
        lv.assemble(lv.s, Dz0, c, 0.0, capi.WANT_F0)
        cdots.append(float(lv.read_scal()[2]))
        if verbose:
            print(f"t={t:.3e} its={row} c.Dz={cdots[-1]:.12e}", file=logfile)
        t *= kappa
    if unconverged:
        import warnings
        warnings.warn(f"amgb: the finest-level Newton iteration reached maxit={maxit} without converging at t={unconverged}; "
                      "the returned iterate is not on the central path there", RuntimeWarning)
    return dict(ts=np.array(ts), its=np.array(its).T, c_dot_Dz=np.array(cdots), t_elapsed=time.time() - t_begin,
                unconverged_t=np.array(unconverged), line_search_stalls_t=np.array(stalls))


def _cm(a: np.ndarray, device) -> torch.Tensor:
    """n x k host matrix -> column-major device buffer (k, n) contiguous."""
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).T)).to(device)


def amgb(geom: Geometry, p: float = 1.0, tol: float = math.sqrt(EPS), t: float = 0.1, kappa: float = 10.0,
         maxit: int = 50, max_newton: Optional[int] = None, state_variables=DEFAULT_STATE, D=None, f=None, g=None,
         verbose: bool = False, logfile=None, device: int = 0, solve_fn: Callable = solve, **extra) -> AMGBSOL:
    """Barrier solve of the default p-Laplace-type problem on ``geom`` with GPU assembly.
    Keyword names follow the reference's documented ``amgb`` keys (docs/src/guide.md:148-152).  The geometry
    constructor's keys are accepted and ignored because ``femNd_mpi_solve`` forwards the same kwargs to the geometry
    constructor and to ``amgb`` (src/MultiGridBarrierMPI.jl:594-600); any other key (e.g. a custom convex set ``Q``)
    raises instead of being dropped silently."""
    unknown = sorted(set(extra) - GEOMETRY_KEYS)
    if unknown:
        raise TypeError(f"amgb: unsupported keyword(s) {unknown}: the GPU path implements the Euclidian power-cone barrier "
                        "of the default problem (p, D, f, g, state_variables) only")
    dim = geom.dim
    D = DEFAULT_D[dim] if D is None else list(D)
    f = DEFAULT_F[dim] if f is None else f
    g = DEFAULT_G[dim] if g is None else g
    M = amg_helper(geom, state_variables, D)
    n = geom.x.shape[0]
    idx = list(range(1, dim + 2))
    z0 = np.array([g(geom.x[i]) for i in range(n)], dtype=float)
    cmat = np.array([f(geom.x[i]) for i in range(n)], dtype=float)
    if max_newton is None:
        max_newton = int(math.ceil(math.log2(1.0 / tol) + 2))
    rank, nranks, group = _dist_layout()
    prob = DeviceProblem(M, idx, p, slack=False, device=device, rank=rank, nranks=nranks, block=geom.block, group=group)
    z = torch.from_numpy(z0.reshape(-1, order="F").copy()).to(prob.device)
    c = _cm(cmat, prob.device)
    # strict feasibility of the start (upstream skips the feasibility phase when it holds)
    lv = prob.level(len(M.R_fine) - 1)
    Dz0 = prob.apply_D(z)
    lv.s.zero_()
    lv.assemble(lv.s, Dz0, c, 0.0, capi.WANT_F0)
    sol_feas = None
    if float(lv.read_scal()[1]) != 1.0:
        z, sol_feas = feasibility_phase(geom, prob, z, cmat, state_variables, D, tol, t, kappa, maxit, solve_fn, device)
    sol_main = amgb_core(prob, z, c, tol, t, kappa, maxit, max_newton, verbose, solve_fn, logfile)
    zz = z.cpu().numpy().reshape(n, M.nu, order="F")
    stats = dict(prob.stats, nranks=nranks)
    if nranks > 1:
        prob.close()
    return AMGBSOL(zz, sol_feas, sol_main, "", geom, stats)


GEOMETRY_KEYS = {"L", "K", "k", "Ti", "backend", "T", "rest", "n"}


def _dist_layout():
    """(rank, nranks, group) of the running job: one process per GPU when torch.distributed is initialised with
    more than one rank (the reference's `mpiexec -n P`, one rank per GPU: test/test_2d.jl:17-20), else (0, 1, None)"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size(), None
    return 0, 1, None


SLACK_COST = 10.0


def feasibility_phase(geom, prob: DeviceProblem, z, cmat, state_variables, D, tol, t0, kappa, maxit, solve_fn, device):
    """Phase 1 with the extra state variable tau (:feasibility_slack, :full; operator :id): central path
    of c.Dz + SLACK_COST*tau s.t. (q, s + tau) in Q, tau > -1, until tau < 0 everywhere."""
    n = geom.x.shape[0]
    sv1 = tuple(state_variables) + (("feasibility_slack", "full"),)
    D1 = list(D) + [("feasibility_slack", "id")]
    M1 = amg_helper(geom, sv1, D1)
    prob1 = DeviceProblem(M1, prob.idx, prob.p, slack=True, device=device, ctx=prob.ctx, rank=prob.rank, nranks=prob.nranks,
                          block=prob.block, group=prob.group)
    Dz = prob.apply_D(z).cpu().numpy().T  # n x nD
    q = Dz[:, prob.idx[:-1]]
    s = Dz[:, prob.idx[-1]]
    need = np.sum(q * q, axis=1) ** (prob.p / 2.0) - s
    slack0 = max(1.0, 2.0 * float(np.max(need)) + 1.0)
    z1 = torch.cat([z, torch.full((n,), slack0, dtype=torch.float64, device=prob.device)])
    c1 = np.hstack([cmat, np.full((n, 1), SLACK_COST)])
    c1d = _cm(c1, prob.device)
    t = t0
    ts, its = [], []
    J = len(M1.R_fine) - 1
    lvm = prob.level(len(prob.M.R_fine) - 1)
    czero = torch.zeros((len(D), n), dtype=torch.float64, device=prob.device)
    while True:
        sol = newton_device(prob1, J, z1, c1d, t, maxit, solve_fn=solve_fn)
        ts.append(t), its.append(sol["k"])
        zt = z1[: prob.N].clone()
        Dz0 = prob.apply_D(zt)
        lvm.s.zero_()
        lvm.assemble(lvm.s, Dz0, czero, 0.0, capi.WANT_F0)
        feasible = float(lvm.read_scal()[1]) == 1.0
        if feasible and float(z1[prob.N:].max().cpu()) < 0:
            if prob1.nranks > 1:
                prob1.close()
            return zt, dict(ts=np.array(ts), its=np.array(its))
        t *= kappa
        if t > 1.0 / tol:
            raise RuntimeError("feasibility phase failed")


# --------------------------------------------------------------------------------------------------
# parabolic_solve (upstream; reference test/test_parabolic.jl:48, docs/src/guide.md:358-380): implicit
# Euler for the p-Laplace gradient flow; every step is a barrier solve on the same geometry, so the
# level plans (symbolic phase) are built once and reused for all steps - the "pattern reuse" config.
# --------------------------------------------------------------------------------------------------
PARABOLIC_STATE = (("u", "dirichlet"), ("s1", "full"), ("s2", "full"))


def parabolic_tables(dim: int):
    D = [("u", "id")] + [("u", "d" + "xyz"[k]) for k in range(dim)] + [("s1", "id"), ("s2", "id")]
    return D, [0, dim + 1], list(range(1, dim + 1)) + [dim + 2]


def parabolic_solve(geom: Geometry, h: float = 0.2, t0: float = 0.0, t1: float = 1.0, p: float = 1.0, f1=None, g=None,
                    tol: float = math.sqrt(EPS), t: float = 0.1, kappa: float = 10.0, maxit: int = 50,
                    verbose: bool = False, device: int = 0, solve_fn: Callable = solve, max_steps: Optional[int] = None,
                    **_ignored) -> ParabolicSOL:
    dim = geom.dim
    f1 = (lambda x: 0.5) if f1 is None else f1
    g = (lambda tt, x: x[0]) if g is None else g
    Dt, idxA, idxB = parabolic_tables(dim)
    M = amg_helper(geom, PARABOLIC_STATE, Dt)
    n = geom.x.shape[0]
    # cone 1 of the plan = (grad u, s2) with p, cone 2 = (u, s1) with p = 2
    rank, nranks, group = _dist_layout()
    prob = DeviceProblem(M, idxB, p, slack=False, device=device, idx2=idxA, p2=2.0, rank=rank, nranks=nranks,
                         block=geom.block, group=group)
    dev = prob.device
    ts = np.arange(t0, t1 + 1e-12 * max(1.0, abs(t1)), h)
    S = geom.subspaces["dirichlet"][-1].tocsr()
    bnd = np.flatnonzero(np.diff(S.indptr) == 0)
    f1v = np.array([f1(geom.x[i]) for i in range(n)], dtype=float)
    max_newton = int(math.ceil(math.log2(1.0 / tol) + 2))
    z = torch.zeros(3 * n, dtype=torch.float64, device=dev)
    Dz = torch.zeros((len(Dt), n), dtype=torch.float64, device=dev)

    def feasible_start(u_host):
        z.zero_()
        z[:n] = torch.from_numpy(u_host).to(dev)
        prob.apply_D(z, Dz)
        z[n:2 * n] = Dz[0] ** 2 + 1.0
        z[2 * n:] = (Dz[1:dim + 1] ** 2).sum(dim=0) ** (p / 2.0) + 1.0

    u = np.array([g(ts[0], geom.x[i]) for i in range(n)], dtype=float)
    feasible_start(u)
    snaps = [z.cpu().numpy().reshape(n, 3, order="F").copy()]
    steps_stats = []
    for k in range(len(ts) - 1):
        if max_steps is not None and k >= max_steps:
            break
        w0, before = time.perf_counter(), dict(prob.stats)
        uk = snaps[-1][:, 0]
        u0 = uk.copy()
        u0[bnd] = np.array([g(ts[k + 1], geom.x[i]) for i in bnd], dtype=float)
        feasible_start(u0)
        c = np.zeros((n, len(Dt)))
        c[:, 0] = h * f1v - uk
        c[:, dim + 1] = 0.5
        c[:, dim + 2] = h / p
        main = amgb_core(prob, z, _cm(c, dev), tol, t, kappa, maxit, max_newton, verbose, solve_fn)
        snaps.append(z.cpu().numpy().reshape(n, 3, order="F").copy())
        steps_stats.append(dict(wall_s=time.perf_counter() - w0, newton_its=int(main["its"].sum()),
                                **{k_: prob.stats.get(k_, 0) - before.get(k_, 0) for k_ in ("assemble_s", "solve_s", "plan_s", "assemblies", "f0_evals")}))
    stats = dict(steps=steps_stats, plan_s_total=prob.stats.get("plan_s", 0.0), nranks=nranks)
    if nranks > 1:
        prob.close()
    return ParabolicSOL(geom, ts[: len(snaps)], snaps, stats)
