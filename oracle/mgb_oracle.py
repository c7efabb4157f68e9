"""CPU ORACLE - TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module.  The product (``multigridbarriermpi.jl_b200``) never does.

PARITY UNPINNED.  The arithmetic of the Newton-step assembly lives in two Julia packages that are
NOT vendored under /root/reference and whose pinned versions are unknown (no Manifest.toml):
``MultiGridBarrier`` (compat "0.11") and ``HPCSparseArrays`` (compat "0.1")
(reference Project.toml:11,15,29,33).  Neither Julia nor MPI exists in this image, so the reference
cannot be executed.  This file restates the algorithm from the reference's own call sites and from
the in-repo restatements of the upstream bodies; every function cites the file:line it follows.
The only absolute known-answer vectors the reference holds for this path are the three
``map_rows`` literals (test/test_helpers.jl:123-167), the literal sparse product
(test/test_basic_ops.jl:27-51) and the structural sizes (docs/src/guide.md:246-253,
test/test_nonsquare.jl:28); ``tests/test_oracle.py`` pins the oracle against all of them.

Numerics: float64 throughout; sparse algebra in CSC with the same product association and loop
order as the reference's f2 loop (test/test_map_rows_compare.jl:102-123); sparse ``+`` drops exact
zeros like Julia's ``SparseMatrixCSC`` ``+`` (the behaviour test/test_matrix_addition.jl:22-24
documents).
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

EPS = np.finfo(np.float64).eps
# Newton stops once the decrement is within the rounding noise of the objective (see solver.py)
NEWTON_NOISE = 1024.0


# --------------------------------------------------------------------------------------
# map_rows  (reference src/MultiGridBarrierMPI.jl:161-170; algorithm restated in
# tools/profile_map_rows_steps.jl:55-150 and tools/profile_comprehension.jl:50-54)
# --------------------------------------------------------------------------------------

def map_rows(f: Callable, *args):
    """Apply ``f`` to aligned rows of every argument.  Matrix arguments contribute their row
    (1-D view), vector arguments a length-1 view (tools/profile_local_rows.jl:54-58).  Scalar
    results stack into a vector; 1 x k rows / length-k vectors stack into an n x k matrix
    (test/test_helpers.jl:123-167, test/test_apply_d.jl:81)."""
    n = args[0].shape[0]
    rows = []
    for i in range(n):
        ri = [a[i] if a.ndim == 2 else a[i:i + 1] for a in args]
        rows.append(f(*ri))
    first = rows[0]
    if np.ndim(first) == 0:
        return np.array(rows, dtype=float)
    return np.vstack([np.asarray(r, dtype=float).reshape(1, -1) for r in rows])


# --------------------------------------------------------------------------------------
# Convex sets / pointwise barriers  (upstream MultiGridBarrier ``convex_Euclidian_power`` and
# ``convex_linear``; shapes evidenced by test/test_map_rows_compare.jl:62-73: F2 is an nD x nD
# symmetric block read at column (j-1)*nD+k)
# --------------------------------------------------------------------------------------

def _mu(p: float) -> float:
    """weight of the extra -log(s) term of the power-cone barrier.  [U]: upstream ``convex_Euclidian_power`` is not
    under /root/reference.  p = 1 (-log(s^2 - |q|^2), Lorentz cone) and p = 2 (-log(s - |q|^2), paraboloid epigraph)
    are self-concordant without it -> 0; 1 < p < 2 -> 1; p > 2 -> 2.  (Round 1 used 2 at p = 2; changed on review:
    the p = 2 set s - |q|^2 > 0 already implies s > 0.)"""
    return 0.0 if (p == 1.0 or p == 2.0) else (1.0 if p < 2.0 else 2.0)


@dataclass
class EuclidianPower:
    """{(q,s): s >= |q|^p}; F = -log(s^(2/p) - |q|^2) - mu(p) log s on y[idx] = (q..., s).
    ``idx`` is 0-based into the Dz row.  ``slack`` (feasibility phase) adds the last Dz column tau to
    s and bounds it below with the extra term -log(1 + tau) (keeps the phase-1 Hessian nonsingular)."""
    idx: Sequence[int]
    p: float
    slack: bool = False

    def _split(self, y):
        yy = y[..., list(self.idx)]
        q, s = yy[..., :-1], yy[..., -1]
        if self.slack:
            s = s + y[..., -1]
        return q, s

    def F(self, x, y):
        q, s = self._split(np.asarray(y, dtype=float))
        a = 2.0 / self.p
        with np.errstate(all="ignore"):
            phi = np.where(s > 0, np.abs(s) ** a, -np.inf) - np.sum(q * q, axis=-1)
            out = -np.log(phi) - _mu(self.p) * np.log(s)
            ok = (phi > 0) & (s > 0)
            if self.slack:
                tau1 = 1.0 + np.asarray(y, dtype=float)[..., -1]
                out = out - np.log(tau1)
                ok = ok & (tau1 > 0)
        return np.where(ok, out, np.inf)

    def F1(self, x, y):
        y = np.asarray(y, dtype=float)
        q, s = self._split(y)
        a = 2.0 / self.p
        phi = s ** a - np.sum(q * q, axis=-1)
        g = np.zeros_like(y)
        idx = list(self.idx)
        gs = -a * s ** (a - 1.0) / phi - _mu(self.p) / s
        for m, k in enumerate(idx[:-1]):
            g[..., k] = 2.0 * q[..., m] / phi
        g[..., idx[-1]] = gs
        if self.slack:
            g[..., -1] = gs - 1.0 / (1.0 + y[..., -1])
        return g

    def F2(self, x, y):
        """Returns (..., nD, nD) symmetric blocks."""
        y = np.asarray(y, dtype=float)
        q, s = self._split(y)
        a = 2.0 / self.p
        nD = y.shape[-1]
        phi = s ** a - np.sum(q * q, axis=-1)
        H = np.zeros(y.shape[:-1] + (nD, nD))
        idx = list(self.idx)
        ds = a * s ** (a - 1.0)
        hss = -a * (a - 1.0) * s ** (a - 2.0) / phi + ds * ds / (phi * phi) + _mu(self.p) / (s * s)
        svars = [idx[-1]] + ([nD - 1] if self.slack else [])
        for m, k in enumerate(idx[:-1]):
            for m2, k2 in enumerate(idx[:-1]):
                H[..., k, k2] = 4.0 * q[..., m] * q[..., m2] / (phi * phi) + (2.0 / phi if m == m2 else 0.0)
            for sv in svars:
                H[..., k, sv] = H[..., sv, k] = -2.0 * q[..., m] * ds / (phi * phi)
        for sv in svars:
            for sv2 in svars:
                H[..., sv, sv2] = hss
        if self.slack:
            H[..., -1, -1] = hss + 1.0 / (1.0 + y[..., -1]) ** 2
        return H


@dataclass
class Intersection:
    """Intersection of convex sets: the barriers add (upstream ``intersect`` / the parabolic problem's
    {s1 >= u^2} with {s2 >= |grad u|^p})."""
    sets: Sequence

    def F(self, x, y):
        return sum(Q.F(x, y) for Q in self.sets)

    def F1(self, x, y):
        return sum(Q.F1(x, y) for Q in self.sets)

    def F2(self, x, y):
        return sum(Q.F2(x, y) for Q in self.sets)


# --------------------------------------------------------------------------------------
# f0 / f1 / f2 : the Newton-step assembly.  Loop order follows the reference's restatement of
# the upstream ``barrier`` body: test/test_map_rows_compare.jl:102-123 (Hessian),
# test/test_apply_d.jl:44 (apply_D), test/test_diag.jl:80-83 + test/test_nonsquare.jl:57-72
# (gradient ops), test/test_map_rows_compare.jl:165-171 (restriction R' * H * R).
# --------------------------------------------------------------------------------------

def apply_D(D: Sequence[sp.spmatrix], z: np.ndarray) -> np.ndarray:
    """hcat([D[k]*z for k]...)  (test/test_apply_d.jl:44)."""
    return np.stack([Dk @ z for Dk in D], axis=1)


def _add_dropzeros(A, B):
    C = (A + B).tocsc()
    C.eliminate_zeros()  # Julia's sparse + drops exact zeros (test/test_matrix_addition.jl:22-24)
    return C


def amgb_diag(z: np.ndarray) -> sp.csc_matrix:
    """spdiagm(m, n, 0 => z) keeping explicit zeros (reference src/MultiGridBarrierMPI.jl:137-147)."""
    n = z.shape[0]
    return sp.csc_matrix((z.copy(), np.arange(n), np.arange(n + 1)), shape=(n, n))


def f0(s, x, w, c, R, D, z0, Q) -> float:
    Dz = apply_D(D, z0 + R @ s)
    y = Q.F(x, Dz)
    if not np.all(np.isfinite(y)):
        return math.inf
    return float(np.dot(w, y) + sum(np.dot(w * c[:, k], Dz[:, k]) for k in range(len(D))))


def f1(s, x, w, c, R, D, z0, Q) -> np.ndarray:
    Dz = apply_D(D, z0 + R @ s)
    y = Q.F1(x, Dz) + c
    ret = np.zeros(D[0].shape[1])
    for k in range(len(D)):
        ret = ret + D[k].T @ (w * y[:, k])
    return R.T @ ret


def hessian_fine(y2, w, D) -> sp.csc_matrix:
    """ret = sum_j D_j' diag(w.*y_jj) D_j + sum_{k<j} (D_j' diag D_k + D_k' diag D_j) with y2 the
    n x nD^2 map_rows output read at column (j-1)*nD+k (test/test_map_rows_compare.jl:102-123)."""
    nD = len(D)
    Dc = [d.tocsc() for d in D]
    Dt = [d.T.tocsc() for d in D]
    ret = None
    for j in range(nD):
        foo = amgb_diag(w * y2[:, j * nD + j])
        bar = (Dt[j] @ foo) @ Dc[j]
        ret = bar if ret is None else _add_dropzeros(ret, bar)
        for k in range(j):
            foo = amgb_diag(w * y2[:, j * nD + k])
            t1 = (Dt[j] @ foo) @ Dc[k]
            t2 = (Dt[k] @ foo) @ Dc[j]
            ret = _add_dropzeros(_add_dropzeros(ret, t1), t2)
    return ret


def f2(s, x, w, c, R, D, z0, Q) -> sp.csc_matrix:
    Dz = apply_D(D, z0 + R @ s)
    nD = len(D)
    y2 = Q.F2(x, Dz).reshape(Dz.shape[0], nD * nD)
    ret = hessian_fine(y2, w, D)
    Rc = R.tocsc()
    return ((Rc.T.tocsc() @ ret) @ Rc).tocsc()


# --------------------------------------------------------------------------------------
# amg_helper  (upstream; restated in test/test_d0_construction.jl:81-100 and
# test/test_amg_structure.jl:42-58: D0[l,k] = hcat(Z.., op, ..Z), R = blockdiag(R_u, R_s))
# --------------------------------------------------------------------------------------

@dataclass
class AMG:
    x: np.ndarray
    w: np.ndarray
    R_fine: List[sp.csr_matrix]
    D: List[sp.csr_matrix]  # finest-level operators, each n x (nu*n)
    nu: int
    nD: int
    op_var: List[int]


def amg_helper(geom, state_variables, D_table) -> AMG:
    L = len(geom.refine)
    n = geom.x.shape[0]
    refine_fine = [None] * L
    refine_fine[L - 1] = geom.refine[L - 1].tocsr()
    for l in range(L - 2, -1, -1):
        refine_fine[l] = (refine_fine[l + 1] @ geom.refine[l]).tocsr()
    nu = len(state_variables)
    R_fine = []
    for l in range(L):
        blocks = [(refine_fine[l] @ geom.subspaces[sv[1]][l]).tocsr() for sv in state_variables]
        R_fine.append(sp.block_diag(blocks, format="csr"))
    var_of = {sv[0]: k for k, sv in enumerate(state_variables)}
    Dm, op_var = [], []
    for (var, opname) in D_table:
        blocks = [sp.csr_matrix((n, n)) for _ in range(nu)]
        blocks[var_of[var]] = geom.operators[opname].tocsr()
        Dm.append(sp.hstack(blocks, format="csr"))
        op_var.append(var_of[var])
    return AMG(geom.x, geom.w, R_fine, Dm, nu, len(D_table), op_var)


# --------------------------------------------------------------------------------------
# Damped Newton + central path (upstream ``newton`` / ``amgb``; bodies NOT in the reference -
# restated from the published barrier method; only the solve seam H \ g is evidenced,
# test/test_newton_matrix_compare.jl:33-51).
# --------------------------------------------------------------------------------------

def solve(H: sp.spmatrix, g: np.ndarray) -> np.ndarray:
    """MultiGridBarrier.solve(A, b) = A \\ b (test/test_newton_matrix_compare.jl:51)."""
    return spla.splu(sp.csc_matrix(H)).solve(g)


def newton(F0, F1, F2, x, maxit=50, alpha=0.1, beta=0.25, solve_fn=solve):
    y = F0(x)
    assert math.isfinite(y), "newton: infeasible start"
    g = F1(x)
    k, converged = 0, False
    while k < maxit:
        H = F2(x)
        n = solve_fn(H, g)
        inc = float(np.dot(g, n))
        if not math.isfinite(inc) or inc <= NEWTON_NOISE * EPS * max(1.0, abs(y)):
            converged = True
            break
        k += 1
        s, ok = 1.0, False
        while s > 1e-12:
            xn = x - s * n
            yn = F0(xn)
            if math.isfinite(yn) and yn <= y - alpha * s * inc:
                ok = True
                break
            s *= beta
        if not ok:
            converged = True  # stagnated at rounding level
            k -= 1
            break
        x, y = xn, yn
        g = F1(x)
    return dict(x=x, y=y, k=k, converged=converged)


@dataclass
class AMGBSOL:
    z: np.ndarray
    SOL_feasibility: Optional[dict]
    SOL_main: dict
    log: str
    geometry: object


def amgb_core(M: AMG, Q, z, c, tol, t0, kappa, maxit_newton, max_newton_fine, verbose=False,
              solve_fn=solve, hook=None):
    """Central path: t <- kappa t from t0 until t > 1/tol.  First t: coarse-to-fine sweep of damped
    Newton solves over R_fine[1..L]; later t: Newton on the finest level, falling back to a sweep."""
    L = len(M.R_fine)
    x, w, D = M.x, M.w, M.D
    t = t0
    ts, its, cdots = [], [], []
    t_begin = time.time()
    kk = 0
    while t <= 1.0 / tol:
        kk += 1
        ts.append(t)
        it_row = [0] * L

        def level(J, maxit):
            nonlocal z
            R = M.R_fine[J]
            ct = t * c
            s0 = np.zeros(R.shape[1])
            z0 = z
            sol = newton(lambda s: f0(s, x, w, ct, R, D, z0, Q),
                         lambda s: f1(s, x, w, ct, R, D, z0, Q),
                         lambda s: f2(s, x, w, ct, R, D, z0, Q), s0, maxit=maxit, solve_fn=solve_fn)
            it_row[J] += sol["k"]
            z = z0 + R @ sol["x"]
            if hook is not None:
                hook(t, J, sol)
            return sol["converged"]

        ok = False
        if kk > 1:
            ok = level(L - 1, max_newton_fine)
        if not ok:
            for J in range(L):
                ok = level(J, maxit_newton)
        its.append(it_row)
        Dz = apply_D(D, z)
        cdots.append(float(sum(np.dot(w * c[:, k], Dz[:, k]) for k in range(len(D)))))
        if verbose:
            print(f"t={t:.3e} its={it_row} c.Dz={cdots[-1]:.12e}")
        t *= kappa
    return z, dict(ts=np.array(ts), its=np.array(its).T, c_dot_Dz=np.array(cdots),
                   t_elapsed=time.time() - t_begin)


DEFAULT_D = {1: [("u", "id"), ("u", "dx"), ("s", "id")],
             2: [("u", "id"), ("u", "dx"), ("u", "dy"), ("s", "id")],
             3: [("u", "id"), ("u", "dx"), ("u", "dy"), ("u", "dz"), ("s", "id")]}
DEFAULT_F = {1: lambda x: [0.5, 0.0, 1.0], 2: lambda x: [0.5, 0.0, 0.0, 1.0],
             3: lambda x: [0.5, 0.0, 0.0, 0.0, 1.0]}  # 3D: reference src/MultiGridBarrierMPI.jl:737
DEFAULT_G = {1: lambda x: [x[0], 2.0], 2: lambda x: [x[0] ** 2 + x[1] ** 2, 100.0],
             3: lambda x: [x[0] ** 2 + x[1] ** 2 + x[2] ** 2, 100.0]}  # 3D: src:738


def amgb(geom, p=1.0, tol=math.sqrt(EPS), t0=0.1, kappa=10.0, maxit=50, max_newton=None,
         state_variables=(("u", "dirichlet"), ("s", "full")), D=None, f=None, g=None, Q=None,
         verbose=False, solve_fn=solve, hook=None) -> AMGBSOL:
    dim = geom.x.shape[1]
    D = DEFAULT_D[dim] if D is None else D
    f = DEFAULT_F[dim] if f is None else f
    g = DEFAULT_G[dim] if g is None else g
    M = amg_helper(geom, state_variables, D)
    if Q is None:
        Q = EuclidianPower(idx=list(range(1, dim + 2)), p=float(p))
    n = geom.x.shape[0]
    z0 = np.array([g(geom.x[i]) for i in range(n)], dtype=float)  # n x nu
    c = np.array([f(geom.x[i]) for i in range(n)], dtype=float)  # n x nD
    z = z0.reshape(-1, order="F").copy()
    if max_newton is None:
        max_newton = int(math.ceil((math.log2(1.0 / tol)) + 2))
    Dz = apply_D(M.D, z)
    sol_feas = None
    if not np.all(np.isfinite(Q.F(geom.x, Dz))):
        z, sol_feas = feasibility_phase(geom, M, Q, z, c, state_variables, D, tol, t0, kappa, maxit,
                                        max_newton, solve_fn)
    z, sol_main = amgb_core(M, Q, z, c, tol, t0, kappa, maxit, max_newton, verbose, solve_fn, hook)
    return AMGBSOL(z.reshape(n, M.nu, order="F"), sol_feas, sol_main, "", geom)


SLACK_COST = 10.0


def feasibility_phase(geom, M, Q, z, c, state_variables, D, tol, t0, kappa, maxit, max_newton, solve_fn):
    """Phase 1: add a slack state variable tau (:feasibility_slack, :full) with operator :id; follow the
    central path of  c.Dz + SLACK_COST*tau  s.t. (q, s + tau) in Q, tau > -1  until tau < 0 everywhere,
    i.e. the original constraints hold strictly."""
    n = geom.x.shape[0]
    sv1 = tuple(state_variables) + (("feasibility_slack", "full"),)
    D1 = list(D) + [("feasibility_slack", "id")]
    M1 = amg_helper(geom, sv1, D1)
    Q1 = EuclidianPower(idx=Q.idx, p=Q.p, slack=True)
    Dz = apply_D(M.D, z)
    q, s = Q._split(Dz)
    need = np.sum(q * q, axis=-1) ** (Q.p / 2.0) - s
    slack0 = max(1.0, 2.0 * float(np.max(need)) + 1.0)
    z1 = np.concatenate([z, np.full(n, slack0)])
    c1 = np.hstack([c, np.full((n, 1), SLACK_COST)])
    t = t0
    ts, its = [], []
    while True:
        R = M1.R_fine[-1]
        ct = t * c1
        z0 = z1
        sol = newton(lambda s_: f0(s_, geom.x, geom.w, ct, R, M1.D, z0, Q1),
                     lambda s_: f1(s_, geom.x, geom.w, ct, R, M1.D, z0, Q1),
                     lambda s_: f2(s_, geom.x, geom.w, ct, R, M1.D, z0, Q1),
                     np.zeros(R.shape[1]), maxit=maxit, solve_fn=solve_fn)
        z1 = z0 + R @ sol["x"]
        ts.append(t), its.append(sol["k"])
        zt = z1[:-n]
        if np.all(np.isfinite(Q.F(geom.x, apply_D(M.D, zt)))) and np.max(z1[-n:]) < 0:
            return zt, dict(ts=np.array(ts), its=np.array(its))
        t *= kappa
        if t > 1.0 / tol:
            raise RuntimeError("feasibility phase failed")


# --------------------------------------------------------------------------------------
# parabolic_solve (upstream; reached through test/test_parabolic.jl:48 and docs/src/guide.md:358-380):
# implicit Euler for the p-Laplace gradient flow  u_t - div(|grad u|^(p-2) grad u) = -f1.
# Each step minimises  int (1/2) s1 + (h/p) s2 + (h f1 - u_k) u   s.t.  s1 >= u^2, s2 >= |grad u|^p
# with the barrier method on the same geometry (identical sparsity every step).  Body not in the
# reference: restated.  The start of every step is made strictly feasible by construction
# (s1 = u^2 + 1, s2 = |grad u|^p + 1), so no feasibility phase is needed.
# --------------------------------------------------------------------------------------

PARABOLIC_STATE = (("u", "dirichlet"), ("s1", "full"), ("s2", "full"))


def parabolic_tables(dim):
    D = [("u", "id")] + [("u", "d" + "xyz"[k]) for k in range(dim)] + [("s1", "id"), ("s2", "id")]
    idxA = [0, dim + 1]                              # (u, s1), p = 2
    idxB = list(range(1, dim + 1)) + [dim + 2]       # (grad u, s2), p
    return D, idxA, idxB


def boundary_mask(geom):
    """broken nodes whose Dirichlet row is empty = boundary nodes"""
    S = geom.subspaces["dirichlet"][-1].tocsr()
    return np.diff(S.indptr) == 0


@dataclass
class ParabolicSOL:
    geometry: object
    ts: np.ndarray
    u: list


def parabolic_feasible_start(M, u, dim, p):
    n = M.x.shape[0]
    z = np.concatenate([u, np.zeros(2 * n)])
    Dz = apply_D(M.D, z)
    s1 = Dz[:, 0] ** 2 + 1.0
    s2 = np.sum(Dz[:, 1:dim + 1] ** 2, axis=1) ** (p / 2.0) + 1.0
    return np.concatenate([u, s1, s2])


def parabolic_solve(geom, h=0.2, t0=0.0, t1=1.0, p=1.0, f1=None, g=None, tol=math.sqrt(EPS), t=0.1, kappa=10.0,
                    maxit=50, verbose=False, solve_fn=solve):
    dim = geom.x.shape[1]
    f1 = (lambda x: 0.5) if f1 is None else f1
    g = (lambda tt, x: x[0]) if g is None else g
    Dt, idxA, idxB = parabolic_tables(dim)
    M = amg_helper(geom, PARABOLIC_STATE, Dt)
    Q = Intersection([EuclidianPower(idx=idxA, p=2.0), EuclidianPower(idx=idxB, p=float(p))])
    n = geom.x.shape[0]
    ts = np.arange(t0, t1 + 1e-12 * max(1.0, abs(t1)), h)
    bnd = boundary_mask(geom)
    f1v = np.array([f1(geom.x[i]) for i in range(n)], dtype=float)
    u = np.array([g(ts[0], geom.x[i]) for i in range(n)], dtype=float)
    z = parabolic_feasible_start(M, u, dim, p)
    snaps = [z.reshape(n, 3, order="F").copy()]
    max_newton = int(math.ceil(math.log2(1.0 / tol) + 2))
    for k in range(len(ts) - 1):
        uk = snaps[-1][:, 0]
        u0 = uk.copy()
        u0[bnd] = np.array([g(ts[k + 1], geom.x[i]) for i in np.flatnonzero(bnd)], dtype=float)
        z = parabolic_feasible_start(M, u0, dim, p)
        c = np.zeros((n, len(Dt)))
        c[:, 0] = h * f1v - uk
        c[:, dim + 1] = 0.5
        c[:, dim + 2] = h / p
        z, _ = amgb_core(M, Q, z, c, tol, t, kappa, maxit, max_newton, verbose, solve_fn)
        snaps.append(z.reshape(n, 3, order="F").copy())
    return ParabolicSOL(geom, ts, snaps)
