"""Mesh + operator generators: the inputs the Newton-step assembly consumes.

These stand in for the upstream ``MultiGridBarrier.fem1d / fem2d / fem3d`` calls made by
``fem{1,2,3}d_mpi`` (reference src/MultiGridBarrierMPI.jl:559-565, 626-632, 696-702).  Upstream is
not vendored in the reference, so node ordering / quadrature are restated from its documentation
and pinned only by the sizes the reference itself pins:

* fem1d: n = 2^(L+1) broken nodes, finest Dirichlet subspace 16x7 at L=3
  (reference test/test_nonsquare.jl:28, test/test_partition_debug.jl:34);
* fem2d: n = 14*4^(L-1) (reference docs/src/guide.md:246-253) - 7-node P2+bubble triangles on the
  two-triangle square [-1,1]^2;
* fem3d: (k+1)^3 nodes per hexahedron (reference src/MultiGridBarrierMPI.jl:682-684).

Like the reference (src/MultiGridBarrierMPI.jl:239-240) every rank builds the same full native
geometry on the host; this is one-time setup, not part of the per-Newton-step path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np
import scipy.sparse as sp


@dataclass
class Geometry:
    """Mirror of ``MultiGridBarrier.Geometry`` (reference src/MultiGridBarrierMPI.jl:329-337):
    discretization tag, finest-level nodes ``x`` (n x dim), quadrature weights ``w`` (n),
    ``subspaces[key][l]`` (n_l x dofs_l), ``operators[key]`` (n x n), ``refine[l]`` / ``coarsen[l]``."""

    discretization: str
    x: np.ndarray
    w: np.ndarray
    subspaces: Dict[str, List[sp.csr_matrix]]
    operators: Dict[str, sp.csr_matrix]
    refine: List[sp.csr_matrix]
    coarsen: List[sp.csr_matrix]
    block: int = 1  # nodes per (broken) element on the finest level
    meta: dict = field(default_factory=dict)

    @property
    def L(self) -> int:
        return len(self.refine)

    @property
    def dim(self) -> int:
        return self.x.shape[1]


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------

def _block_diag_uniform(blocks: np.ndarray) -> sp.csr_matrix:
    """blocks: (E, r, c) dense -> block diagonal CSR (E*r x E*c) keeping every entry of every
    block that is non-zero in ANY element (uniform structure across elements)."""
    E, r, c = blocks.shape
    mask = np.any(blocks != 0.0, axis=0)  # (r, c) uniform structural mask
    rr, cc = np.nonzero(mask)
    order = np.lexsort((cc, rr))
    rr, cc = rr[order], cc[order]
    counts = np.bincount(rr, minlength=r)
    nnz_blk = rr.size
    rows_ptr_blk = np.concatenate([[0], np.cumsum(counts)])
    indptr = (np.arange(E)[:, None] * nnz_blk + rows_ptr_blk[None, :-1]).reshape(-1)
    indptr = np.concatenate([indptr, [E * nnz_blk]]).astype(np.int64)
    indices = (np.arange(E)[:, None] * c + cc[None, :]).reshape(-1).astype(np.int64)
    data = blocks[:, rr, cc].reshape(-1)
    return sp.csr_matrix((data, indices, indptr), shape=(E * r, E * c))


def _selection(rows: np.ndarray, cols: np.ndarray, shape) -> sp.csr_matrix:
    m = sp.csr_matrix((np.ones(rows.size), (rows, cols)), shape=shape)
    m.sum_duplicates()
    m.sort_indices()
    return m


def _unique_nodes(x: np.ndarray, scale: float):
    """Global (continuous) node ids for broken nodes by coordinate match; ids are assigned in
    lexicographic coordinate order so every rank derives the same numbering."""
    key = np.round(x * scale).astype(np.int64)
    _, inv = np.unique(key, axis=0, return_inverse=True)
    return inv.reshape(-1)


# --------------------------------------------------------------------------------------
# fem1d
# --------------------------------------------------------------------------------------

def fem1d(L: int = 4, dtype=np.float64) -> Geometry:
    """Piecewise-linear broken elements on [-1,1]; level l has 2^l elements, 2 nodes each."""
    assert L >= 1
    full, dirichlet, uniform, refine, coarsen = [], [], [], [], []
    for l in range(1, L + 1):
        ne = 2 ** l
        n_l = 2 * ne
        e = np.arange(ne)
        rows = np.arange(n_l)
        gid = np.stack([e, e + 1], axis=1).reshape(-1)  # continuous node id of each broken node
        full.append(_selection(rows, gid, (n_l, ne + 1)))
        interior = (gid > 0) & (gid < ne)
        dirichlet.append(_selection(rows[interior], gid[interior] - 1, (n_l, ne - 1)))
        uniform.append(sp.csr_matrix(np.ones((n_l, 1))))
        if l < L:
            blk = np.array([[1.0, 0.0], [0.5, 0.5], [0.5, 0.5], [0.0, 1.0]])
            refine.append(_block_diag_uniform(np.broadcast_to(blk, (ne, 4, 2)).copy()))
            # parent left <- child(2e).left ; parent right <- child(2e+1).right
            prow = np.arange(n_l)
            pcol = np.stack([4 * e, 4 * e + 3], axis=1).reshape(-1)
            coarsen.append(_selection(prow, pcol, (n_l, 4 * ne)))
        else:
            refine.append(sp.identity(n_l, format="csr"))
            coarsen.append(sp.identity(n_l, format="csr"))
    ne = 2 ** L
    h = 2.0 / ne
    e = np.arange(ne)
    x = np.stack([-1.0 + h * e, -1.0 + h * (e + 1)], axis=1).reshape(-1, 1).astype(dtype)
    w = np.full(2 * ne, h / 2.0, dtype=dtype)
    dxb = np.array([[-1.0, 1.0], [-1.0, 1.0]]) / h
    ops = {
        "id": sp.identity(2 * ne, format="csr", dtype=dtype),
        "dx": _block_diag_uniform(np.broadcast_to(dxb, (ne, 2, 2)).copy()),
    }
    return Geometry("fem1d", x, w, {"dirichlet": dirichlet, "full": full, "uniform": uniform},
                    ops, refine, coarsen, block=2, meta={"L": L})


# --------------------------------------------------------------------------------------
# fem2d : P2 + cubic bubble, 7 nodes / triangle
# --------------------------------------------------------------------------------------

_REF_NODES_2D = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0],
                          [0.5, 0.0], [0.5, 0.5], [0.0, 0.5],
                          [1.0 / 3.0, 1.0 / 3.0]])
_REF_W_2D = np.array([3.0, 3.0, 3.0, 8.0, 8.0, 8.0, 27.0]) / 60.0  # x area; exact for cubics


def _phi2d(P):
    xi, et = P[:, 0], P[:, 1]
    b = xi * et * (1.0 - xi - et)
    return np.stack([np.ones_like(xi), xi, et, xi * xi, xi * et, et * et, b], axis=1)


def _dphi2d(P):
    xi, et = P[:, 0], P[:, 1]
    z, o = np.zeros_like(xi), np.ones_like(xi)
    dxi = np.stack([z, o, z, 2 * xi, et, z, et * (1 - xi - et) - xi * et], axis=1)
    det = np.stack([z, z, o, z, xi, 2 * et, xi * (1 - xi - et) - xi * et], axis=1)
    return dxi, det


_C2D = np.linalg.inv(_phi2d(_REF_NODES_2D))  # nodal basis coefficients


def _ref_tab_2d():
    N = _phi2d(_REF_NODES_2D) @ _C2D
    dxi, det = _dphi2d(_REF_NODES_2D)
    return N, dxi @ _C2D, det @ _C2D


# children in parent reference coordinates (v1,m12,m31),(m12,v2,m23),(m31,m23,v3),(m12,m23,m31)
_CHILD_VERTS_2D = np.array([[0, 3, 5], [3, 1, 4], [5, 4, 2], [3, 4, 5]])


def _child_interp_2d():
    blocks = []
    for ch in range(4):
        v = _REF_NODES_2D[_CHILD_VERTS_2D[ch]]  # (3,2) child vertices in parent ref coords
        a, b, c = v
        P = a[None, :] + np.outer(_REF_NODES_2D[:, 0], b - a) + np.outer(_REF_NODES_2D[:, 1], c - a)
        blk = _phi2d(P) @ _C2D
        blk[np.abs(blk) < 1e-14] = 0.0
        blocks.append(blk)
    return np.concatenate(blocks, axis=0)  # (28, 7)


def _tri_nodes(verts):
    """verts: (T,3,2) -> (T,7,2) element nodes."""
    a, b, c = verts[:, 0], verts[:, 1], verts[:, 2]
    return np.stack([a, b, c, (a + b) / 2, (b + c) / 2, (c + a) / 2, (a + b + c) / 3], axis=1)


def _subdivide(verts):
    nodes = _tri_nodes(verts)  # (T,7,2)
    ch = nodes[:, _CHILD_VERTS_2D, :]  # (T,4,3,2)
    return ch.reshape(-1, 3, 2)


def _subspaces_2d(verts, scale):
    T = verts.shape[0]
    nodes = _tri_nodes(verts).reshape(-1, 2)
    gid = _unique_nodes(nodes, scale)
    ndof = int(gid.max()) + 1
    n_l = 7 * T
    rows = np.arange(n_l)
    full = _selection(rows, gid, (n_l, ndof))
    # boundary edges: edges (by midpoint id) seen by exactly one triangle
    gid_e = gid.reshape(T, 7)
    mids = gid_e[:, 3:6].reshape(-1)
    cnt = np.bincount(mids, minlength=ndof)
    bnd_edge = (cnt[gid_e[:, 3:6]] == 1)  # (T,3): edges 12, 23, 31
    is_b = np.zeros(ndof, dtype=bool)
    ends = np.array([[0, 1], [1, 2], [2, 0]])
    for k in range(3):
        sel = bnd_edge[:, k]
        is_b[gid_e[sel, 3 + k]] = True
        is_b[gid_e[sel, ends[k, 0]]] = True
        is_b[gid_e[sel, ends[k, 1]]] = True
    newid = np.cumsum(~is_b) - 1
    keep = ~is_b[gid]
    dirichlet = _selection(rows[keep], newid[gid[keep]], (n_l, int((~is_b).sum())))
    uniform = sp.csr_matrix(np.ones((n_l, 1)))
    return full, dirichlet, uniform


def fem2d(L: int = 2, K=None, dtype=np.float64) -> Geometry:
    """K: (3T,2) triangle vertices of the coarse mesh (default: the square [-1,1]^2 split in two,
    the same coordinates as upstream's documented default)."""
    assert L >= 1
    if K is None:
        K = np.array([[-1, -1], [1, -1], [-1, 1], [1, -1], [1, 1], [-1, 1]], dtype=float)
    K = np.asarray(K, dtype=float)
    verts = K.reshape(-1, 3, 2)
    span = float(np.max(np.abs(K))) or 1.0
    full, dirichlet, uniform, refine, coarsen = [], [], [], [], []
    interp = _child_interp_2d()
    for l in range(1, L + 1):
        T = verts.shape[0]
        scale = 3.0 * (2.0 ** (l + 2)) / span * 64.0
        f, d, u = _subspaces_2d(verts, scale)
        full.append(f), dirichlet.append(d), uniform.append(u)
        if l < L:
            refine.append(_block_diag_uniform(np.broadcast_to(interp, (T, 28, 7)).copy()))
            # parent node <- coincident child node (child-major numbering 4t+ch, 7 nodes each)
            src = np.array([0 * 7 + 0, 1 * 7 + 1, 2 * 7 + 2, 0 * 7 + 1, 1 * 7 + 2, 0 * 7 + 2, 3 * 7 + 6])
            prow = np.arange(7 * T)
            pcol = (np.arange(T)[:, None] * 28 + src[None, :]).reshape(-1)
            coarsen.append(_selection(prow, pcol, (7 * T, 28 * T)))
            verts = _subdivide(verts)
        else:
            refine.append(sp.identity(7 * T, format="csr"))
            coarsen.append(sp.identity(7 * T, format="csr"))
    T = verts.shape[0]
    x = _tri_nodes(verts).reshape(-1, 2).astype(dtype)
    a, b, c = verts[:, 0], verts[:, 1], verts[:, 2]
    J = np.stack([b - a, c - a], axis=2)  # (T,2,2) columns = edge vectors
    detJ = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
    Jinv = np.empty_like(J)
    Jinv[:, 0, 0] = J[:, 1, 1] / detJ
    Jinv[:, 0, 1] = -J[:, 0, 1] / detJ
    Jinv[:, 1, 0] = -J[:, 1, 0] / detJ
    Jinv[:, 1, 1] = J[:, 0, 0] / detJ
    w = (np.abs(detJ)[:, None] / 2.0 * _REF_W_2D[None, :]).reshape(-1).astype(dtype)
    _, Dxi, Det = _ref_tab_2d()
    Dxi = np.where(np.abs(Dxi) < 1e-13, 0.0, Dxi)
    Det = np.where(np.abs(Det) < 1e-13, 0.0, Det)
    dxb = Jinv[:, 0, 0, None, None] * Dxi[None] + Jinv[:, 1, 0, None, None] * Det[None]
    dyb = Jinv[:, 0, 1, None, None] * Dxi[None] + Jinv[:, 1, 1, None, None] * Det[None]
    ops = {
        "id": sp.identity(7 * T, format="csr", dtype=dtype),
        "dx": _block_diag_uniform(dxb),
        "dy": _block_diag_uniform(dyb),
    }
    return Geometry("fem2d", x, w, {"dirichlet": dirichlet, "full": full, "uniform": uniform},
                    ops, refine, coarsen, block=7, meta={"L": L})


# --------------------------------------------------------------------------------------
# fem3d : Q_k hexahedra, (k+1)^3 nodes / element, Gauss-Lobatto-Legendre nodes
# --------------------------------------------------------------------------------------

def _gll(k: int):
    """k+1 Gauss-Lobatto-Legendre nodes/weights on [-1,1]."""
    if k == 1:
        return np.array([-1.0, 1.0]), np.array([1.0, 1.0])
    from numpy.polynomial import legendre as leg
    Pk = leg.Legendre.basis(k)
    xi = np.concatenate([[-1.0], np.sort(Pk.deriv().roots().real), [1.0]])
    w = 2.0 / (k * (k + 1) * Pk(xi) ** 2)
    return xi, w


def _lagrange_tab(nodes: np.ndarray, pts: np.ndarray):
    """values and derivatives of the Lagrange basis on ``nodes`` at ``pts`` -> (len(pts), len(nodes))."""
    m = nodes.size
    V = np.vander(nodes, m, increasing=True)
    C = np.linalg.inv(V)
    P = np.vander(pts, m, increasing=True)
    dP = np.zeros_like(P)
    dP[:, 1:] = P[:, :-1] * np.arange(1, m)[None, :]
    return P @ C, dP @ C


def fem3d(L: int = 2, k: int = 3, dtype=np.float64) -> Geometry:
    """Unit-cube [0,1]^3 (reference src/MultiGridBarrierMPI.jl:684) with 8^(l-1) Q_k hexahedra at
    level l; nodes are tensor GLL points, x fastest."""
    assert L >= 1 and k >= 1
    m = k + 1
    b = m ** 3
    xi, wq = _gll(k)
    xi01, w01 = (xi + 1.0) / 2.0, wq / 2.0
    Nid, D1 = _lagrange_tab(xi01, xi01)  # derivative w.r.t. the [0,1] reference coordinate
    D1 = np.where(np.abs(D1) < 1e-12, 0.0, D1)
    I1 = np.eye(m)
    # child interpolation along one axis: child c in {0,1} covers [c/2,(c+1)/2]
    A = [np.where(np.abs(t) < 1e-13, 0.0, t) for t in
         (_lagrange_tab(xi01, xi01 / 2.0)[0], _lagrange_tab(xi01, 0.5 + xi01 / 2.0)[0])]

    def elem_origins(l):
        ne1 = 2 ** (l - 1)
        if l == 1:
            return np.zeros((1, 3), dtype=np.int64), ne1
        prev, _ = elem_origins(l - 1)
        ch = np.array([[cx, cy, cz] for cz in (0, 1) for cy in (0, 1) for cx in (0, 1)])
        return (2 * prev[:, None, :] + ch[None, :, :]).reshape(-1, 3), ne1

    full, dirichlet, uniform, refine, coarsen = [], [], [], [], []
    for l in range(1, L + 1):
        org, ne1 = elem_origins(l)
        E = org.shape[0]
        h = 1.0 / ne1
        loc = np.stack(np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij"), axis=-1)
        loc = loc.transpose(2, 1, 0, 3).reshape(-1, 3)  # x fastest
        # integer lattice position of each node (exact match across elements)
        pos1 = np.round(xi01 * 10 ** 9).astype(np.int64)
        lat = org[:, None, :] * 10 ** 9 + pos1[loc][None, :, :]
        lat = lat.reshape(-1, 3)
        _, gid = np.unique(lat, axis=0, return_inverse=True)
        gid = gid.reshape(-1)
        ndof = int(gid.max()) + 1
        n_l = E * b
        rows = np.arange(n_l)
        full.append(_selection(rows, gid, (n_l, ndof)))
        onb = np.any((lat == 0) | (lat == ne1 * 10 ** 9), axis=1)
        is_b = np.zeros(ndof, dtype=bool)
        is_b[gid[onb]] = True
        newid = np.cumsum(~is_b) - 1
        keep = ~is_b[gid]
        dirichlet.append(_selection(rows[keep], newid[gid[keep]], (n_l, int((~is_b).sum()))))
        uniform.append(sp.csr_matrix(np.ones((n_l, 1))))
        if l < L:
            blocks = []
            for cz in (0, 1):
                for cy in (0, 1):
                    for cx in (0, 1):
                        blocks.append(np.kron(A[cz], np.kron(A[cy], A[cx])))
            blk = np.concatenate(blocks, axis=0)  # (8b, b)
            refine.append(_block_diag_uniform(np.broadcast_to(blk, (E, 8 * b, b)).copy()))
            # coarsen: parent node <- coincident child node when one exists (left inverse of refine)
            rsel = np.full(b, -1)
            for r in range(8 * b):
                row = blk[r]
                j = np.flatnonzero(row)
                if j.size == 1 and abs(row[j[0]] - 1.0) < 1e-12 and rsel[j[0]] < 0:
                    rsel[j[0]] = r
            if np.all(rsel >= 0):
                prow = np.arange(E * b)
                pcol = (np.arange(E)[:, None] * 8 * b + rsel[None, :]).reshape(-1)
                coarsen.append(_selection(prow, pcol, (E * b, 8 * E * b)))
            else:  # GLL nodes of order k>=3 are not nested: use the least-squares left inverse
                pinv = np.linalg.pinv(blk)
                pinv[np.abs(pinv) < 1e-13] = 0.0
                coarsen.append(_block_diag_uniform(np.broadcast_to(pinv, (E, b, 8 * b)).copy()))
        else:
            refine.append(sp.identity(n_l, format="csr"))
            coarsen.append(sp.identity(n_l, format="csr"))
    org, ne1 = elem_origins(L)
    E = org.shape[0]
    h = 1.0 / ne1
    loc = np.stack(np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij"), axis=-1)
    loc = loc.transpose(2, 1, 0, 3).reshape(-1, 3)
    x = ((org[:, None, :] + xi01[loc][None, :, :]) * h).reshape(-1, 3).astype(dtype)
    w1 = w01 * h
    wloc = w1[loc[:, 0]] * w1[loc[:, 1]] * w1[loc[:, 2]]
    w = np.tile(wloc, E).astype(dtype)
    Dh = D1 / h
    dxb = np.kron(I1, np.kron(I1, Dh))
    dyb = np.kron(I1, np.kron(Dh, I1))
    dzb = np.kron(Dh, np.kron(I1, I1))
    ops = {
        "id": sp.identity(E * b, format="csr", dtype=dtype),
        "dx": _block_diag_uniform(np.broadcast_to(dxb, (E, b, b)).copy()),
        "dy": _block_diag_uniform(np.broadcast_to(dyb, (E, b, b)).copy()),
        "dz": _block_diag_uniform(np.broadcast_to(dzb, (E, b, b)).copy()),
    }
    return Geometry("fem3d", x, w, {"dirichlet": dirichlet, "full": full, "uniform": uniform},
                    ops, refine, coarsen, block=b, meta={"L": L, "k": k})
