"""Level operators (upstream ``amg_helper``): once-per-solve setup on the host.

Restates reference test/test_d0_construction.jl:81-100 (D0[l,k] = hcat(Z.., op, ..Z),
R = blockdiag(...)) and test/test_amg_structure.jl:42-58 (refine/coarsen chains).  As in the
reference every rank builds these from the replicated native geometry
(src/MultiGridBarrierMPI.jl:239-240); the per-Newton-step work happens on the GPU through the plans
built from them (``capi.Plan``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

from .geometry import Geometry

DEFAULT_STATE = (("u", "dirichlet"), ("s", "full"))
DEFAULT_D = {1: [("u", "id"), ("u", "dx"), ("s", "id")],
             2: [("u", "id"), ("u", "dx"), ("u", "dy"), ("s", "id")],
             3: [("u", "id"), ("u", "dx"), ("u", "dy"), ("u", "dz"), ("s", "id")]}  # 3D: src:736
DEFAULT_F: Dict[int, Callable] = {1: lambda x: [0.5, 0.0, 1.0], 2: lambda x: [0.5, 0.0, 0.0, 1.0],
                                  3: lambda x: [0.5, 0.0, 0.0, 0.0, 1.0]}  # 3D: src:737
DEFAULT_G: Dict[int, Callable] = {1: lambda x: [x[0], 2.0], 2: lambda x: [x[0] ** 2 + x[1] ** 2, 100.0],
                                  3: lambda x: [x[0] ** 2 + x[1] ** 2 + x[2] ** 2, 100.0]}  # 3D: src:738


@dataclass
class AMG:
    x: np.ndarray
    w: np.ndarray
    R_fine: List[sp.csr_matrix]   # per level: N x m_l
    D: List[sp.csr_matrix]        # finest-level operators, n x N
    nu: int
    nD: int
    state_variables: Tuple[Tuple[str, str], ...]
    D_table: List[Tuple[str, str]]


def amg_helper(geom: Geometry, state_variables: Sequence[Tuple[str, str]], D_table: Sequence[Tuple[str, str]]) -> AMG:
    L = len(geom.refine)
    n = geom.x.shape[0]
    refine_fine: List = [None] * L
    refine_fine[L - 1] = geom.refine[L - 1].tocsr()
    for l in range(L - 2, -1, -1):
        refine_fine[l] = (refine_fine[l + 1] @ geom.refine[l]).tocsr()
    R_fine = []
    for l in range(L):
        blocks = [(refine_fine[l] @ geom.subspaces[sub][l]).tocsr() for (_, sub) in state_variables]
        R = sp.block_diag(blocks, format="csr")
        R.sort_indices()
        R_fine.append(R)
    var_of = {name: k for k, (name, _) in enumerate(state_variables)}
    nu = len(state_variables)
    Dm = []
    for (var, opname) in D_table:
        blocks = [sp.csr_matrix((n, n)) for _ in range(nu)]
        blocks[var_of[var]] = geom.operators[opname].tocsr()
        M = sp.hstack(blocks, format="csr")
        M.sort_indices()
        Dm.append(M)
    return AMG(geom.x, geom.w, R_fine, Dm, nu, len(D_table), tuple(state_variables), list(D_table))


def amg(geom: Geometry, state_variables=DEFAULT_STATE, D_table=None):
    """Returns (M_main, M_feasibility): the second has the extra (:feasibility_slack, :full) state
    variable with operator :id appended (upstream ``amg``)."""
    D_table = DEFAULT_D[geom.dim] if D_table is None else list(D_table)
    M1 = amg_helper(geom, state_variables, D_table)
    sv2 = tuple(state_variables) + (("feasibility_slack", "full"),)
    M2 = amg_helper(geom, sv2, list(D_table) + [("feasibility_slack", "id")])
    return M1, M2
