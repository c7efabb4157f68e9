"""End-to-end differential tests, same pattern as the reference's integration suite
(test/test_quick.jl:108-140, test/test_2d.jl:120-161): solve with the GPU assembly, solve with the
CPU oracle, compare.  north_star bar: identical Newton iteration counts, solutions within 1e-9 rel."""
import numpy as np
import pytest

import mgb_b200
from mgb_b200 import solver
import mgb_oracle as O

pytestmark = pytest.mark.gpu


def _compare(geom, p, **kw):
    sol_o = O.amgb(geom, p=p, **kw)
    sol_g = solver.amgb(geom, p=p, **kw)
    assert np.array_equal(sol_g.SOL_main["ts"], sol_o.SOL_main["ts"])
    assert np.array_equal(sol_g.SOL_main["its"], sol_o.SOL_main["its"]), (sol_g.SOL_main["its"], sol_o.SOL_main["its"])
    rel = np.linalg.norm(sol_g.z - sol_o.z) / np.linalg.norm(sol_o.z)
    assert rel < 1e-9, rel
    assert np.allclose(sol_g.SOL_main["c_dot_Dz"], sol_o.SOL_main["c_dot_Dz"], rtol=1e-10)
    return sol_g, sol_o


def test_fem1d_L3_p1_reference_quick_case():
    sol_g, _ = _compare(mgb_b200.fem1d(3), 1.0)
    assert sol_g.z.shape == (16, 2)
    assert sol_g.SOL_feasibility is None


def test_fem2d_L2_p2_reference_2d_case():
    _compare(mgb_b200.fem2d(2), 2.0)


def test_fem2d_L3_p1_baseline_config_C1():
    sol_g, _ = _compare(mgb_b200.fem2d(3), 1.0)
    assert sol_g.stats["assemblies"] > 0 and sol_g.stats["f0_evals"] > 0


def test_fem2d_L3_p1p5():
    _compare(mgb_b200.fem2d(3), 1.5)


def test_feasibility_phase_runs_and_matches():
    g = lambda x: [x[0] ** 2 + x[1] ** 2, 0.5]   # s = 0.5 < |grad u| near the corners: infeasible start
    sol_g, sol_o = _compare(mgb_b200.fem2d(2), 1.0, g=g)
    assert sol_g.SOL_feasibility is not None and sol_o.SOL_feasibility is not None
    assert np.array_equal(sol_g.SOL_feasibility["its"], sol_o.SOL_feasibility["its"])


def test_fem3d_L3_whole_solve_uses_dense_coarse_levels():
    """config C4 as a whole solve: fem3d (Q3 hexahedra, reference problem data src/MultiGridBarrierMPI.jl:735-745) on a
    3-level hierarchy - coarse levels on the dense contraction path, the finest on the CSR path - with the oracle's
    t-schedule, Newton counts per level and iterate"""
    _compare(mgb_b200.fem3d(3), 1.0)
