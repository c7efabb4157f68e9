"""CUDA-graph launch of an assembly (mgb_assemble's cache) and of a recorded call sequence (mgb_graph_begin /
mgb_graph_end / mgb_graph_launch): bit-identical to the plain launches."""
import numpy as np
import pytest
import torch

import mgb_b200
from mgb_b200 import capi

from helpers import problem

pytestmark = pytest.mark.gpu


def _dev_inputs(pr, dev):
    Dz0 = np.stack([Dk @ pr["z0"] for Dk in pr["D"]], axis=1)
    cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)
    return torch.from_numpy(pr["s"]).to(dev), cm(Dz0), cm(pr["c"])


@pytest.mark.parametrize("gen,L,level", [("fem2d", 5, None), ("fem2d", 4, 1), ("fem1d", 6, None), ("fem3d", 2, None)])
def test_graph_cache_matches_plain_launches(gen, L, level, monkeypatch):
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        ctx = capi.Context(0, stream.cuda_stream)
        pr = problem(getattr(mgb_b200, gen)(L), level=level)
        s_d, Dz0_d, c_d = _dev_inputs(pr, dev)
        outs = {}
        for mode in ("0", "1"):
            monkeypatch.setenv("MGB_GRAPH", mode)
            plan = capi.Plan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0)
            scal = torch.zeros(4, dtype=torch.float64, device=dev)
            grad = torch.zeros(plan.m, dtype=torch.float64, device=dev)
            hval = torch.zeros(plan.nnzH, dtype=torch.float64, device=dev)
            for t in (0.7, 0.7, 1.3, 0.7):     # repeated and changing scalars: cache hits and re-captures
                plan.assemble(s_d, Dz0_d, c_d, t, 7, scal, grad, hval)
            plan.assemble(s_d, Dz0_d, c_d, 0.7, 1, scal)   # objective only: another graph
            plan.assemble(s_d, Dz0_d, c_d, 0.7, 7, scal, grad, hval)
            ctx.sync()
            st = plan.graph_stats()
            if mode == "1":
                assert st["state"] == 1 and st["captures"] == 3 and st["launches"] == 3, st
            else:
                assert st["state"] == -1 and st["captures"] == 0
            outs[mode] = (scal.cpu().numpy().copy(), grad.cpu().numpy().copy(), hval.cpu().numpy().copy())
            plan.close()
        for a, b in zip(outs["0"], outs["1"]):
            assert np.array_equal(a, b)
        ctx.close()


def test_recorded_level_sweep_replays_bit_identically():
    """one assembly on every level of a fem2d hierarchy, recorded once and replayed with a single launch"""
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        ctx = capi.Context(0, stream.cuda_stream)
        geom = mgb_b200.fem2d(4)
        levels = []
        for J in range(4):
            pr = problem(geom, level=J)
            plan = capi.Plan(ctx, pr["D"], pr["R"], pr["x"], pr["w"], pr["idx"], 1.0)
            s_d, Dz0_d, c_d = _dev_inputs(pr, dev)
            bufs = (torch.zeros(4, dtype=torch.float64, device=dev), torch.zeros(plan.m, dtype=torch.float64, device=dev),
                    torch.zeros(plan.nnzH, dtype=torch.float64, device=dev))
            levels.append((plan, s_d, Dz0_d, c_d, bufs))
        def sweep():
            for plan, s_d, Dz0_d, c_d, (scal, grad, hval) in levels:
                plan.assemble(s_d, Dz0_d, c_d, 0.9, 7, scal, grad, hval)
        sweep()
        ctx.sync()
        ref = [tuple(b.cpu().numpy().copy() for b in lv[4]) for lv in levels]
        for lv in levels:
            for b in lv[4]:
                b.zero_()
        ctx.graph_begin()
        sweep()                      # recorded, not executed
        graph = ctx.graph_end()
        ctx.sync()
        assert all(float(lv[4][2].abs().sum()) == 0.0 for lv in levels), "capture must not execute"
        n0 = capi.launch_count()
        graph.launch()
        ctx.sync()
        assert capi.launch_count() - n0 >= 2 * len(levels)
        for lv, r in zip(levels, ref):
            for b, a in zip(lv[4], r):
                assert np.array_equal(b.cpu().numpy(), a)
        graph.close()
        for lv in levels:
            lv[0].close()
        ctx.close()
