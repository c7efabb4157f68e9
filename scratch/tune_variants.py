"""Build tuning variants of the library (in-tree, scratch/variants/) for an A/B bench run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mgb_b200
from mgb_b200 import build
os.makedirs(os.path.join(ROOT, "scratch", "variants"), exist_ok=True)
variants = {
    "mb4": ["-DMGB_ELEM_MINBLOCKS=4"], "mb6": ["-DMGB_ELEM_MINBLOCKS=6"],
    "t256mb2": ["-DMGB_ELEM_THREADS=256", "-DMGB_ELEM_MINBLOCKS=2"], "t256mb3": ["-DMGB_ELEM_THREADS=256", "-DMGB_ELEM_MINBLOCKS=3"],
    "t64mb10": ["-DMGB_ELEM_THREADS=64", "-DMGB_ELEM_MINBLOCKS=10"],
    "gu2": ["-DMGB_GATHER_UNROLL=2"], "gu8": ["-DMGB_GATHER_UNROLL=8"], "gu6": ["-DMGB_GATHER_UNROLL=6"],
}
for name, extra in variants.items():
    out = os.path.join(ROOT, "scratch", "variants", f"libmgb_{name}.so")
    build.build(force=True, out=out, extra=extra)
    print("built", out)
