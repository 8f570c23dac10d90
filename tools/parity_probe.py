#!/usr/bin/env python
"""Prints the per-tensor gradient errors (vs the fp64 oracle) and the relu units on the other side of their kink for the
cases of tests/test_parity_wide_gpu.py.  usage: python tools/parity_probe.py [humanoid256|chain15|bushy15|sweep]..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from sgrl_b200 import graph as G, morphologies as M, synth  # noqa: E402
import gpu_util  # noqa: E402
import test_parity_wide_gpu as W  # noqa: E402


def case(name, parents, B, seed):
    actor, critic, pa, pc = gpu_util.make_modules(use_tc=int(os.environ.get("USE_TC", "1")))
    g = G.build_graph(parents, device="cuda")
    b = gpu_util.to_cuda(synth.make_batch(B, len(parents), seed=seed))
    print(f"# {name}: N={len(parents)} B={B}")
    print("  " + str(W._critic_case(critic, pc, g, b))[:900])
    print("  " + str(W._actor_case(actor, critic, pa, pc, g, b))[:900])


for what in sys.argv[1:] or ["humanoid256", "chain15", "bushy15"]:
    if what == "humanoid256":
        case(what, M.ALL["3d_humanoid_9_full"], 256, 1)
    elif what == "chain15":
        case(what, [-1] + list(range(14)), 32, 3)
    elif what == "bushy15":
        case(what, [-1, 0, 1, 2, 2, 1, 5, 5, 0, 8, 9, 9, 8, 12, 12], 16, 3)
    elif what == "cheetah64":
        case(what, M.ALL["3d_cheetah_14_full"], 64, 1)
