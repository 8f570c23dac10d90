#!/usr/bin/env python
"""One net forward / backward in isolation, eager, for `ncu --metrics gpu__time_duration.sum` launch lists:
python tools/fwd_probe.py [--net actor|critic] [--batch B] [--bwd] [--reps R]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgrl_b200 import graph as G, morphologies as M, synth  # noqa: E402
from sgrl_b200.agent import Agent  # noqa: E402
from sgrl_b200.config import default_args  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--net", default="actor")
ap.add_argument("--morph", default="3d_humanoid_9_full")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--bwd", action="store_true")
ap.add_argument("--keep", type=int, default=1)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
torch.cuda.set_device(0)
torch.manual_seed(0)
ag = Agent(default_args())
par = M.ALL[a.morph]
g = G.build_graph(par, device="cuda")
ag.change_morphology(g)
B, N = a.batch, len(par)
b = {k: v.cuda() for k, v in synth.make_batch(B, N, seed=1).items()}
tb = ag.actor._tables(B)
T = tb.T
obs, act = b["obs"].reshape(T, 41).contiguous(), b["action"].reshape(T, 3).contiguous()
for _ in range(a.reps):
    if a.net == "actor":
        out, stash = ag.actor.forward_raw(tb, obs, None, keep=bool(a.keep), nb=1)
        if a.bwd:
            ag.actor.backward_raw(tb, stash, torch.ones_like(out), 1, ag.actor.grad_arena(), False)
    else:
        out, stash = ag.critic.forward_raw(tb, obs, act, keep=bool(a.keep), nb=2)
        if a.bwd:
            ag.critic.backward_raw(tb, stash, torch.ones_like(out), 2, ag.critic.grad_arena(), False)
    torch.cuda.synchronize()
print("done", T)
