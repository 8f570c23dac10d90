#!/usr/bin/env python
"""Workload for tools/sanitize.sh: small batches so that compute-sanitizer's 10-100x slowdown stays within minutes, but
T3*32*128 >= 2^21 so the tcgen05 / fused forward paths are the ones that run (B=64 humanoid-9: 576 tokens)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from sgrl_b200 import graph as G, morphologies as M, synth
import gpu_util

tc = int(os.environ.get("SGRL_SAN_TC", "1"))
actor, critic, pa, pc = gpu_util.make_modules(use_tc=tc)
par = M.ALL["3d_humanoid_9_full"]
g = G.build_graph(par, device="cuda")
B = 64
b = gpu_util.to_cuda(synth.make_batch(B, len(par), seed=1))
actor.change_morphology(g); critic.change_morphology(g)
q1, q2 = critic(b["obs"], b["action"])
((q1 - b["reward"]) ** 2 + (q2 - b["reward"]) ** 2).mean().backward()
(-critic.Q1(b["obs"], actor(b["obs"])).mean()).backward()
torch.cuda.synchronize()
if os.environ.get("SGRL_SAN_UPDATE", "1") == "1":
    from test_agent_gpu import make_agent
    ag, _, _ = make_agent(tc)
    ag.change_morphology(g)
    for it in range(4):          # eager, eager, capture + replay, capture + replay
        ag.update(b, it)
    torch.cuda.synchronize()
print("sanitize_target done")
