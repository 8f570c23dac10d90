#!/usr/bin/env python
"""Is the rollout forward bit-deterministic?  (no atomics on that path: any run-to-run difference is a race)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sgrl_b200 import graph as G, morphologies as M, synth
from sgrl_b200.agent import Agent
from sgrl_b200.config import default_args
torch.manual_seed(0)
ag = Agent(default_args())
par = M.ALL["3d_humanoid_9_full"]
ag.change_morphology(G.build_graph(par, device="cuda"))
for envs in (256, 8192):
    obs = synth.make_obs(min(envs, 4096), len(par), seed=7).cuda()
    obs = obs.repeat((envs + obs.shape[0] - 1) // obs.shape[0], 1)[:envs].contiguous()
    with torch.no_grad():
        outs = [ag.actor(obs).clone() for _ in range(6)]
    diffs = [(o - outs[0]).abs().max().item() for o in outs[1:]]
    print(f"envs {envs}: max |delta| over 5 repeats = {max(diffs):.3e}  ({'bit-identical' if max(diffs) == 0 else 'DIFFERS'})")
    critic_in = torch.rand(envs, 27, device="cuda") * 2 - 1
    with torch.no_grad():
        qs = [ag.critic(obs, critic_in)[0].clone() for _ in range(4)]
    print(f"   critic: {max((q - qs[0]).abs().max().item() for q in qs[1:]):.3e}")
