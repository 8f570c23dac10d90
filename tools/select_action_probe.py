import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from sgrl_b200 import graph as G, morphologies as M, synth
from sgrl_b200.agent import Agent
from sgrl_b200.config import default_args
torch.manual_seed(0)
ag = Agent(default_args())
ag.use_graphs = False
par = M.ALL["3d_humanoid_9_full"]
ag.change_morphology(G.build_graph(par, device="cuda"))
obs = synth.make_obs(1, len(par), seed=3)[0].numpy()
for _ in range(6):
    ag.select_action(obs)
torch.cuda.synchronize()
