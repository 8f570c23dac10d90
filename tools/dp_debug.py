#!/usr/bin/env python
"""torchrun debug aid: per step, are the all-reduced gradient arenas and the parameter arenas of all ranks bit-identical?
usage: torchrun --nproc-per-node 2 tools/dp_debug.py [set] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from sgrl_b200 import graph as G, morphologies as M, synth
from sgrl_b200.agent import Agent
from sgrl_b200.config import default_args

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sname = sys.argv[1] if len(sys.argv) > 1 else "3d_cwhh"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.manual_seed(0)
ag = Agent(default_args())
ag.use_graphs = os.environ.get("SGRL_GRAPHS", "1") != "0"
for m in (ag.actor, ag.actor_target, ag.critic, ag.critic_target):
    dist.broadcast(m.full_arena, 0)
names = sorted(M.SETS[sname])
mine = names[rank::world]
graphs = {n: G.build_graph(M.SETS[sname][n], device=dev) for n in mine}
bat = {n: {k: v.to(dev) for k, v in synth.make_batch(100, len(M.SETS[sname][n]), seed=300 + i).items()} for i, n in enumerate(mine)}


def same(t):
    bits = t.contiguous().view(torch.int32).to(torch.int64)
    s = torch.stack([bits.sum(), (bits * (torch.arange(bits.numel(), device=dev) % 8191 + 1)).sum()])
    allv = [torch.zeros_like(s) for _ in range(world)]
    dist.all_gather(allv, s)
    return all(bool((v == allv[0]).all()) for v in allv)


for it in range(steps):
    ag.update_packed([(graphs[n], bat[n]) for n in mine], it, morph_count=len(names) / world)
    torch.cuda.synchronize()
    r = {"critic_grad": same(ag.critic.grad_arena()), "actor_grad": same(ag.actor.grad_arena()), "critic": same(ag.critic.live_arena),
         "actor": same(ag.actor.live_arena), "critic_t": same(ag.critic_target.full_arena), "actor_t": same(ag.actor_target.full_arena),
         "adam_m": same(ag.critic_optimizer.exp_avg), "adam_v": same(ag.critic_optimizer.exp_avg_sq), "sumsq": same(ag.critic_optimizer.sumsq)}
    if rank == 0:
        print(it, {k: v for k, v in r.items()}, flush=True)
torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)
