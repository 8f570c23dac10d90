#!/usr/bin/env python
"""Where do the tcgen05 and the fp32 SIMT GEMM paths disagree?  For one morphology / batch: d Q1 / d action of the critic,
then the actor's parameter gradients for a FIXED upstream gradient, tensor by tensor.  usage: tc_vs_simt_probe.py N B"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from sgrl_b200 import graph as G, synth  # noqa: E402
import gpu_util  # noqa: E402
import parity  # noqa: E402

N, B = int(sys.argv[1]), int(sys.argv[2])
parents = [-1] + list(range(N - 1))
g = G.build_graph(parents, device="cuda")
b = gpu_util.to_cuda(synth.make_batch(B, N, seed=3))
res = {}
for tc in (0, 1):
    actor, critic, pa, pc = gpu_util.make_modules(use_tc=tc)
    actor.change_morphology(g); critic.change_morphology(g)
    act_in = b["action"].clone().requires_grad_(True)
    critic.zero_grad(set_to_none=True)
    critic.Q1(b["obs"], act_in).mean().backward()
    res[tc, "dact"] = act_in.grad.clone()
    up = torch.randn(B, 3 * N, generator=torch.Generator().manual_seed(1)).cuda()
    actor.zero_grad(set_to_none=True)
    out = actor(b["obs"])
    res[tc, "a"] = out.detach().clone()
    out.backward(up)
    res[tc, "g"] = {k: p.grad.clone() for k, p in actor.named_parameters() if p.grad is not None}
print(f"N={N} B={B} T={N * B}: actions tc vs simt {parity.rel_err(res[1, 'a'], res[0, 'a']):.2e}; dQ1/daction {parity.rel_err(res[1, 'dact'], res[0, 'dact']):.2e}")
rows = sorted(((parity.rel_err(res[1, 'g'][k], v), k) for k, v in res[0, "g"].items()), reverse=True)
print("actor grads for a fixed upstream gradient, worst tensors:", [(k.split("encoder.")[-1], f"{e:.1e}") for e, k in rows[:8]])
