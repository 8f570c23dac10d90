"""Not a test: prints per-intermediate and per-gradient errors of the CUDA path vs the oracle.
usage (on the GPU box): python tests/debug_stash.py [morph] [B] [use_tc]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import set_oracle as O
from sgrl_b200 import graph as G, morphologies as M, synth
import gpu_util, parity

name = sys.argv[1] if len(sys.argv) > 1 else "3d_humanoid_9_full"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
use_tc = int(sys.argv[3]) if len(sys.argv) > 3 else 0
actor, critic, pa, pc = gpu_util.make_modules(use_tc=use_tc)
par = M.ALL[name]; N = len(par)
g = G.build_graph(par, device="cuda")
b = gpu_util.to_cuda(synth.make_batch(B, N, seed=3))
g64 = dict(g); g64["relation"] = g["relation"].double()
for mod, params, prefix in ((actor, pa, "actor."), (critic, pc, "critic1."), (critic, pc, "critic2.")):
    mod.change_morphology(g)
    tb = mod._tables(B)
    z = 1 if prefix == "critic2." else 0
    act = b["action"].contiguous() if mod is critic else None
    out, stash = mod.forward_raw(tb, b["obs"].contiguous(), act, keep=True)
    torch.cuda.synchronize()
    x = b["obs"].view(B, N, 41) if mod is actor else torch.cat([b["obs"].view(B, N, 41), b["action"].view(B, N, 3)], 2)
    trace = {}
    p = {k: v.cuda().double() for k, v in O.sub(params, prefix).items()}
    with torch.no_grad():
        ref = O.transformer_model(p, x.double(), g64, trace=trace)
    errs = gpu_util.compare_stash(mod, stash, tb, mod._nb, z, trace)
    print(f"== {prefix} forward: worst intermediates")
    for k, e in sorted(errs, key=lambda t: -t[1])[:12]:
        print(f"   {k:12s} {e:.3e}")
    o = out[z]
    refo = torch.tanh(ref) if mod is actor else ref
    print("   OUT", parity.rel_err(o.reshape(-1), refo.reshape(-1)))

# gradients
critic.zero_grad(set_to_none=True)
q1, q2 = critic(b["obs"], b["action"]); tgt = b["reward"].expand_as(q1)
(F.mse_loss(q1, tgt) + F.mse_loss(q2, tgt)).backward()
p = {k: v.cuda().double().requires_grad_(not O.is_dead(k)) for k, v in pc.items()}
o1, o2 = O.critic_forward(p, b["obs"].double(), b["action"].double(), g64); t = b["reward"].double().expand_as(o1)
(F.mse_loss(o1, t) + F.mse_loss(o2, t)).backward()
rows = []
gn = torch.sqrt(sum((v.grad ** 2).sum() for v in p.values() if v.grad is not None)).item()
for k, prm in critic.named_parameters():
    w = p[k].grad
    if w is None: continue
    gg = prm.grad
    d = (gg.double() - w).norm().item() if gg is not None else float("nan")
    rows.append((d / max(w.norm().item(), 1e-4 * gn), k, w.norm().item()))
print("== critic grads: worst tensors (err/scale, name, ||g||)  global", gn)
for r in sorted(rows, key=lambda t: -(t[0] if t[0] == t[0] else 1e9))[:25]:
    print(f"   {r[0]:.3e} {r[1]} {r[2]:.3e}")
