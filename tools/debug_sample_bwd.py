"""Not a test: per-sample backward of the twin critic (module path) vs the fp64 oracle, to find samples whose gradient differs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import set_oracle as O
from sgrl_b200 import graph as G, morphologies as M, synth
import gpu_util, parity

B = 100
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 51
actor, critic, pa, pc = gpu_util.make_modules(use_tc=int(os.environ.get("USE_TC", "1")))
par = M.ALL["3d_humanoid_9_full"]; N = len(par)
g = G.build_graph(par, device="cuda")
g64 = dict(g); g64["relation"] = g["relation"].double()
critic.change_morphology(g)
b = {k: v.cuda() for k, v in synth.make_batch(B, N, seed=seed).items()}
p64 = {k: v.cuda().double().requires_grad_(not O.is_dead(k)) for k, v in pc.items()}
w = torch.randn(2, B, N, generator=torch.Generator().manual_seed(5)).cuda()

def ours(mask):
    critic.zero_grad(set_to_none=True)
    q1, q2 = critic(b["obs"], b["action"])
    ((q1 * w[0] + q2 * w[1]) * mask[:, None]).sum().backward()
    return {k: (p.grad.detach().double().clone() if p.grad is not None else None) for k, p in critic.named_parameters()}

def ref(mask):
    for v in p64.values():
        v.grad = None
    q1, q2 = O.critic_forward(p64, b["obs"].double(), b["action"].double(), g64)
    ((q1 * w[0].double() + q2 * w[1].double()) * mask.double()[:, None]).sum().backward()
    return {k: (v.grad.detach().clone() if v.grad is not None else None) for k, v in p64.items()}

def err(a, r, net):
    num = den = 0.0
    for k, v in r.items():
        if v is None or not k.startswith(net):
            continue
        num += (a[k] - v).pow(2).sum().item(); den += v.pow(2).sum().item()
    return (num / max(den, 1e-300)) ** 0.5

full = torch.ones(B, device="cuda")
a, r = ours(full), ref(full)
print(f"seed {seed} all samples: critic1 {err(a, r, 'critic1'):.2e} critic2 {err(a, r, 'critic2'):.2e}")
worst = []
for s in range(B):
    m = torch.zeros(B, device="cuda"); m[s] = 1
    a, r = ours(m), ref(m)
    worst.append((max(err(a, r, 'critic1'), err(a, r, 'critic2')), s, err(a, r, 'critic1'), err(a, r, 'critic2')))
worst.sort(reverse=True)
print("worst samples (max err, sample, critic1, critic2):", [(f"{e:.2e}", s, f"{e1:.2e}", f"{e2:.2e}") for e, s, e1, e2 in worst[:6]])
s = worst[0][1]
o = b["obs"][s].view(N, 41)
print("worst sample obs stats: min", o.min().item(), "max", o.max().item(), "done", b["done"][s].item(), "reward", b["reward"][s].item())
print("action", b["action"][s].view(N, 3))
