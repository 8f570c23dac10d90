"""Not a test: per-iteration gradient errors of Agent.update vs a re-synchronised fp64 oracle. usage: python tests/debug_update_grads.py [B] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import set_oracle as O
from sgrl_b200 import graph as G, morphologies as M, synth
import parity
from test_agent_gpu import make_agent, _arena_grads
from test_backward_gpu import grad_report

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ag, _, _ = make_agent(int(os.environ.get("USE_TC", "1")))
par = M.ALL["3d_humanoid_9_full"]
g = G.build_graph(par, device="cuda")
g64 = dict(g); g64["relation"] = g["relation"].double()
ag.change_morphology(g)
for it in range(int(os.environ.get("START_IT", "0")), int(os.environ.get("START_IT", "0")) + iters):
    cur = lambda mod: {k: v.detach().double().clone() for k, v in mod.state_dict().items()}
    td = O.TD3Oracle(cur(ag.actor), cur(ag.critic))
    td.actor_t, td.critic_t = cur(ag.actor_target), cur(ag.critic_target)
    cur32 = lambda mod: {k: v.detach().clone() for k, v in mod.state_dict().items()}
    td32 = O.TD3Oracle(cur32(ag.actor), cur32(ag.critic))
    td32.actor_t, td32.critic_t = cur32(ag.actor_target), cur32(ag.critic_target)
    b = {k: v.cuda() for k, v in synth.make_batch(B, len(par), seed=50 + it).items()}
    noise = torch.randn(B, 27, generator=torch.Generator().manual_seed(100 + it)).cuda() * 0.2
    ld = ag.update(b, it, noise=noise)
    ref = td.update({k: v.double() for k, v in b.items()}, it, noise.double(), g64)
    td32.update(b, it, noise, g)
    bad32, glob32 = grad_report(td32.critic_grad, td.critic_grad)
    print(f"it {it}: torch-fp32 oracle vs fp64 oracle: critic grads global {glob32:.2e}, {len(bad32)} bad")
    got = _arena_grads(ag.critic)
    bad, glob = grad_report(got, td.critic_grad)
    big = sorted(((k, (got[k].double() - w).norm().item()) for k, w in td.critic_grad.items() if w is not None), key=lambda x: -x[1])[:3]
    print("     largest absolute errors:", [(k, f"{d:.2e}", f"ref {td.critic_grad[k].norm().item():.2e}") for k, d in big])
    print(f"it {it}: loss {ld['loss/critic_loss'].item():.6e} vs {ref['loss/critic_loss'].item():.6e}; critic grads global {glob:.2e}, {len(bad)} bad")
    for k, e in sorted(bad, key=lambda x: -x[1])[:5]:
        w = td.critic_grad[k]
        print(f"     {k:70s} err/scale {e:.2e}  |ref| {w.norm().item():.3e}  |got| {got[k].double().norm().item():.3e}")
    if it % 2 == 0:
        got = _arena_grads(ag.actor)
        bad, glob = grad_report(got, td.actor_grad)
        print(f"       actor grads global {glob:.2e}, {len(bad)} bad")
