"""CUDA backward (loss.backward() through the drop-in modules) vs the oracle's autograd
and vs gradient summaries recorded from the reference.

Metric: per tensor ||dg|| <= 1e-4 * max(||g_t||, 1e-4 * ||g||_global)
and globally ||dg|| / ||g|| <= 1e-4.  rel_encoder.bias has an exactly-zero true gradient
(softmax shift invariance); the reference produces rounding noise there, we produce 0."""
import pytest
import torch
import torch.nn.functional as F

from oracle import set_oracle as O
from sgrl_b200 import graph as G, morphologies as M, synth
import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[1, 0], ids=["tcgen05", "simt"])
def mods(request):
    import gpu_util
    return gpu_util.make_modules(use_tc=request.param)


@pytest.fixture(scope="module")
def gold():
    return parity.load_golden()


def grad_report(got: dict, want: dict, rtol=parity.RTOL, floor=1e-4):
    gn = torch.sqrt(sum((v.double() ** 2).sum() for v in want.values() if v is not None)).item()
    bad, tot = [], 0.0
    for k, w in want.items():
        if w is None:
            assert got.get(k) is None or float(got[k].abs().max()) == 0.0, f"{k}: dead tensor received a gradient"
            continue
        g = got[k]
        assert g is not None, f"{k}: missing gradient"
        d = (g.double().cpu() - w.double().cpu()).norm().item()
        tot += d * d
        if k.endswith("rel_encoder.bias"):
            continue
        scale = max(w.double().norm().item(), floor * gn)
        if d > rtol * scale:
            bad.append((k, d / scale))
    return bad, (tot ** 0.5) / gn


def oracle_grads(params, loss_fn):
    p = {k: v.cuda().double().requires_grad_(not O.is_dead(k)) for k, v in params.items()}
    loss = loss_fn(p)
    loss.backward()
    return loss.item(), {k: v.grad for k, v in p.items()}


@pytest.mark.parametrize("name,B", parity.CASES[:4] + [("3d_humanoid_9_full", 100)])
def test_critic_gradients(mods, gold, name, B):
    _, critic, _, pc = mods
    par = M.ALL[name]
    if B == 100:
        import gpu_util
        g = G.build_graph(par, device="cuda"); b = gpu_util.to_cuda(synth.make_batch(B, len(par), seed=1))
    else:
        g = parity.golden_graph(gold, name, par, device="cuda"); b = parity.golden_batch(gold, name, device="cuda")
    critic.change_morphology(g)
    critic.zero_grad(set_to_none=True)
    q1, q2 = critic(b["obs"], b["action"])
    tgt = b["reward"].expand_as(q1)
    loss = F.mse_loss(q1, tgt) + F.mse_loss(q2, tgt)
    loss.backward()
    got = {k: p.grad for k, p in critic.named_parameters()}
    g64 = dict(g); g64["relation"] = g["relation"].double()

    def lf(p):
        o1, o2 = O.critic_forward(p, b["obs"].double(), b["action"].double(), g64)
        t = b["reward"].double().expand_as(o1)
        return F.mse_loss(o1, t) + F.mse_loss(o2, t)
    want_loss, want = oracle_grads(pc, lf)
    assert abs(loss.item() - want_loss) < 1e-5 * abs(want_loss)
    bad, glob = grad_report(got, want)
    assert not bad and glob < parity.RTOL, f"global {glob:.2e}; worst {sorted(bad, key=lambda x: -x[1])[:8]}"
    if B != 100:
        parity.check_summary(parity.summarize(got), gold[name + "/critic_grad"], what="critic grad vs reference golden")


@pytest.mark.parametrize("name,B", parity.CASES[:4])
def test_actor_gradients_and_daction(mods, gold, name, B):
    actor, critic, pa, pc = mods
    par = M.ALL[name]
    g = parity.golden_graph(gold, name, par, device="cuda"); b = parity.golden_batch(gold, name, device="cuda")
    actor.change_morphology(g); critic.change_morphology(g)
    # d Q1 / d action
    act_in = b["action"].clone().requires_grad_(True)
    critic.zero_grad(set_to_none=True)
    critic.Q1(b["obs"], act_in).mean().backward()
    assert parity.rel_err(act_in.grad, gold[name + "/dq1_daction"]) < parity.RTOL
    # actor loss through critic1
    actor.zero_grad(set_to_none=True); critic.zero_grad(set_to_none=True)
    aloss = -critic.Q1(b["obs"], actor(b["obs"])).mean()
    aloss.backward()
    assert abs(aloss.item() - float(gold[name + "/aloss0"])) < 1e-4 * abs(float(gold[name + "/aloss0"]))
    got = {k: p.grad for k, p in actor.named_parameters()}
    g64 = dict(g); g64["relation"] = g["relation"].double()
    pc64 = {k: v.cuda().double() for k, v in pc.items()}

    def lf(p):
        return -O.critic_forward(pc64, b["obs"].double(), O.actor_forward(p, b["obs"].double(), g64), g64, which=(1,)).mean()
    _, want = oracle_grads(pa, lf)
    bad, glob = grad_report(got, want)
    assert not bad and glob < parity.RTOL, f"global {glob:.2e}; worst {sorted(bad, key=lambda x: -x[1])[:8]}"
    parity.check_summary(parity.summarize(got), gold[name + "/actor_grad"], what="actor grad vs reference golden")
    # critic2 did not run in Q1: no gradient there; critic1 received (discarded) gradients like the reference
    assert all(p.grad is None for k, p in critic.named_parameters() if k.startswith("critic2."))


def test_gradient_accumulates_like_torch(mods, gold):
    _, critic, _, _ = mods
    name = "3d_hopper_3_shin"
    g = parity.golden_graph(gold, name, M.ALL[name], device="cuda"); b = parity.golden_batch(gold, name, device="cuda")
    critic.change_morphology(g)
    critic.zero_grad(set_to_none=True)
    critic(b["obs"], b["action"])[0].sum().backward()
    g1 = {k: p.grad.clone() for k, p in critic.named_parameters() if p.grad is not None}
    critic(b["obs"], b["action"])[0].sum().backward()
    k = "critic1.linear2_g.weight"
    assert parity.rel_err(dict(critic.named_parameters())[k].grad, 2 * g1[k]) < 1e-5
