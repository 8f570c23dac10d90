"""Packed mixed-morphology batches (BASELINE configs "Walker++ multi-morphology update", "cwhh"):
one TD3 update over the limb-tokens of several morphologies at once must reproduce, on the real limbs of every morphology,
what the per-morphology path computes, and its step must equal the oracle's step on the loss
mean_over_morphologies( per-morphology reference loss ) — the reference itself steps the morphologies one after the other
(src/trainer.py:245-250), so the packed gradient is the average of the gradients of those separate updates."""
import pytest
import torch

from oracle import set_oracle as O
from sgrl_b200 import graph as G, morphologies as M, synth
import parity
from test_agent_gpu import make_agent

pytestmark = pytest.mark.gpu

MORPHS = [("3d_hopper_5_full", 12), ("3d_humanoid_9_full", 7), ("3d_cheetah_14_full", 5), ("3d_walker_7_full", 9)]


def _parts(device="cuda"):
    out = []
    for i, (name, B) in enumerate(MORPHS):
        par = M.ALL[name]
        g = G.build_graph(par, device=device)
        b = {k: v.to(device) for k, v in synth.make_batch(B, len(par), seed=40 + i).items()}
        nz = (torch.randn(B, 3 * len(par), generator=torch.Generator().manual_seed(70 + i)) * 0.2).to(device)
        out.append((g, b, nz))
    return out


def test_packed_forward_equals_per_morphology():
    from sgrl_b200.modules import make_packed_tables
    ag, pa, pc = make_agent()
    parts = _parts()
    tb = make_packed_tables([(g, b["obs"].shape[0]) for g, b, _ in parts], "cuda")
    assert tb.T == sum(b["obs"].shape[0] * len(g["parents"]) for g, b, _ in parts)
    obs = torch.cat([b["obs"].reshape(-1, 41) for _, b, _ in parts]).contiguous()
    act = torch.cat([b["action"].reshape(-1, 3) for _, b, _ in parts]).contiguous()
    with torch.no_grad():
        pi, _ = ag.actor.forward_raw(tb, obs, None, keep=False)
        q, _ = ag.critic.forward_raw(tb, obs, act, keep=False, nb=2)
        for (t0, t1, g0, g1, n), (g, b, _) in zip(tb.parts, parts):
            ag.change_morphology(g)
            a_one = ag.actor(b["obs"])                       # (B, 3n) through the single-morphology module path
            q1, q2 = ag.critic(b["obs"], b["action"])
            # same kernels, but the tile shapes / accumulator rotation depend on the token count: fp32 reassociation only
            assert parity.rel_err(pi[0, t0:t1].reshape(a_one.shape), a_one) < 1e-5
            assert parity.rel_err(q[0, t0:t1].reshape(q1.shape), q1) < 1e-5
            assert parity.rel_err(q[1, t0:t1].reshape(q2.shape), q2) < 1e-5
            # and against the oracle on this morphology
            gc = G.build_graph(g["parents"])
            assert parity.rel_err(a_one, O.actor_forward(pa, b["obs"].cpu(), gc)) < parity.RTOL


def _oracle_packed_update(pa, pc, pa_t, pc_t, parts, it, opt_c, opt_a, lr=1e-4):
    """The reference TD3 step (src/agent.py:117-183) restated for the packed loss: mean over morphologies of the per-morphology loss."""
    m = len(parts)
    loss_c = 0.0
    for g, b, nz in parts:
        with torch.no_grad():
            a2 = (O.actor_forward(pa_t, b["next_obs"], g) + nz.clamp(-0.5, 0.5)).clamp(-1.0, 1.0)
            tq1, tq2 = O.critic_forward(pc_t, b["next_obs"], a2, g)
            y = b["reward"] + (1 - b["done"]) * 0.99 * torch.min(tq1, tq2)
        q1, q2 = O.critic_forward(pc, b["obs"], b["action"], g)
        loss_c = loss_c + (torch.nn.functional.mse_loss(q1, y) + torch.nn.functional.mse_loss(q2, y)) / m
    opt_c.zero_grad()
    loss_c.backward()
    torch.nn.utils.clip_grad_norm_([p for p in pc.values() if p.requires_grad], 0.1)
    opt_c.step()
    out = {"critic": loss_c.item()}
    if it % 2 == 0:
        loss_a = 0.0
        for g, b, _ in parts:
            loss_a = loss_a - O.critic_forward(pc, b["obs"], O.actor_forward(pa, b["obs"], g), g, which=(1,)).mean() / m
        opt_a.zero_grad()
        loss_a.backward()
        torch.nn.utils.clip_grad_norm_([p for p in pa.values() if p.requires_grad], 0.1)
        opt_a.step()
        out["actor"] = loss_a.item()
        with torch.no_grad():
            for src, dst in ((pc, pc_t), (pa, pa_t)):
                for k in src:
                    dst[k].mul_(1 - 0.005).add_(src[k].detach(), alpha=0.005)
    return out


@pytest.mark.parametrize("graphs", [True, False], ids=["graph", "eager"])
def test_packed_update_matches_oracle_step(graphs):
    ag, pa0, pc0 = make_agent()
    ag.use_graphs = graphs
    parts = _parts()
    dd = torch.float64
    conv = lambda d, grad: {k: v.cuda().to(dd).requires_grad_(grad and not O.is_dead(k)) for k, v in d.items()}
    pa, pc, pa_t, pc_t = conv(pa0, True), conv(pc0, True), conv(pa0, False), conv(pc0, False)
    opt_c = torch.optim.Adam([p for p in pc.values() if p.requires_grad], lr=1e-4)
    opt_a = torch.optim.Adam([p for p in pa.values() if p.requires_grad], lr=1e-4)
    parts64 = []
    for g, b, nz in parts:
        g64 = dict(g); g64["relation"] = g["relation"].double()
        parts64.append((g64, {k: v.double() for k, v in b.items()}, nz.double()))
    for it in range(4):          # it = 2, 3 run through the captured graphs when graphs=True
        ld = ag.update_packed([(g, b) for g, b, _ in parts], it, noises=[nz for _, _, nz in parts])
        ref = _oracle_packed_update(pa, pc, pa_t, pc_t, parts64, it, opt_c, opt_a)
        assert abs(ld["loss/critic_loss"].item() - ref["critic"]) < 2e-4 * abs(ref["critic"]), (it, ld["loss/critic_loss"].item(), ref["critic"])
        if it % 2 == 0:
            assert abs(ld["loss/actor_loss"].item() - ref["actor"]) < 5e-4 * abs(ref["actor"]), (it, ld["loss/actor_loss"].item(), ref["actor"])

    def traj_err(mod, ref_p, init):
        num = den = 0.0
        for k, p in mod.named_parameters():
            d0 = init[k].cuda().double()
            num += ((p.double() - d0) - (ref_p[k].detach() - d0)).pow(2).sum().item()
            den += (ref_p[k].detach() - d0).pow(2).sum().item()
        return (num / den) ** 0.5

    # fp32 Adam trajectories against an fp64 oracle: same scale as the single-morphology test (test_agent_gpu.py)
    assert traj_err(ag.critic, pc, pc0) < 3e-2
    assert traj_err(ag.actor, pa, pa0) < 3e-2
    assert traj_err(ag.critic_target, pc_t, pc0) < 3e-2
