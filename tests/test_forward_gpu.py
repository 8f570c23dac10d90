"""CUDA forward (through the C ABI) vs golden vectors from the reference and vs the oracle."""
import pytest
import torch

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


@pytest.mark.parametrize("name,B", parity.CASES)
def test_forward_matches_reference_golden(mods, gold, name, B):
    actor, critic, _, _ = mods
    g = parity.golden_graph(gold, name, M.ALL[name], device="cuda")
    b = parity.golden_batch(gold, name, device="cuda")
    actor.change_morphology(g); critic.change_morphology(g)
    with torch.no_grad():
        a = actor(b["obs"])
        q1, q2 = critic(b["obs"], b["action"])
        q1b = critic.Q1(b["obs"], b["action"])
    assert a.shape == (B, 3 * len(M.ALL[name])) and q1.shape == (B, len(M.ALL[name]))
    assert parity.rel_err(a, gold[name + "/actions"]) < parity.RTOL
    assert parity.rel_err(q1, gold[name + "/q1"]) < parity.RTOL
    assert parity.rel_err(q2, gold[name + "/q2"]) < parity.RTOL
    assert torch.equal(q1, q1b)


@pytest.mark.parametrize("name,B", [("3d_humanoid_9_full", 8), ("3d_walker_2_right_leg_left_knee", 4), ("3d_cheetah_14_full", 3)])
def test_every_intermediate_matches_oracle(mods, gold, name, B):
    """Kernel-by-kernel parity: each stash buffer vs the oracle's traced intermediate."""
    import gpu_util
    actor, critic, pa, pc = mods
    par = M.ALL[name]
    g = parity.golden_graph(gold, name, par, device="cuda")
    b = parity.golden_batch(gold, name, device="cuda")
    N = len(par)
    for mod, params, prefix, x in (
        (actor, pa, "actor.", b["obs"].view(B, N, 41)),
        (critic, pc, "critic1.", torch.cat([b["obs"].view(B, N, 41), b["action"].view(B, N, 3)], 2)),
        (critic, pc, "critic2.", torch.cat([b["obs"].view(B, N, 41), b["action"].view(B, N, 3)], 2)),
    ):
        mod.change_morphology(g)
        tb = mod._tables(B)
        nb = mod._nb
        z = 1 if prefix == "critic2." else 0
        act = b["action"].contiguous() if mod is critic else None
        out, stash = mod.forward_raw(tb, b["obs"].contiguous(), act, keep=True)
        trace = {}
        p = {k: v.cuda().double() for k, v in O.sub(params, prefix).items()}
        g64 = dict(g); g64["relation"] = g["relation"].double()
        with torch.no_grad():
            O.transformer_model(p, x.double(), g64, trace=trace)
        errs = gpu_util.compare_stash(mod, stash, tb, nb, z, trace)
        bad = [(k, e) for k, e in errs if e > 2e-5]
        assert not bad, f"{prefix} {name}: {bad[:8]}"


def test_rotation_about_gravity_invariance(mods):
    actor, critic, _, _ = mods
    par = M.ALL["3d_humanoid_9_full"]
    g = G.build_graph(par, device="cuda")
    actor.change_morphology(g); critic.change_morphology(g)
    b = synth.make_batch(64, len(par), seed=5)
    obs = b["obs"].cuda(); act = b["action"].cuda()
    rot = synth.rotate_about_gravity(obs, len(par), 0.7)
    with torch.no_grad():
        assert parity.rel_err(actor(rot), actor(obs)) < parity.RTOL_ROT
        q, qr = critic(obs, act), critic(rot, act)
        assert parity.rel_err(qr[0], q[0]) < parity.RTOL_ROT and parity.rel_err(qr[1], q[1]) < parity.RTOL_ROT


def test_large_batch_matches_oracle(mods):
    """BASELINE sizes (B=256, humanoid-9): compare with the fp32 oracle on the same GPU."""
    import gpu_util
    actor, critic, pa, pc = mods
    par = M.ALL["3d_humanoid_9_full"]
    g = G.build_graph(par, device="cuda")
    actor.change_morphology(g); critic.change_morphology(g)
    b = gpu_util.to_cuda(synth.make_batch(256, len(par), seed=1))
    pa_c = {k: v.cuda() for k, v in pa.items()}; pc_c = {k: v.cuda() for k, v in pc.items()}
    with torch.no_grad():
        assert parity.rel_err(actor(b["obs"]), O.actor_forward(pa_c, b["obs"], g)) < parity.RTOL
        q1, q2 = critic(b["obs"], b["action"])
        o1, o2 = O.critic_forward(pc_c, b["obs"], b["action"], g)
        assert parity.rel_err(q1, o1) < parity.RTOL and parity.rel_err(q2, o2) < parity.RTOL


def test_critic_width_assert(mods):
    _, critic, _, _ = mods
    g = G.build_graph(M.ALL["3d_hopper_3_shin"], device="cuda")
    critic.change_morphology(g)
    with pytest.raises(AssertionError):
        critic(torch.zeros(2, 41 * 4, device="cuda"), torch.zeros(2, 12, device="cuda"))
