"""Rollout front-end (Agent.select_action as a replayed CUDA graph, Agent.select_actions over mixed
morphologies) vs the oracle's actor forward and vs the env-by-env path."""
import numpy as np
import pytest
import torch

from oracle import set_oracle as O
from sgrl_b200 import graph as G, morphologies as M, synth
import parity
from test_agent_gpu import make_agent

pytestmark = pytest.mark.gpu

NAMES = ["3d_humanoid_9_full", "3d_hopper_3_shin", "3d_walker_7_full", "3d_cheetah_14_full", "3d_walker_2_right_leg_left_knee",
         "3d_humanoid_9_full"]


def _oracle_actions(pa, obs, g):
    pa_c = {k: v.cuda() for k, v in pa.items()}
    with torch.no_grad():
        return O.actor_forward(pa_c, torch.as_tensor(obs, dtype=torch.float32).cuda().reshape(1, -1), g).cpu().numpy()


def test_select_action_graph_replay_matches_oracle():
    ag, pa, _ = make_agent()
    par = M.ALL["3d_humanoid_9_full"]
    g = G.build_graph(par, device="cuda")
    ag.change_morphology(g)
    for i in range(4):                                   # call 0 eager, call 1 captures, calls 2.. replay
        obs = synth.make_obs(1, len(par), seed=40 + i)[0].numpy().astype(np.float64)     # ModularEnv hands out float64
        a = ag.select_action(obs)
        assert a.shape == (1, 3 * len(par)) and a.dtype == np.float32
        assert parity.rel_err(a, _oracle_actions(pa, obs, g)) < parity.RTOL
    plan = next(iter(ag._rollout_plans.values()))
    assert plan.graph is not None and plan.graph_launches > 30
    # a (B, 41N) batch goes through the same path
    obs = synth.make_obs(5, len(par), seed=3).numpy()
    a = ag.select_action(obs)
    with torch.no_grad():
        want = O.actor_forward({k: v.cuda() for k, v in pa.items()}, torch.tensor(obs).cuda(), g)
    assert a.shape == (5, 27) and parity.rel_err(a, want) < parity.RTOL
    with pytest.raises(RuntimeError, match="invalid for input"):
        ag.select_action(np.zeros(41 * 4))


def test_select_actions_mixed_morphologies():
    ag, pa, _ = make_agent()
    graphs = {n: G.build_graph(M.ALL[n], device="cuda") for n in set(NAMES)}
    glist = [graphs[n] for n in NAMES]
    for rep in range(3):
        obs_list = [synth.make_obs(1, len(M.ALL[n]), seed=100 * rep + i)[0].numpy() for i, n in enumerate(NAMES)]
        acts = ag.select_actions(obs_list, glist)
        assert len(acts) == len(NAMES)
        for n, o, a in zip(NAMES, obs_list, acts):
            assert a.shape == (1, 3 * len(M.ALL[n]))
            assert parity.rel_err(a, _oracle_actions(pa, o, graphs[n])) < parity.RTOL
            ag.change_morphology(graphs[n])                 # the reference's env-by-env loop (src/trainer.py:174-196)
            assert parity.rel_err(a, ag.select_action(o)) < 3e-5       # packed batch: tcgen05 3xTF32 projections; B=1: fp32 SIMT
    with pytest.raises(ValueError):
        ag.select_actions(obs_list[:2], glist)
    with pytest.raises(RuntimeError, match="limbs x 41"):
        ag.select_actions([obs_list[1]] + obs_list[1:], glist)


def test_rollout_graph_sees_updated_weights():
    ag, _, _ = make_agent()
    par = M.ALL["3d_walker_7_full"]
    g = G.build_graph(par, device="cuda")
    ag.change_morphology(g)
    obs = synth.make_obs(64, len(par), seed=9).numpy()     # 448 tokens: the tcgen05 path with its pre-split weights
    for _ in range(3):
        before = ag.select_action(obs)
    b = {k: v.cuda() for k, v in synth.make_batch(32, len(par), seed=2).items()}
    for it in range(2):
        ag.update(b, it)
    after = ag.select_action(obs)                           # replayed graph, weights changed by the fused Adam
    with torch.no_grad():
        want = ag.actor(torch.tensor(obs).cuda()).cpu().numpy()
    assert parity.rel_err(after, want) < 1e-6
    assert parity.rel_err(after, before) > 1e-5
    sd = {k: v.clone() for k, v in ag.state_dict().items()}
    ag.load_state_dict({k: (v * 0.5 if k.startswith("actor.") and "decoder_g" in k else v) for k, v in sd.items()})   # edit behind our back
    half = ag.select_action(obs)
    with torch.no_grad():
        want = ag.actor(torch.tensor(obs).cuda()).cpu().numpy()
    assert parity.rel_err(half, want) < 1e-6


def test_large_batch_forward_two_cta_gemm_variant():
    """At rollout sizes (several 128-row tiles per SM) the short-K projections of an inference pass run in the
    two-CTAs-per-SM tcgen05 variant with ONE TMEM accumulator for all three 3xTF32 terms: same 1e-4 bar."""
    ag, pa, _ = make_agent()
    par = M.ALL["3d_walker_7_full"]
    g = G.build_graph(par, device="cuda")
    ag.change_morphology(g)
    obs = synth.make_obs(8192, len(par), seed=21).cuda()              # 57 344 tokens: 448 row tiles
    with torch.no_grad():
        got = ag.actor(obs)
        want = O.actor_forward({k: v.cuda() for k, v in pa.items()}, obs, g)
    err = parity.rel_err(got, want)
    print(f"rollout forward at 57k tokens: rel err {err:.2e}")
    assert err < parity.RTOL
    rot = synth.rotate_about_gravity(obs.cpu(), len(par), 0.7).cuda()
    with torch.no_grad():
        err_rot = parity.rel_err(ag.actor(rot), got)
    print(f"rotation about gravity at 57k tokens: rel err {err_rot:.2e}")
    assert err_rot < parity.RTOL_ROT
