"""Pin the oracle: restatement vs golden vectors (everywhere) and vs the unmodified
reference modules (when /root/reference is on this machine)."""
import numpy as np
import pytest
import torch

from oracle import ref_loader, set_oracle as O
from sgrl_b200 import morphologies as M, synth
import parity


def _params():
    a = {"actor." + k: v for k, v in O.synth_params("actor", parity.WEIGHT_SEED).items()}
    c = {"critic1." + k: v for k, v in O.synth_params("critic", parity.WEIGHT_SEED + 1).items()}
    c.update({"critic2." + k: v for k, v in O.synth_params("critic", parity.WEIGHT_SEED + 2).items()})
    return a, c


@pytest.fixture(scope="module")
def gold():
    return parity.load_golden()


@pytest.fixture(scope="module")
def params():
    return _params()


def test_param_spec_counts():
    na = sum(int(np.prod(s)) for _, s in O.param_spec("actor"))
    nc = sum(int(np.prod(s)) for _, s in O.param_spec("critic"))
    assert (na, 2 * nc) == (4712712, 8761330)
    live = sum(int(np.prod(s)) for n, s in O.param_spec("actor") if not O.is_dead(n))
    assert live == 4514568
    assert len(O.param_spec("actor")) == 139 and len(O.param_spec("critic")) == 135


@pytest.mark.parametrize("name,B", parity.CASES)
def test_forward_against_golden(gold, params, name, B):
    a, c = params
    g = parity.golden_graph(gold, name, M.ALL[name])
    b = parity.golden_batch(gold, name)
    with torch.no_grad():
        act = O.actor_forward(a, b["obs"], g)
        q1, q2 = O.critic_forward(c, b["obs"], b["action"], g)
    assert parity.rel_err(act, gold[name + "/actions"]) < 2e-5
    assert parity.rel_err(q1, gold[name + "/q1"]) < 2e-5
    assert parity.rel_err(q2, gold[name + "/q2"]) < 2e-5


@pytest.mark.parametrize("name,B", parity.CASES[:4])
def test_gradients_against_golden(gold, params, name, B):
    a, c = params
    g = parity.golden_graph(gold, name, M.ALL[name])
    b = parity.golden_batch(gold, name)
    cc = {k: v.clone().requires_grad_(not O.is_dead(k)) for k, v in c.items()}
    q1, q2 = O.critic_forward(cc, b["obs"], b["action"], g)
    tgt = b["reward"].expand_as(q1)
    loss = torch.nn.functional.mse_loss(q1, tgt) + torch.nn.functional.mse_loss(q2, tgt)
    loss.backward()
    assert abs(loss.item() - float(gold[name + "/closs0"])) < 1e-5 * abs(float(gold[name + "/closs0"]))
    parity.check_summary(parity.summarize({k: v.grad for k, v in cc.items()}), gold[name + "/critic_grad"], what="critic grad")

    aa = {k: v.clone().requires_grad_(not O.is_dead(k)) for k, v in a.items()}
    aloss = -O.critic_forward(c, b["obs"], O.actor_forward(aa, b["obs"], g), g, which=(1,)).mean()
    aloss.backward()
    assert abs(aloss.item() - float(gold[name + "/aloss0"])) < 2e-5 * abs(float(gold[name + "/aloss0"]))
    # golden actor summaries are keyed like SEPolicy.named_parameters(): 'actor.<name>' == keys of aa
    parity.check_summary(parity.summarize({k: v.grad for k, v in aa.items()}), gold[name + "/actor_grad"], what="actor grad")

    act_in = b["action"].clone().requires_grad_(True)
    O.critic_forward(c, b["obs"], act_in, g, which=(1,)).mean().backward()
    assert parity.rel_err(act_in.grad, gold[name + "/dq1_daction"]) < 1e-4


@pytest.mark.parametrize("name,B", [parity.CASES[1], parity.CASES[3]])
def test_td3_update_against_golden(gold, params, name, B):
    a, c = params
    g = parity.golden_graph(gold, name, M.ALL[name])
    b = parity.golden_batch(gold, name)
    td3 = O.TD3Oracle(a, c)
    names = [str(s) for s in gold["state_names"]]
    for it in range(2):
        before = _agent_state(td3)
        out = td3.update(b, it, torch.tensor(gold[name + "/noise"][it]), g)
        assert abs(out["loss/critic_loss"].item() - float(gold[name + f"/upd{it}/critic_loss"])) < 2e-5 * float(gold[name + f"/upd{it}/critic_loss"])
        if it == 0:
            assert abs(out["loss/actor_loss"].item() - float(gold[name + "/upd0/actor_loss"])) < 1e-4 * abs(float(gold[name + "/upd0/actor_loss"]))
        assert abs(out["misc/train_reward_var"] - float(gold[name + f"/upd{it}/reward_var"])) < 1e-6
        after = _agent_state(td3)
        step = {n: after[n] - before[n] for n in names}
        # Adam steps are ~lr*sign(g): compare step summaries with a looser floor (eps=1e-8 dominates tiny grads)
        parity.check_summary(parity.summarize(step), gold[name + f"/upd{it}/step"], rtol=2e-3, floor=1e-3, what=f"step it={it}")
    with torch.no_grad():
        assert parity.rel_err(O.actor_forward(td3.actor, b["obs"], g), gold[name + "/actions_after"]) < 1e-4
        assert parity.rel_err(O.critic_forward(td3.critic_t, b["obs"], b["action"], g, which=(1,)), gold[name + "/tq1_after"]) < 1e-4


def _agent_state(td3):
    s = {}
    for pre, d in (("actor.", td3.actor), ("actor_target.", td3.actor_t), ("critic.", td3.critic), ("critic_target.", td3.critic_t)):
        s.update({pre + k: v.detach().clone() for k, v in d.items()})
    return s


def test_rotation_invariance_of_oracle(params):
    a, c = params
    from sgrl_b200 import graph as G
    par = M.ALL["3d_humanoid_9_full"]
    g = G.build_graph(par)
    b = synth.make_batch(6, len(par), seed=5)
    rot = synth.rotate_about_gravity(b["obs"], len(par), 0.7)
    with torch.no_grad():
        assert parity.rel_err(O.actor_forward(a, rot, g), O.actor_forward(a, b["obs"], g)) < parity.RTOL_ROT
        assert parity.rel_err(O.critic_forward(c, rot, b["action"], g)[0], O.critic_forward(c, b["obs"], b["action"], g)[0]) < parity.RTOL_ROT


@pytest.mark.skipif(ref_loader.find_reference() is None, reason="reference not on this machine")
def test_oracle_against_live_reference():
    ref = ref_loader.load_reference()
    torch.manual_seed(0)
    ag = ref.agent.Agent(ref_loader.default_args())
    par = M.ALL["3d_humanoid_8_left_knee"]
    g = ref.utils.getGraphDict(par, ["pre", "inlcrs", "postlcrs"], device=torch.device("cpu"))
    ag.change_morphology(g)
    sd = ag.state_dict()
    b = synth.make_batch(6, len(par), seed=9)
    pa = {k[len("actor."):]: v for k, v in sd.items() if k.startswith("actor.")}
    pc = {k[len("critic."):]: v for k, v in sd.items() if k.startswith("critic.")}
    with torch.no_grad():
        assert parity.rel_err(O.actor_forward(pa, b["obs"], g), ag.actor(b["obs"])) < 1e-5
        q = ag.critic(b["obs"], b["action"])
        qo = O.critic_forward(pc, b["obs"], b["action"], g)
        assert parity.rel_err(qo[0], q[0]) < 1e-5 and parity.rel_err(qo[1], q[1]) < 1e-5
    with pytest.raises(AssertionError):
        O.critic_forward(pc, b["obs"][:, :-41], b["action"], g)
