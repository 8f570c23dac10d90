"""Parity at the sizes and shapes the first round left out ("What's weak" 1-2):

* gradients (critic and actor) at the benched configuration, humanoid-9 B=256, and on the largest morphology
  (3d_cheetah_14_full), against the fp64 oracle;
* the limb-count limits: a 15-limb chain (the positional tables hold 15 rows, src/SEActor.py:19), a 2-limb pair, and the
  error for 16 limbs;
* the forward of every morphology of sgrl_b200/morphologies.py;
* a 20-seed sweep of weights and batches that prints the pass fraction and attributes every miss to relu units whose
  pre-activation sits on the kink (|a| < 1e-5 of its row's scale in fp64), i.e. whose side is decided by fp32 summation order.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import set_oracle as O
from sgrl_b200 import graph as G, morphologies as M, synth
import gpu_util
import parity
from test_backward_gpu import grad_report, oracle_grads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    return gpu_util.make_modules(use_tc=1)


KINK = 1e-5      # a relu unit whose |activation| is below this fraction of its row's largest sits "on the kink"


class Verdict:
    """Outcome of one gradient comparison.  ok: within 1e-4 of the fp64 oracle.  Otherwise `explained` says whether the miss is
    fully accounted for by relu units on their kink: every unit that the CUDA forward and the fp64 forward put on different
    sides of zero has |activation| < KINK of its row scale, AND the gradients are within 1e-4 of the fp64 oracle evaluated
    with the CUDA forward's activation pattern imposed (oracle `masks=`)."""

    def __init__(self, what, bad, glob):
        self.what, self.bad, self.glob = what, bad, glob
        self.ok = not bad and glob < parity.RTOL
        self.flips, self.bad_forced, self.glob_forced = [], None, None

    @property
    def explained(self):
        return (not self.ok and self.flips and max(r for _, r in self.flips) < KINK and self.bad_forced is not None
                and not self.bad_forced and self.glob_forced < parity.RTOL)

    def __str__(self):
        worst = sorted(self.bad, key=lambda x: -x[1])[:6]
        s = f"{self.what}: global {self.glob:.2e}, {len(self.bad)} tensors out {worst}"
        if not self.ok:
            s += f"; {len(self.flips)} relu units on the other side of zero {[(k, f'{r:.1e}') for k, r in self.flips[:6]]}"
            if self.bad_forced is not None:
                s += f"; against the oracle with the CUDA activation pattern: global {self.glob_forced:.2e}, {len(self.bad_forced)} tensors out"
        return s


def _traces(params64, prefixes, x, g64):
    out = {}
    for pre in prefixes:
        tr = {}
        with torch.no_grad():
            O.transformer_model(O.sub(params64, pre), x, g64, trace=tr)
        out[pre] = tr
    return out


def _critic_case(critic, pc, g, b):
    critic.change_morphology(g)
    critic.zero_grad(set_to_none=True)
    q1, q2 = critic(b["obs"], b["action"])
    tgt = b["reward"].expand_as(q1)
    loss = F.mse_loss(q1, tgt) + F.mse_loss(q2, tgt)
    loss.backward()
    got = {k: p.grad for k, p in critic.named_parameters()}
    g64 = _g64(g)

    def lf(p, masks=None):
        o1, o2 = O.critic_forward(p, b["obs"].double(), b["action"].double(), g64, masks=masks)
        t = b["reward"].double().expand_as(o1)
        return F.mse_loss(o1, t) + F.mse_loss(o2, t)
    want_loss, want = oracle_grads(pc, lf)
    assert abs(loss.item() - want_loss) < 1e-5 * abs(want_loss)
    v = Verdict("critic", *grad_report(got, want))
    if not v.ok:
        B, N = b["obs"].shape[0], len(g["parents"])
        tb = critic._tables(B)
        _, stash = critic.forward_raw(tb, b["obs"].contiguous(), b["action"].contiguous(), keep=True)
        torch.cuda.synchronize()
        x = torch.cat([b["obs"].view(B, N, 41), b["action"].view(B, N, 3)], 2).double()
        tr = _traces({k: t.cuda().double() for k, t in pc.items()}, ("critic1.", "critic2."), x, g64)
        masks = {}
        for z, pre in enumerate(("critic1.", "critic2.")):
            v.flips += gpu_util.relu_flips(critic, stash, tb, 2, z, tr[pre], pre)
            masks[z + 1] = gpu_util.relu_masks(critic, stash, tb, 2, z)
        _, want2 = oracle_grads(pc, lambda p: lf(p, masks))
        v.bad_forced, v.glob_forced = grad_report(got, want2)
    return v


def _actor_case(actor, critic, pa, pc, g, b):
    actor.change_morphology(g); critic.change_morphology(g)
    actor.zero_grad(set_to_none=True); critic.zero_grad(set_to_none=True)
    aloss = -critic.Q1(b["obs"], actor(b["obs"])).mean()
    aloss.backward()
    got = {k: p.grad for k, p in actor.named_parameters()}
    g64 = _g64(g)
    pc64 = {k: t.cuda().double() for k, t in pc.items()}

    def lf(p, mk_a=None, mk_c=None):
        a = O.actor_forward(p, b["obs"].double(), g64, masks=mk_a)
        return -O.critic_forward(pc64, b["obs"].double(), a, g64, which=(1,), masks=None if mk_c is None else {1: mk_c}).mean()
    want_loss, want = oracle_grads(pa, lf)
    assert abs(aloss.item() - want_loss) < 1e-4 * abs(want_loss)
    v = Verdict("actor", *grad_report(got, want))
    if not v.ok:
        B, N = b["obs"].shape[0], len(g["parents"])
        tb = actor._tables(B)
        obs = b["obs"].contiguous()
        a_out, st_a = actor.forward_raw(tb, obs, None, keep=True, nb=1)
        _, st_c = critic.forward_raw(tb, obs, a_out[0].contiguous(), keep=True, nb=1)
        torch.cuda.synchronize()
        pa64 = {k: t.cuda().double() for k, t in pa.items()}
        xa = b["obs"].view(B, N, 41).double()
        with torch.no_grad():
            a_ref = O.actor_forward(pa64, b["obs"].double(), g64)
        xc = torch.cat([xa, a_ref.view(B, N, 3)], 2)
        v.flips += gpu_util.relu_flips(actor, st_a, tb, 1, 0, _traces(pa64, ("actor.",), xa, g64)["actor."], "actor.")
        v.flips += gpu_util.relu_flips(critic, st_c, tb, 1, 0, _traces(pc64, ("critic1.",), xc, g64)["critic1."], "critic1.")
        mk_a, mk_c = gpu_util.relu_masks(actor, st_a, tb, 1, 0), gpu_util.relu_masks(critic, st_c, tb, 1, 0)
        _, want2 = oracle_grads(pa, lambda p: lf(p, mk_a, mk_c))
        v.bad_forced, v.glob_forced = grad_report(got, want2)
    return v


def _accept(v, report):
    """within tolerance, or a miss fully explained by relu units on their kink (printed either way when it was a miss)"""
    if not v.ok:
        report.append(str(v))
    assert v.ok or v.explained, str(v)


@pytest.mark.parametrize("name,B", [("3d_humanoid_9_full", 256), ("3d_cheetah_14_full", 3), ("3d_cheetah_14_full", 64)])
def test_gradients_headline_and_largest_morphology(mods, capsys, name, B):
    actor, critic, pa, pc = mods
    par = M.ALL[name]
    g = G.build_graph(par, device="cuda")
    b = gpu_util.to_cuda(synth.make_batch(B, len(par), seed=1))
    report = []
    _accept(_critic_case(critic, pc, g, b), report)
    _accept(_actor_case(actor, critic, pa, pc, g, b), report)
    if report:
        with capsys.disabled():
            print("\n  " + "\n  ".join(report))


@pytest.mark.parametrize("parents,B", [([-1] + list(range(14)), 32), ([-1, 0], 64), ([-1, 0, 1, 2, 2, 1, 5, 5, 0, 8, 9, 9, 8, 12, 12], 16)],
                         ids=["chain15", "pair2", "bushy15"])
def test_limb_count_limits(mods, capsys, parents, B):
    actor, critic, pa, pc = mods
    g = G.build_graph(parents, device="cuda")
    N = len(parents)
    b = gpu_util.to_cuda(synth.make_batch(B, N, seed=3))
    actor.change_morphology(g); critic.change_morphology(g)
    with torch.no_grad():
        a = actor(b["obs"])
        q1, q2 = critic(b["obs"], b["action"])
    pa_c = {k: v.cuda() for k, v in pa.items()}
    pc_c = {k: v.cuda() for k, v in pc.items()}
    with torch.no_grad():
        a_ref = O.actor_forward(pa_c, b["obs"], g)
        q1_ref, q2_ref = O.critic_forward(pc_c, b["obs"], b["action"], g)
    assert a.shape == (B, 3 * N) and q1.shape == q1_ref.shape
    assert parity.rel_err(a, a_ref) < parity.RTOL
    assert parity.rel_err(q1, q1_ref) < parity.RTOL and parity.rel_err(q2, q2_ref) < parity.RTOL
    report = []
    _accept(_critic_case(critic, pc, g, b), report)
    _accept(_actor_case(actor, critic, pa, pc, g, b), report)
    if report:
        with capsys.disabled():
            print("\n  " + "\n  ".join(report))


def test_sixteen_limbs_are_rejected_like_the_reference(mods):
    """A 16-limb graph has traversal ranks up to 15, one past the 15-row positional tables: the reference raises an index
    error inside nn.Embedding (src/SEActor.py:19,34-38); the drop-in refuses the morphology up front."""
    actor = mods[0]
    g = G.build_graph([-1] + list(range(15)), device="cuda")
    actor.change_morphology(g)
    with pytest.raises((ValueError, IndexError, RuntimeError)):
        actor(torch.zeros(2, 16 * 41, device="cuda"))


@pytest.mark.parametrize("name", sorted(M.ALL))
def test_forward_every_morphology(mods, name):
    actor, critic, pa, pc = mods
    par = M.ALL[name]
    if len(par) < 2:
        pytest.skip("single-limb morphology: the reference builds no traversals for it (utils.py:452-453)")
    g = G.build_graph(par, device="cuda")
    B = 64
    b = gpu_util.to_cuda(synth.make_batch(B, len(par), seed=5))
    actor.change_morphology(g); critic.change_morphology(g)
    with torch.no_grad():
        a = actor(b["obs"])
        q1, q2 = critic(b["obs"], b["action"])
        a_ref = O.actor_forward({k: v.cuda().double() for k, v in pa.items()}, b["obs"].double(), _g64(g))
        q1_ref, q2_ref = O.critic_forward({k: v.cuda().double() for k, v in pc.items()}, b["obs"].double(), b["action"].double(), _g64(g))
    assert parity.rel_err(a, a_ref) < parity.RTOL, name
    assert parity.rel_err(q1, q1_ref) < parity.RTOL and parity.rel_err(q2, q2_ref) < parity.RTOL, name


def _g64(g):
    g64 = dict(g); g64["relation"] = g["relation"].double()
    return g64


def test_seed_sweep_reports_pass_fraction_and_attributes_misses_to_kinks(capsys):
    """20 (weight seed, batch seed) pairs at humanoid-9, B=32, critic and actor gradients.  A pair passes when every gradient
    tensor is within 1e-4 of the fp64 oracle.  Every miss must be fully explained by relu units on their kink (Verdict):
    such a unit's side is decided by fp32 summation order — the reference's own fp32 forward has the same freedom — and with
    the CUDA forward's activation pattern imposed on the fp64 oracle the gradients agree to 1e-4 again.  A miss that is not
    explained this way fails the test; the pass fraction is printed."""
    par = M.ALL["3d_humanoid_9_full"]
    g = G.build_graph(par, device="cuda")
    n, clean, report = 20, 0, []
    for s in range(n):
        actor, critic, pa, pc = gpu_util.make_modules(seed=100 + 7 * s, use_tc=1)
        b = gpu_util.to_cuda(synth.make_batch(32, len(par), seed=200 + s))
        vc, va = _critic_case(critic, pc, g, b), _actor_case(actor, critic, pa, pc, g, b)
        clean += vc.ok and va.ok
        for v in (vc, va):
            if not v.ok:
                report.append(f"pair {s} {v}")
            assert v.ok or v.explained, f"seed pair {s}: {v}"
    with capsys.disabled():
        print(f"\nseed sweep: {clean}/{n} pairs within 1e-4 of the fp64 oracle on every gradient tensor; "
              f"{n - clean} explained by relu units on their kink")
        for r in report:
            print("  " + r[:400])
