#!/usr/bin/env python
"""Device time of each phase of Agent.update in isolation (each phase captured as its own CUDA graph and replayed),
next to the whole captured update.  Shows how much of the step is the dependency chain of one phase and how much the
phases overlap.  python tools/phase_probe.py [--morph M] [--batch B]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgrl_b200 import _lib, graph as G, morphologies as M, synth  # noqa: E402
from sgrl_b200.agent import Agent, _UpdatePlan  # noqa: E402
from sgrl_b200.config import default_args  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--morph", default="3d_humanoid_9_full")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()

torch.cuda.set_device(0)
torch.manual_seed(0)
ag = Agent(default_args())
par = M.ALL[a.morph]
g = G.build_graph(par, device="cuda")
ag.change_morphology(g)
B, N = a.batch, len(par)
b = {k: v.cuda() for k, v in synth.make_batch(B, N, seed=1).items()}
for it in range(4):
    ag.update(b, it)
torch.cuda.synchronize()
tb = ag.actor._tables(B)
p = ag._plan(tb)
T = tb.T
ac, cr, at, ct = ag.actor, ag.critic, ag.actor_target, ag.critic_target


def timed(name, fn, reps=a.reps):
    fn(); torch.cuda.synchronize()
    l0 = _lib.lib.sgrl_launch_count()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    n = _lib.lib.sgrl_launch_count() - l0
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:44s} {ms * 1e3:9.1f} us  {n:4d} launches  {ms * 1e3 / max(n, 1):6.1f} us/launch", flush=True)
    return ms


print(f"# {a.morph} B={B} T={T}")
tot = 0.0
tot += timed("actor_target fwd (nb=1, keep=0)", lambda: at.forward_raw(tb, p.nobs, None, keep=False, trusted_split=True, out=p.a_t, stash=p.stash_at))
tot += timed("critic_target fwd (nb=2, keep=0)", lambda: ct.forward_raw(tb, p.nobs, p.next_action, keep=False, nb=2, trusted_split=True, out=p.tq, stash=p.stash_ct))
tot += timed("critic fwd (nb=2, keep=1)", lambda: cr.forward_raw(tb, p.obs, p.act, keep=True, nb=2, trusted_split=True, out=p.q, stash=p.stash_c))
tot += timed("critic bwd (nb=2, wgrad)", lambda: cr.backward_raw(tb, p.stash_c, p.dq, 2, cr.grad_arena(), False, trusted_split=True, ws=p.ws))
tot += timed("critic clip+adam", lambda: ag.critic_optimizer.step(max_norm=0.1))
half = 0.0
half += timed("actor fwd (nb=1, keep=1)", lambda: ac.forward_raw(tb, p.obs, None, keep=True, trusted_split=True, out=p.pi, stash=p.stash_a))
half += timed("critic1 fwd (nb=1, keep=1)", lambda: cr.forward_raw(tb, p.obs, p.pi[0], keep=True, nb=1, trusted_split=True, out=p.q1, stash=p.stash_c))
half += timed("critic1 bwd (nb=1, data only)", lambda: cr.backward_raw(tb, p.stash_c, p.dq1, 1, None, True, trusted_split=True, ws=p.ws, dact=p.dact))
half += timed("actor bwd (nb=1, wgrad)", lambda: ac.backward_raw(tb, p.stash_a, p.dact, 1, ac.grad_arena(), False, trusted_split=True, ws=p.ws))
half += timed("actor clip+adam", lambda: ag.actor_optimizer.step(max_norm=0.1))
half += timed("polyak x2", lambda: ag.try_update_target_network())
print(f"# serial sum: critic-only step {tot:.3f} ms, actor step {tot + half:.3f} ms, policy_freq=2 average {tot + half / 2:.3f} ms")
for step in (False, True):
    ms = timed(f"whole update (actor_step={step})", lambda: ag._update_impl(p, step))


# ---- where the critic-only step's time beyond its critical chain goes: chain A -> loss -> backward -> Adam WITHOUT the concurrent
# critic forward (its outputs are in the plan from the runs above), and the two forward chains alone
def chain_a():
    at.forward_raw(tb, p.nobs, None, keep=False, trusted_split=True, out=p.a_t, stash=p.stash_at)
    ct.forward_raw(tb, p.nobs, p.a_t[0], keep=False, nb=2, trusted_split=True, out=p.tq, stash=p.stash_ct)


def tail():
    cr.backward_raw(tb, p.stash_c, p.dq, 2, cr.grad_arena(), False, trusted_split=True, ws=p.ws)
    ag.critic_optimizer.step(max_norm=0.1)


def fwd_ab():
    main = torch.cuda.current_stream()
    p.ev_start.record(main)
    with torch.cuda.stream(p.s1):
        p.s1.wait_event(p.ev_start)
        chain_a()
        p.ev_a.record(p.s1)
    cr.forward_raw(tb, p.obs, p.act, keep=True, nb=2, trusted_split=True, out=p.q, stash=p.stash_c)
    main.wait_event(p.ev_a)


timed("chain A alone (actor_t -> critic_t)", chain_a)
timed("chain A || critic fwd", fwd_ab)
timed("chain A -> critic bwd -> adam (no critic fwd)", lambda: (chain_a(), tail()))
timed("(chain A || critic fwd) -> critic bwd -> adam", lambda: (fwd_ab(), tail()))
