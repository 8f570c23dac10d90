"""TD3 agent over the SET modules — drop-in for the reference ``agent.Agent`` with
``actor_type == critic_type == 'set'`` (src/agent.py:20-217): same constructor argument,
``update(data_batch, it)``, ``select_action``, ``change_morphology``, ``models2train/eval``,
``try_update_target_network`` and attributes; it is an ``nn.Module`` whose ``state_dict()``
has the reference's 818 keys (src/common/agents.py:7-11, src/common/trainer.py:256-259).

``update`` does not go through autograd: it drives the forward/backward kernels on the flat
arenas directly and finishes with the fused clip+Adam pass and (every ``policy_freq``-th
call) the Polyak pass.  Under ``torch.distributed`` (one process per GPU, NCCL) the flat
gradient arena is all-reduced once per optimizer step, so every rank applies the identical
update (data parallel over replay samples / morphologies).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from ._lib import check, lib, ptr, stream
from .modules import SECritic, SEPolicy, SetNetModule


class FusedAdam:
    """torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8) + clip_grad_norm_ over a module's flat
    live arena in one kernel pass (src/agent.py:104-105,152-156,172-176)."""

    def __init__(self, module: SetNetModule, lr: float, betas=(0.9, 0.999), eps: float = 1e-8):
        self.module = module
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self._alloc()

    def _alloc(self):
        live = self.module.live_arena
        self.exp_avg = torch.zeros_like(live)
        self.exp_avg_sq = torch.zeros_like(live)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=live.device)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=live.device)
        self._arena_ptr = live.data_ptr()

    def _sync_device(self):
        if self.module.live_arena.data_ptr() != self._arena_ptr:      # module was moved / re-flattened
            old = (self.exp_avg, self.exp_avg_sq, self.step_count)
            self._alloc()
            self.exp_avg.copy_(old[0]); self.exp_avg_sq.copy_(old[1]); self.step_count.copy_(old[2])

    def zero_grad(self, set_to_none: bool = False):
        self._sync_device()
        self.module.grad_arena().zero_()

    def step(self, max_norm: float = 0.0, world_size: int = 1):
        """Clip (if max_norm > 0) and apply one Adam step from module.grad_arena()."""
        self._sync_device()
        m = self.module
        g, p = m.grad_arena(), m.live_arena
        n, st = p.numel(), stream()
        self.sumsq.zero_()
        if max_norm > 0:
            check(lib.sgrl_sumsq(ptr(g), n, ptr(self.sumsq), st), "sgrl_sumsq")
        check(lib.sgrl_bump_step(ptr(self.step_count), st), "sgrl_bump_step")
        fresh = m._split is not None and m._split_fresh and m._split_version == m._arena._version
        hi, lo = (m._split[0], m._split[1]) if fresh else (None, None)     # keep a valid tf32 split valid (else it is rebuilt lazily)
        check(lib.sgrl_adam_clip(ptr(p), ptr(g), ptr(self.exp_avg), ptr(self.exp_avg_sq), n, ptr(self.sumsq), ptr(self.step_count),
                                 self.lr, self.betas[0], self.betas[1], self.eps, float(max_norm), 1.0 / world_size, ptr(hi), ptr(lo), st),
              "sgrl_adam_clip")

    def state_dict(self):
        return {"step": self.step_count.clone(), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "lr": self.lr, "betas": self.betas, "eps": self.eps}

    def load_state_dict(self, sd):
        self._sync_device()
        self.step_count.copy_(sd["step"]); self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.lr, self.betas, self.eps = sd["lr"], tuple(sd["betas"]), sd["eps"]


def soft_update_network(source: SetNetModule, target: SetNetModule, tau: float):
    """theta_t <- tau*theta + (1-tau)*theta_t over ALL parameters incl. the dead ones
    (src/common/functional.py:7-10) as one pass over the flat arenas."""
    s, t = source.full_arena, target.full_arena
    if s.device.type != "cuda":
        with torch.no_grad():
            t.mul_(1 - tau).add_(s, alpha=tau)
        return
    fresh = target._split is not None and target._split_fresh and target._split_version == target._arena._version
    hi, lo = (target._split[0], target._split[1]) if fresh else (None, None)
    check(lib.sgrl_polyak(ptr(t), ptr(s), s.numel(), float(tau), ptr(hi), ptr(lo), target.live_arena.numel(), stream()), "sgrl_polyak")


class Agent(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        if args.actor_type != "set" or args.critic_type != "set":
            raise NotImplementedError("sgrl_b200 implements the SET actor/critic only (actor_type=critic_type='set')")

        def mk_actor():
            return SEPolicy(args.limb_obs_size, args.limb_action_size, args.msg_dim, args.batch_size, args.max_action,
                            args.max_children, args.disable_fold, args.td, args.bu, args)

        def mk_critic():
            return SECritic(args.limb_obs_size, args.limb_action_size, args.msg_dim, args.batch_size,
                            args.max_children, args.disable_fold, args.td, args.bu, args)

        self.actor, self.actor_target = mk_actor(), mk_actor()
        self.critic, self.critic_target = mk_critic(), mk_critic()
        # sync network parameters (agent.py:100-101, tau = 1.0)
        with torch.no_grad():
            self.actor_target.full_arena.copy_(self.actor.full_arena)
            self.critic_target.full_arena.copy_(self.critic.full_arena)
        self.actor_optimizer = FusedAdam(self.actor, lr=args.lr)
        self.critic_optimizer = FusedAdam(self.critic, lr=args.lr)
        self.models2eval()
        self.tot_update_count = 0
        self.target_smoothing_tau = args.agent.target_smoothing_tau
        self.reward_scale = args.agent.reward_scale
        self.lazy_stats = False      # True: reward statistics returned as 0-dim tensors (no host sync)
        self._loss = None

    # ------------------------------------------------------------------ TD3 step
    def _world(self):
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def _allreduce(self, g: torch.Tensor, world: int):
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(g, op=dist.ReduceOp.SUM)

    def update(self, data_batch: Dict, it: int, noise: Optional[torch.Tensor] = None):
        """One TD3 update (src/agent.py:117-183).  data_batch: obs (B,41N), action (B,3N),
        next_obs, reward (B,1), done (B,1) — torch tensors (any device) or numpy arrays.
        `noise` (B,3N) optionally injects the target-policy noise (drawn on device otherwise)."""
        a = self.args
        dev = self.actor.full_arena.device
        reward_in = data_batch["reward"]
        t = {k: _to_dev(data_batch[k], dev) for k in ("obs", "action", "next_obs", "reward", "done")}
        obs, act, nobs, rew, done = t["obs"], t["action"], t["next_obs"], t["reward"].reshape(-1), t["done"].reshape(-1)
        B = obs.shape[0]
        tb = self.actor._tables(B)
        T, st = tb.T, stream()
        world = self._world()
        if noise is None:
            noise = torch.randn(B, act.shape[1], device=dev) * a.policy_noise                    # agent.py:128
        else:
            noise = _to_dev(noise, dev)
        # ---- target:  y = r + (1-d) * gamma * min_i Q_i'(s', clip(pi'(s') + clip(eps)))       agent.py:127-139
        a_t, _ = self.actor_target.forward_raw(tb, nobs, None, keep=False, trusted_split=True)
        next_action = torch.empty(T, 3, device=dev)
        check(lib.sgrl_td3_smooth_action(ptr(a_t), ptr(noise), ptr(next_action), float(a.noise_clip), float(a.max_action), T * 3, st))
        tq, _ = self.critic_target.forward_raw(tb, nobs, next_action, keep=False, nb=2, trusted_split=True)
        # ---- critic step                                                                       agent.py:142-156
        q, stash = self.critic.forward_raw(tb, obs, act, keep=True, nb=2, trusted_split=True)
        scal = torch.zeros(2, device=dev)            # [critic_loss, actor_loss]
        target = torch.empty(T, device=dev)
        dq = torch.empty(2, T, 1, device=dev)
        check(lib.sgrl_td3_critic_loss(ptr(q[0]), ptr(q[1]), ptr(tq[0]), ptr(tq[1]), ptr(rew), ptr(done), ptr(tb.tok_graph), ptr(target),
                                       ptr(dq[0]), ptr(dq[1]), ptr(scal), float(a.discount), float(self.reward_scale), T, st))
        self.critic_optimizer.zero_grad()
        self.critic.backward_raw(tb, stash, dq, 2, self.critic.grad_arena(), False, trusted_split=True)
        del stash
        self._allreduce(self.critic.grad_arena(), world)
        self.critic_optimizer.step(max_norm=float(a.grad_clipping_value), world_size=world)
        loss_dict = {"loss/critic_loss": scal[0]}
        loss_dict.update(self._reward_stats(reward_in, rew))
        # ---- delayed actor step + Polyak                                                       agent.py:165-180
        if it % a.policy_freq == 0:
            pi, stash_a = self.actor.forward_raw(tb, obs, None, keep=True, trusted_split=True)
            q1, stash_c = self.critic.forward_raw(tb, obs, pi[0], keep=True, nb=1, trusted_split=True)
            dq1 = torch.empty(1, T, 1, device=dev)
            check(lib.sgrl_td3_actor_loss(ptr(q1), ptr(dq1), ptr(scal[1:]), T, st))
            dact = self.critic.backward_raw(tb, stash_c, dq1, 1, None, True, trusted_split=True)     # only d/d(action) is needed
            self.actor_optimizer.zero_grad()
            self.actor.backward_raw(tb, stash_a, dact, 1, self.actor.grad_arena(), False, trusted_split=True)
            del stash_a, stash_c
            self._allreduce(self.actor.grad_arena(), world)
            self.actor_optimizer.step(max_norm=float(a.grad_clipping_value), world_size=world)
            self.try_update_target_network()
            loss_dict["loss/actor_loss"] = scal[1]
        self.tot_update_count += 1
        self._last_target = target
        return loss_dict

    train_step = update   # BASELINE.json calls the TD3 step "train()"; the reference name is update (agent.py:117)

    def _reward_stats(self, reward_in, rew_dev):
        """agent.py:158-162 returns Python floats (two device syncs in the reference).  When the batch
        arrived from host memory the statistics are computed there and nothing synchronises."""
        s = self.reward_scale
        if isinstance(reward_in, np.ndarray) or (torch.is_tensor(reward_in) and not reward_in.is_cuda):
            r = torch.as_tensor(reward_in, dtype=torch.float32).reshape(-1) * s
            return {"misc/train_reward_mean": r.mean().item(), "misc/train_reward_var": r.var().item() if r.numel() > 1 else float("nan")}
        r = rew_dev * s
        if self.lazy_stats:
            return {"misc/train_reward_mean": r.mean(), "misc/train_reward_var": r.var()}
        return {"misc/train_reward_mean": r.mean().item(), "misc/train_reward_var": r.var().item()}

    def try_update_target_network(self):
        soft_update_network(self.critic, self.critic_target, self.target_smoothing_tau)
        soft_update_network(self.actor, self.actor_target, self.target_smoothing_tau)

    # ------------------------------------------------------------------ acting
    @torch.no_grad()
    def select_action(self, obs, deterministic=False):
        """src/agent.py:189-198: numpy (41N,) or (B,41N) -> numpy (B,3N)."""
        if len(obs.shape) == 1:
            obs = obs[None,]
        if not isinstance(obs, torch.Tensor):
            obs = torch.as_tensor(np.asarray(obs), dtype=torch.float32)
        obs = obs.to(self.actor.full_arena.device, non_blocking=True)
        return self.actor(obs).cpu().numpy()

    def change_morphology(self, graph):
        self.actor.change_morphology(graph)
        self.actor_target.change_morphology(graph)
        self.critic.change_morphology(graph)
        self.critic_target.change_morphology(graph)

    def models2eval(self):
        self.actor = self.actor.eval()
        self.actor_target = self.actor_target.eval()
        self.critic = self.critic.eval()
        self.critic_target = self.critic_target.eval()

    def models2train(self):
        self.actor = self.actor.train()
        self.actor_target = self.actor_target.train()
        self.critic = self.critic.train()
        self.critic_target = self.critic_target.train()


def _to_dev(x, dev):
    if not torch.is_tensor(x):
        x = torch.as_tensor(np.asarray(x), dtype=torch.float32)
        if dev.type == "cuda":
            x = x.pin_memory()
    return x.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
