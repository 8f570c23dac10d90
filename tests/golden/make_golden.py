"""Generate tests/golden/set_golden.npz from the UNMODIFIED reference modules.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

The reference ships no tests or golden files, so the vectors are produced
here by importing its SET modules (oracle/ref_loader.py), loading deterministic
init-scale weights (oracle.set_oracle.synth_params — regenerated bit-identically on any
machine from numpy PCG64, so the 100 MB state need not be stored) and recording, per
morphology case:

* inputs (obs, next_obs, action, reward, done, injected policy noise)
* SEPolicy.forward actions, SECritic.forward Q1/Q2
* per-tensor gradient summaries of the critic loss and of the actor loss
  (L2 norm and dot product with a fixed pseudo-random probe, in param_spec order)
* two consecutive Agent.update calls (it=0 with actor step + Polyak, it=1 without):
  losses, target_Q, and per-tensor summaries of the parameter *steps* of all 4 networks.
"""
from __future__ import annotations

import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader, set_oracle as O  # noqa: E402
from sgrl_b200 import graph as G, morphologies as M, synth  # noqa: E402

CASES = [
    ("3d_walker_2_right_leg_left_knee", 4),
    ("3d_hopper_3_shin", 4),
    ("3d_walker_7_full", 5),
    ("3d_humanoid_9_full", 8),
    ("3d_cheetah_14_full", 3),
]
WEIGHT_SEED = 11


def probe(name: str, shape) -> torch.Tensor:
    rng = np.random.Generator(np.random.PCG64([977, zlib.crc32(name.encode())]))
    return torch.tensor(rng.standard_normal(tuple(shape)), dtype=torch.float64)


def summarize(named: dict) -> np.ndarray:
    """(n_tensors, 2) float64: [L2 norm, <t, probe>] per tensor (zeros for None)."""
    rows = []
    for k, v in named.items():
        if v is None:
            rows.append((0.0, 0.0))
        else:
            v = v.detach().double()
            rows.append((v.norm().item(), (v * probe(k, v.shape)).sum().item()))
    return np.array(rows, dtype=np.float64)


def full_state(seed: int) -> dict:
    """Agent.state_dict()-style dict (actor.actor.*, critic.critic{1,2}.*, targets = copies)."""
    a = O.synth_params("actor", seed)
    c1 = O.synth_params("critic", seed + 1)
    c2 = O.synth_params("critic", seed + 2)
    sd = {}
    for pre in ("actor.actor.", "actor_target.actor."):
        sd.update({pre + k: v.clone() for k, v in a.items()})
    for pre in ("critic.", "critic_target."):
        sd.update({pre + "critic1." + k: v.clone() for k, v in c1.items()})
        sd.update({pre + "critic2." + k: v.clone() for k, v in c2.items()})
    return sd


def main():
    ref = ref_loader.load_reference()
    torch.set_num_threads(8)
    out = {}
    for name, B in CASES:
        par = M.ALL[name]
        N = len(par)
        g = ref.utils.getGraphDict(par, ["pre", "inlcrs", "postlcrs"], device=torch.device("cpu"))
        args = ref_loader.default_args()
        ag = ref.agent.Agent(args)
        ag.load_state_dict(full_state(WEIGHT_SEED))
        ag.change_morphology(g)
        batch = synth.make_batch(B, N, seed=3)
        noises = []
        for it in range(2):
            torch.manual_seed(4242 + it)
            noises.append(torch.zeros_like(batch["action"]).normal_(0, args.policy_noise))
        k = name + "/"
        for f in ("obs", "next_obs", "action", "reward", "done"):
            out[k + f] = batch[f].numpy()
        out[k + "noise"] = torch.stack(noises).numpy()
        out[k + "relation"] = g["relation"].numpy()
        out[k + "traversals"] = torch.stack(g["traversals"]).numpy()

        with torch.no_grad():
            out[k + "actions"] = ag.actor(batch["obs"]).numpy()
            q1, q2 = ag.critic(batch["obs"], batch["action"])
            out[k + "q1"], out[k + "q2"] = q1.numpy(), q2.numpy()

        # gradients of the two losses at the initial weights, target = reward broadcast (no bootstrap)
        ag.critic.zero_grad(); ag.actor.zero_grad()
        q1, q2 = ag.critic(batch["obs"], batch["action"])
        tgt = batch["reward"].expand_as(q1)
        closs = torch.nn.functional.mse_loss(q1, tgt) + torch.nn.functional.mse_loss(q2, tgt)
        closs.backward()
        out[k + "closs0"] = np.float64(closs.item())
        out[k + "critic_grad"] = summarize({n: p.grad for n, p in ag.critic.named_parameters()})
        ag.critic.zero_grad(); ag.actor.zero_grad()
        act_in = batch["action"].clone().requires_grad_(True)
        ag.critic.Q1(batch["obs"], act_in).mean().backward()
        out[k + "dq1_daction"] = act_in.grad.numpy()
        ag.critic.zero_grad()
        aloss = -ag.critic.Q1(batch["obs"], ag.actor(batch["obs"])).mean()
        aloss.backward()
        out[k + "aloss0"] = np.float64(aloss.item())
        out[k + "actor_grad"] = summarize({n: p.grad for n, p in ag.actor.named_parameters()})
        ag.critic.zero_grad(); ag.actor.zero_grad()

        # two reference TD3 updates
        for it in range(2):
            before = {n: v.detach().clone() for n, v in ag.state_dict().items()}
            torch.manual_seed(4242 + it)
            ld = ag.update(batch, it)
            out[k + f"upd{it}/critic_loss"] = np.float64(ld["loss/critic_loss"].item())
            if "loss/actor_loss" in ld:
                out[k + f"upd{it}/actor_loss"] = np.float64(ld["loss/actor_loss"].item())
            out[k + f"upd{it}/reward_mean"] = np.float64(ld["misc/train_reward_mean"])
            out[k + f"upd{it}/reward_var"] = np.float64(ld["misc/train_reward_var"])
            after = ag.state_dict()
            out[k + f"upd{it}/step"] = summarize({n: after[n] - before[n] for n in after})
        with torch.no_grad():
            out[k + "actions_after"] = ag.actor(batch["obs"]).numpy()
            q1, q2 = ag.critic_target(batch["obs"], batch["action"])
            out[k + "tq1_after"] = q1.numpy()
        print(name, "B", B, "closs0", out[k + "closs0"], "aloss0", out[k + "aloss0"])
    names = list(full_state(WEIGHT_SEED).keys())
    out["state_names"] = np.array(names)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "set_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
