"""Synthetic replay batches with the per-limb observation layout of the reference
environments (41 floats per limb, src/environments/ModularEnv.py:107-128; distribution
spec in SURVEY.md Appendix E).  Generated on the CPU with a seeded ``torch.Generator`` so
the same batch can be fed to the CUDA path, the oracle and the reference.

Per-limb layout: [0:3] pos-torso pos, [3:6] gravity (0,0,-9.81), [6:9] unit target
direction, [9:12] lin-vel (clip +-10), [12:15] ang-vel, [15:24] world axes of the limb's
x/y/z hinges (zero for the torso), [24:27] joint angles, [27:36] 3x(norm angle, lo, hi),
[36:40] limb-type one-hot, [40] limb height."""
from __future__ import annotations

import math
from typing import Dict

import torch

LIMB_OBS = 41
LIMB_ACT = 3


def make_obs(batch: int, n_limbs: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(int(seed))
    B, N = batch, n_limbs

    def rn(*s):
        return torch.randn(*s, generator=g)

    def ru(*s):
        return torch.rand(*s, generator=g)

    o = torch.zeros(B, N, LIMB_OBS)
    o[:, :, 0:3] = 0.3 * rn(B, N, 3)
    o[:, 0, 0:3] = 0.0
    o[:, :, 5] = -9.81
    th = (ru(B, 1) * 2 - 1) * math.pi
    o[:, :, 6] = torch.cos(th)
    o[:, :, 7] = torch.sin(th)
    o[:, :, 9:12] = (2.0 * rn(B, N, 3)).clamp(-10, 10)
    o[:, :, 12:15] = 3.0 * rn(B, N, 3)
    ax = rn(B, N, 3, 3)
    ax = ax / ax.norm(dim=-1, keepdim=True)
    o[:, :, 15:24] = ax.reshape(B, N, 9)
    o[:, 0, 15:24] = 0.0
    o[:, :, 24:27] = 0.5 * rn(B, N, 3)
    o[:, 0, 24:27] = 0.0
    o[:, :, 27:36] = ru(B, N, 9)
    o[:, 0, 27:36] = 0.5
    for n in range(N):
        o[:, n, 36 + (0 if n == 0 else 1 + (n - 1) % 3)] = 1.0
    o[:, :, 40] = 1.5 * ru(B, N)
    return o.reshape(B, N * LIMB_OBS).contiguous()


def make_batch(batch: int, n_limbs: int, seed: int = 1) -> Dict[str, torch.Tensor]:
    """TD3 minibatch dict with the keys ``Agent.update`` reads (src/agent.py:118-122)."""
    g = torch.Generator().manual_seed(int(seed) + 7919)
    B, N = batch, n_limbs
    return {
        "obs": make_obs(B, N, seed),
        "next_obs": make_obs(B, N, seed + 1),
        "action": torch.rand(B, N * LIMB_ACT, generator=g) * 2 - 1,
        "reward": torch.randn(B, 1, generator=g),
        "done": (torch.rand(B, 1, generator=g) < 0.01).float(),
    }


def rotate_about_gravity(obs: torch.Tensor, n_limbs: int, angle: float) -> torch.Tensor:
    """Rotate the eight 3-vectors of every limb by R_z(angle); scalars untouched.
    SET outputs must not change (subequivariance)."""
    B = obs.shape[0]
    o = obs.reshape(B, n_limbs, LIMB_OBS).clone()
    c, s = math.cos(angle), math.sin(angle)
    R = torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], dtype=obs.dtype, device=obs.device)
    v = o[:, :, :24].reshape(B, n_limbs, 8, 3)
    o[:, :, :24] = (v @ R.T).reshape(B, n_limbs, 24)
    return o.reshape(B, n_limbs * LIMB_OBS)
