"""Shared parity helpers.  Tolerances (north_star): outputs and gradients within 1e-4
relative in fp32; rotation-about-gravity invariance within 1e-5.  Relative means
||delta||_2 / ||ref||_2 — never absolute: |action|~2e-2 and |Q|~5e-3 at init
(SURVEY.md Appendix H)."""
import os
import zlib

import numpy as np
import torch

RTOL = 1e-4
RTOL_ROT = 1e-5
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "set_golden.npz")
WEIGHT_SEED = 11
CASES = [
    ("3d_walker_2_right_leg_left_knee", 4),
    ("3d_hopper_3_shin", 4),
    ("3d_walker_7_full", 5),
    ("3d_humanoid_9_full", 8),
    ("3d_cheetah_14_full", 3),
]


def rel_err(a, b) -> float:
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def probe(name, shape):
    rng = np.random.Generator(np.random.PCG64([977, zlib.crc32(name.encode())]))
    return torch.tensor(rng.standard_normal(tuple(shape)), dtype=torch.float64)


def summarize(named) -> np.ndarray:
    rows = []
    for k, v in named.items():
        if v is None:
            rows.append((0.0, 0.0))
        else:
            v = v.detach().double().cpu()
            rows.append((v.norm().item(), (v * probe(k, v.shape)).sum().item()))
    return np.array(rows, dtype=np.float64)


def check_summary(got: np.ndarray, want: np.ndarray, rtol=RTOL, floor=1e-4, what=""):
    """Per-tensor check of [norm, probe-dot] summaries with the global floor of SURVEY.md
    Appendix H: |delta| <= rtol * max(norm_t, floor * norm_global).  The probe dot of a
    tensor with norm n has standard deviation n, so the same scale bounds it."""
    gn = float(np.sqrt((want[:, 0] ** 2).sum()))
    scale = np.maximum(want[:, 0], floor * gn)
    bad = []
    for i in range(want.shape[0]):
        dn = abs(got[i, 0] - want[i, 0])
        dp = abs(got[i, 1] - want[i, 1])
        if dn > rtol * scale[i] or dp > 8 * rtol * scale[i]:
            bad.append((i, got[i].tolist(), want[i].tolist()))
    assert not bad, f"{what}: {len(bad)} tensors out of tolerance, first: {bad[:4]}"
    g2 = float(np.sqrt((got[:, 0] ** 2).sum()))
    assert abs(g2 - gn) <= rtol * gn, f"{what}: global norm {g2} vs {gn}"


def load_golden():
    return np.load(GOLDEN, allow_pickle=False)


def golden_graph(gold, name, parents, device="cpu"):
    return {
        "parents": list(parents),
        "traversals": [torch.tensor(r, dtype=torch.long, device=device) for r in gold[name + "/traversals"]],
        "relation": torch.tensor(gold[name + "/relation"], device=device),
    }


def golden_batch(gold, name, device="cpu"):
    return {k: torch.tensor(gold[f"{name}/{k}"], device=device) for k in ("obs", "next_obs", "action", "reward", "done")}
