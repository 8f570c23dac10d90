"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference SET modules on CPU.

Only ``tests/``, ``tests/golden/make_golden.py`` and ``bench.py``'s reference arm may use
this file.  Nothing under ``sgrl_b200/`` imports it.

The reference (alpc91/SGRL, ``src/``) cannot be imported directly in this image because
``SEActor -> ModularActor -> utils -> xmltodict, gym, wrappers`` (src/ModularActor.py:5,
src/utils.py:6-9) are not installed.  Those imports are host-only glue that the SET hot
path never touches, so we register empty stub modules for them
and import the reference files where they lie.  Nothing is copied out of the reference.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

_SEARCH = (
    os.environ.get("SGRL_REF", ""),
    "/root/reference/src",
    os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "src"),
)


def find_reference():
    """Path of the reference ``src`` directory, or None when it is not on this machine (SGRL_REF_DISABLE=1: pretend so)."""
    if os.environ.get("SGRL_REF_DISABLE"):
        return None
    for p in _SEARCH:
        if p and os.path.isfile(os.path.join(p, "SEActor.py")):
            return p
    return None


class AttrDict(dict):
    """Attribute access over a dict (the reference wraps its config the same way)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def default_args(**over):
    """Hyper-parameters the SET hot path reads (src/arguments.py, src/configs/default.py,
    src/main.py:54,104-125) with the values start.sh / start_humanoid.sh run with."""
    a = AttrDict(
        actor_type="set", critic_type="set",
        limb_obs_size=41, limb_action_size=3, msg_dim=32, batch_size=100,
        max_action=1.0, max_children=None, disable_fold=False, td=False, bu=False,
        attention_embedding_size=128, attention_heads=2, attention_hidden_size=256,
        attention_layers=3, dropout_rate=0.0, condition_decoder_on_features=1,
        transformer_norm=1, traversal_types=["pre", "inlcrs", "postlcrs"], rel_size=3,
        lr=1e-4, discount=0.99, policy_noise=0.2, noise_clip=0.5, policy_freq=2,
        grad_clipping_value=0.1, expl_noise=0.126,
        agent=AttrDict(target_smoothing_tau=0.005, reward_scale=1.0),
    )
    a.update(over)
    return a


_loaded = None


def load_reference():
    """Import the reference modules (CPU).  Returns a namespace with SEActor, SECritic,
    utils, agent, util.  Raises FileNotFoundError when the reference is absent."""
    global _loaded
    if _loaded is not None:
        return _loaded
    src = find_reference()
    if src is None:
        raise FileNotFoundError("reference src/ not found (looked in $SGRL_REF, /root/reference/src, baseline/_ref/src)")
    import numpy
    import torch

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    stub("xmltodict")
    spaces = stub("gym.spaces", Box=object, Discrete=object, MultiBinary=object, space=types.ModuleType("space"))
    stub("gym.spaces.discrete", Discrete=object)
    stub("gym.spaces.box", Box=object)
    reg = stub("gym.envs.registration", register=lambda **k: None)
    envs = stub("gym.envs", registration=reg)
    stub("gym", Wrapper=object, Space=object, spaces=spaces, envs=envs)
    stub("wrappers")
    stub("numpy.lib.arraysetops", isin=numpy.isin)

    if src not in sys.path:
        sys.path.insert(0, src)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from common import util  # type: ignore
        util.device = torch.device("cpu")
        import SEActor  # type: ignore
        import SECritic  # type: ignore
        import utils  # type: ignore
        import agent  # type: ignore
    ns = types.SimpleNamespace(SEActor=SEActor, SECritic=SECritic, utils=utils, agent=agent, util=util, src=src)
    _loaded = ns
    return ns
