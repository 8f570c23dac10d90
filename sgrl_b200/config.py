"""Hyper-parameters the SET hot path reads from the reference's ``args`` object, with the values
``start.sh`` / ``start_humanoid.sh`` run with (src/arguments.py, src/configs/default.py:10,61,
src/configs/3d.py, src/main.py:54,104-125).  ``Agent(default_args())`` builds the reference setup."""
from __future__ import annotations


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def default_args(**over):
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
