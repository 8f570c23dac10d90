"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference replay buffer
(src/common/buffer.py:35-145).  Only ``tests/`` and ``bench.py``'s CPU legs may import it;
nothing under ``sgrl_b200/`` does.

Pinned: tests/test_buffer.py checks it against the unmodified reference ``ReplayBuffer``
(imported through oracle/ref_loader.py with stub gym spaces) whenever /root/reference is on
the machine: same storage contents, same ``curr`` / ``max_sample_size`` bookkeeping and,
for equal ``random`` / ``numpy.random`` seeds, the same sampled indices and batches.
"""
from __future__ import annotations

import random
import warnings

import numpy as np


class ReplayOracle:
    """buffer.py:35-66: five host arrays, a write cursor and the number of valid rows."""

    def __init__(self, obs_dim: int, action_dim: int, max_buffer_size: int = 1000000):
        self.max_buffer_size = max_buffer_size
        self.curr = 0
        self.obs_dim, self.action_dim = obs_dim, action_dim
        self.obs_buffer = np.zeros((max_buffer_size, obs_dim), dtype=np.float32)
        self.action_buffer = np.zeros((max_buffer_size, action_dim), dtype=np.float32)
        self.next_obs_buffer = np.zeros((max_buffer_size, obs_dim), dtype=np.float32)
        self.reward_buffer = np.zeros((max_buffer_size,), dtype=np.float32)
        self.done_buffer = np.zeros((max_buffer_size,), dtype=np.float32)
        self.max_sample_size = 0

    def clear(self):                                            # buffer.py:67-69
        self.max_sample_size = 0
        self.curr = 0

    def add_transition(self, obs, action, next_obs, reward, done):   # buffer.py:75-84
        self.obs_buffer[self.curr] = obs
        self.action_buffer[self.curr] = action
        self.next_obs_buffer[self.curr] = next_obs
        self.reward_buffer[self.curr] = reward
        self.done_buffer[self.curr] = done
        self.curr = (self.curr + 1) % self.max_buffer_size
        self.max_sample_size = min(self.max_sample_size + 1, self.max_buffer_size)

    def add_traj(self, obs_list, action_list, next_obs_list, reward_list, done_list):   # buffer.py:71-73
        for t in zip(obs_list, action_list, next_obs_list, reward_list, done_list):
            self.add_transition(*t)

    def draw_indices(self, batch_size, sequential=False, allow_duplicate=False):        # buffer.py:87-101
        if not allow_duplicate:
            if batch_size > self.max_sample_size:
                warnings.warn("Sampling larger than buffer size")
            batch_size = min(self.max_sample_size, batch_size)
        if sequential:
            start = random.choice(range(self.max_sample_size))
            return [(start + i) % self.max_sample_size for i in range(batch_size)]
        if allow_duplicate:
            return np.random.choice(range(self.max_sample_size), batch_size)
        return random.sample(range(self.max_sample_size), batch_size)

    def get_batch(self, indices):                                                       # buffer.py:103-111,127-134
        return dict(obs=self.obs_buffer[indices], action=self.action_buffer[indices], next_obs=self.next_obs_buffer[indices],
                    reward=self.reward_buffer[indices].reshape(-1, 1), done=self.done_buffer[indices].reshape(-1, 1))

    def sample(self, batch_size, sequential=False, allow_duplicate=False):
        return self.get_batch(self.draw_indices(batch_size, sequential, allow_duplicate))
