"""sgrl_b200 — B200-native SET (subequivariant transformer) TD3 hot path of alpc91/SGRL.

Importing the compute modules loads libsgrl_b200.so; there is no CPU fallback.
Host-only helpers (graph, morphologies, synth, names) import without it.
"""
__all__ = ["graph", "morphologies", "synth", "names"]
__version__ = "0.1.0"
