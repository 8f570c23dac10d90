#!/usr/bin/env python
"""Stand-alone timing of the HBM-bound kernels at rollout size through the C ABI: K1 (sgrl_inv_feature_fwd) and
K2 (sgrl_attention_fwd).  GB/s = algorithmic bytes / device time.  python tools/k_bench.py [T] [n_limbs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sgrl_b200._lib import lib, ptr, stream, check

T = int(sys.argv[1]) if len(sys.argv) > 1 else 147456
n = int(sys.argv[2]) if len(sys.argv) > 2 else 9
T = T // n * n
G = T // n
dev = "cuda"
f = lambda *s: torch.randn(*s, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[reps // 2]


X, gd, P1, P2, V0, Ph = f(T, 3, 128), f(T, 3, 2), f(30, 128) * 0.1, f(30, 128) * 0.1, f(T, 3, 8), f(30, 136) * 0.1
Z, Z2, Gp, Fn = f(T, 3, 32), f(T, 3, 32), f(T, 544), f(T)
for name, fn, byts in (
    ("K1 layer (1 projection)", lambda: check(lib.sgrl_inv_feature_fwd(ptr(X), None, ptr(gd), ptr(P1), None, ptr(Z), None, ptr(Gp), ptr(Fn), T, stream())), 12 * 128 + 24 + 2176 + 4 + 384),
    ("K1 layer (2 projections)", lambda: check(lib.sgrl_inv_feature_fwd(ptr(X), None, ptr(gd), ptr(P1), ptr(P2), ptr(Z), ptr(Z2), ptr(Gp), ptr(Fn), T, stream())), 12 * 128 + 24 + 2176 + 4 + 768),
    ("K1 head (C=136)", lambda: check(lib.sgrl_inv_feature_fwd(ptr(X), ptr(V0), ptr(gd), ptr(Ph), None, ptr(Z), None, ptr(Gp), ptr(Fn), T, stream())), 12 * 136 + 24 + 2176 + 4 + 384),
):
    ms = timed(fn)
    print(f"{name:28s} T={T}: {ms * 1e3:8.1f} us  {byts * T / ms / 1e6:8.1f} GB/s")

qkv, vgp = f(T, 768) * 0.1, f(T, 3, 252)
o, og, p = f(T, 256), f(T, 3, 256), f(T, 2, 16)
cu = torch.arange(0, T + 1, n, dtype=torch.int32, device=dev)
rel, rw, rb = f(n, n, 3), f(2, 3), f(2)
for name, w, b in (("K2 fwd (layer 0: bias)", rw, rb), ("K2 fwd", None, None)):
    fn = lambda: check(lib.sgrl_attention_fwd(ptr(qkv), ptr(vgp), ptr(gd), ptr(w), ptr(b), ptr(cu), None, ptr(rel), G, n, ptr(o), ptr(og), ptr(p), stream()))
    ms = timed(fn)
    print(f"{name:28s} T={T} n={n}: {ms * 1e3:8.1f} us  {(3072 + 3024 + 24 + 1024 + 3072 + 128) * T / ms / 1e6:8.1f} GB/s")
