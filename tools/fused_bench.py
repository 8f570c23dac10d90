#!/usr/bin/env python
"""Warm device time (CUDA graph of back-to-back launches) of the fused projections of the tcgen05 forward schedule next to
the launches they replace.  usage: python tools/fused_bench.py [T] [nrep]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sgrl_b200._lib import lib, ptr, stream, check
from tools.gemm_bench import timeit


def split(w):
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    check(lib.sgrl_split_tf32(ptr(w), ptr(hi), ptr(lo), w.numel(), stream()), "split")
    return hi, lo


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 2304
    nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    print(f"# T={T} tokens; us per launch, CUDA graph of {nrep} back-to-back launches, warm L2")
    # ---- K1 + linear_g1 vs [gd projection] + [gram GEMM]
    X, gd, P = r(3 * T, 128), r(T, 3, 2), r(64, 128) / 11.3
    Phi, Plo = split(P)
    Z, Z2, G, F = r(T, 3, 32), r(T, 3, 32), torch.zeros(T, 544, device=dev), torch.zeros(T, device=dev)
    Wf = r(256, 544) / 23.0
    Wh, Wl = split(Wf)
    b256, A1 = r(256), torch.zeros(T, 256, device=dev)
    t_k1 = timeit(lambda: check(lib.sgrl_inv_feature_fwd(ptr(X), None, ptr(gd), ptr(P), None, ptr(Z), None, ptr(G), ptr(F), T, stream())), nrep)
    t_k1b = timeit(lambda: check(lib.sgrl_inv_feature_fwd(ptr(X), None, ptr(gd), ptr(P), ptr(P[32:]), ptr(Z), ptr(Z2), ptr(G), ptr(F), T, stream())), nrep)
    t_g1 = timeit(lambda: check(lib.sgrl_gemm_presplit(ptr(G), 544, 0, ptr(Wh), ptr(Wl), 544, 0, ptr(A1), 256, T, 256, 544, 1.0, ptr(b256), None, 1, 0, 1, stream())), nrep)
    t_gd = timeit(lambda: check(lib.sgrl_gemm_gd(ptr(X), 128, ptr(Phi), ptr(Plo), 128, ptr(gd), ptr(Z), 3 * T, 128, stream())), nrep)
    t_gr0 = timeit(lambda: check(lib.sgrl_gemm_gram(ptr(Z), ptr(Wh), ptr(Wl), ptr(b256), ptr(A1), 256, ptr(F), None, T, 256, 1, stream())), nrep)
    t_gr1 = timeit(lambda: check(lib.sgrl_gemm_gram(ptr(Z), ptr(Wh), ptr(Wl), ptr(b256), ptr(A1), 256, ptr(F), ptr(G), T, 256, 1, stream())), nrep)
    print(f"K1 (1 proj) {t_k1:6.1f} | K1 (2 proj) {t_k1b:6.1f} | linear_g1 from G {t_g1:6.1f}   ->   gd projection {t_gd:6.1f} | gram GEMM {t_gr0:6.1f} (keep G: {t_gr1:6.1f})")
    # ---- ng_out + LayerNorm
    O, Wn, bn = r(T, 256), r(128, 256) / 16.0, r(128)
    Wnh, Wnl = split(Wn)
    res, gam, bet = r(T, 256), r(128), r(128)
    y, x, st = torch.zeros(T, 256, device=dev), torch.zeros(T, 128, device=dev), torch.zeros(T, 2, device=dev)
    t_ngo = timeit(lambda: check(lib.sgrl_gemm_presplit(ptr(O), 256, 0, ptr(Wnh), ptr(Wnl), 256, 0, ptr(x), 128, T, 128, 256, 1.0, ptr(bn), None, 0, 0, 1, stream())), nrep)
    t_ln = timeit(lambda: check(lib.sgrl_gemm_ln(ptr(O), 256, ptr(Wnh), ptr(Wnl), ptr(bn), None, ptr(res[:, 128:]), 256, ptr(gam), ptr(bet), None, None,
                                                 ptr(y[:, 128:]), 256, ptr(x), None, ptr(st), None, 0, None, T, 256, stream())), nrep)
    print(f"ng_out GEMM {t_ngo:6.1f} (+ LayerNorm kernel ~3-4)   ->   GEMM with LayerNorm epilogue {t_ln:6.1f}")
    # ---- grouped pairs
    def pair(s0, s1):
        ops = []
        for M, N, K in (s0, s1):
            A, W, b = r(M, K), r(N, K) / K ** 0.5, r(N)
            h, l = split(W)
            ops.append((A, h, l, b, torch.zeros(M, N, device=dev), M, N, K))
        a, c = ops
        one = lambda o: check(lib.sgrl_gemm_presplit(ptr(o[0]), o[7], 0, ptr(o[1]), ptr(o[2]), o[7], 0, ptr(o[4]), o[6], o[5], o[6], o[7], 1.0, ptr(o[3]), None, 0, 0, 1, stream()))
        ta, tc = timeit(lambda: one(a), nrep), timeit(lambda: one(c), nrep)
        tp = timeit(lambda: check(lib.sgrl_gemm_pair(ptr(a[0]), a[7], ptr(a[1]), ptr(a[2]), ptr(a[3]), ptr(a[4]), a[6], a[5], a[6], a[7],
                                                      ptr(c[0]), c[7], ptr(c[1]), ptr(c[2]), ptr(c[3]), ptr(c[4]), c[6], c[5], c[6], c[7], 0, stream())), nrep)
        print(f"{s0} {ta:6.1f} + {s1} {tc:6.1f}   ->   grouped {tp:6.1f}")
    pair((T, 128, 256), (3 * T, 252, 128))
    pair((T, 128, 256), (3 * T, 128, 256))
    pair((T, 128, 256), (T, 1024, 256))
    pair((3 * T, 32, 128), (3 * T, 32, 128))


if __name__ == "__main__":
    main()
