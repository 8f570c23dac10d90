#!/usr/bin/env python
"""Times every projection shape of one SET TD3 update (forward / dgrad / wgrad forms) through
the C ABI on the current GPU: tcgen05 kernel with in-kernel split of both operands, with pre-split weights, and the
fp32 SIMT kernel; reports us per launch, TFLOP/s and the relative error against fp64.
usage: python tools/gemm_bench.py [T] [nrep]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sgrl_b200._lib import lib, ptr, stream, check


def timeit(fn, nrep):
    """us per call of fn, replayed from a CUDA graph of nrep back-to-back launches (no host launch cost in the number)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(nrep):
                fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / nrep


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 2304
    nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    # (name, M, N, K) forward Y = X W^T ; dgrad dX = dY W (M, K<-N) ; wgrad dW = dY^T X
    lin = [("G1", T, 256, 544), ("G2", T, 128, 256), ("QKV", T, 768, 256), ("VG", 3 * T, 252, 128), ("NGO", T, 128, 256),
           ("GO", 3 * T, 128, 256), ("L3L1", T, 512, 256), ("L4", T, 1024, 256), ("L2", T, 128, 256), ("H1G", T, 128, 544),
           ("H2G", T, 128, 128)]
    print(f"# T={T} tokens, {nrep} launches each; us = device time per launch, CUDA graph of back-to-back launches, warm L2")
    print(f"# {'shape':34s} {'tc us':>8s} {'pre us':>8s} {'simt us':>8s} {'tc TF':>7s} {'pre TF':>7s} {'simt TF':>7s}  err tc/pre/simt")
    tot = [0.0, 0.0, 0.0]
    for name, M, N, K in lin:
        X = torch.randn(M, K, device=dev, generator=g)
        W = torch.randn(N, K, device=dev, generator=g) / K ** 0.5
        dY = torch.randn(M, N, device=dev, generator=g)
        hi, lo = torch.empty_like(W), torch.empty_like(W)
        if W.numel() % 4 == 0:
            check(lib.sgrl_split_tf32(ptr(W), ptr(hi), ptr(lo), W.numel(), stream()))
        forms = []
        Y = torch.empty(M, N, device=dev)
        forms.append((f"fwd  {name} {M}x{N}x{K}", (X, K, 0, W, K, 0, Y, N, M, N, K), (hi, lo), 1, (X.double() @ W.double().T)))
        dX = torch.empty(M, K, device=dev)
        forms.append((f"dgrd {name} {M}x{K}x{N}", (dY, N, 0, W, K, 1, dX, K, M, K, N), (hi, lo), 1, (dY.double() @ W.double())))
        dW = torch.zeros(N, K, device=dev)
        tiles = ((N + 63) // 64) * ((K + 63) // 64)
        sk = max(1, min(64, (296 + tiles - 1) // tiles, (M // 16) // 4))
        forms.append((f"wgrd {name} {N}x{K}x{M} sk{sk}", (dY, N, 1, X, K, 1, dW, K, N, K, M), None, sk, (dY.double().T @ X.double())))
        for label, a, split, sk, ref in forms:
            A, lda, ta, B, ldb, tb, Cm, ldc, m, n, k = a
            flops = 2.0 * m * n * k
            res = []
            for mode in ("tc", "pre", "simt"):
                if mode == "pre" and split is None:
                    res.append((float("nan"), float("nan"))); continue
                acc = 1 if sk > 1 else 0

                def run():
                    if sk > 1:
                        Cm.zero_()
                    if mode == "pre":
                        check(lib.sgrl_gemm_presplit(ptr(A), lda, ta, ptr(split[0]), ptr(split[1]), ldb, tb, ptr(Cm), ldc, m, n, k, 1.0, None, None, 0, acc, sk, stream()))
                    else:
                        check(lib.sgrl_gemm(ptr(A), lda, ta, ptr(B), ldb, tb, ptr(Cm), ldc, m, n, k, 1.0, None, None, 0, acc, sk, 1 if mode == "tc" else 0, stream()))
                try:
                    us = timeit(run, nrep)
                    if sk > 1:
                        us -= timeit(lambda: Cm.zero_(), nrep)
                    run(); torch.cuda.synchronize()
                    err = ((Cm.double() - ref).norm() / ref.norm()).item()
                except Exception as ex:  # not eligible
                    us, err = float("nan"), float("nan")
                res.append((us, err))
            for i in range(3):
                if res[i][0] == res[i][0]:
                    tot[i] += res[i][0]
            tf = [flops / (r[0] * 1e-6) / 1e12 if r[0] == r[0] else float("nan") for r in res]
            print(f"{label:36s} {res[0][0]:8.1f} {res[1][0]:8.1f} {res[2][0]:8.1f} {tf[0]:7.1f} {tf[1]:7.1f} {tf[2]:7.1f}  {res[0][1]:.1e} {res[1][1]:.1e} {res[2][1]:.1e}")
    print(f"# sum us: tc {tot[0]:.0f}  pre(+wgrad tc) {tot[1]:.0f}  simt {tot[2]:.0f}")


if __name__ == "__main__":
    main()
