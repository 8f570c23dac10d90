"""The fused projections of the tcgen05 forward schedule through the C ABI, each against fp64 torch:
Gram-generating GEMM (K1 -> linear_g1 fusion), [X P^T | gd] projection, residual + LayerNorm epilogue, grouped launch."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 2e-5


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()


def split(w):
    from sgrl_b200._lib import lib, ptr, stream, check
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    check(lib.sgrl_split_tf32(ptr(w), ptr(hi), ptr(lo), w.numel(), stream()), "split")
    return hi, lo


def fold(W):
    """(N,1024) -> triangle-folded (N,544), csrc/layout.h"""
    from sgrl_b200.packing import pack_indices
    slots, ri, ci = pack_indices()
    W3 = W.reshape(W.shape[0], 32, 32)
    Wf = torch.zeros(W.shape[0], 544, dtype=W.dtype, device=W.device)
    ri_t, ci_t = torch.tensor(ri, device=W.device), torch.tensor(ci, device=W.device)
    Wf[:, slots] = torch.where(ri_t == ci_t, W3[:, ri, ci], W3[:, ri, ci] + W3[:, ci, ri])
    return Wf


@pytest.mark.parametrize("T,N,keep", [(2304, 256, 1), (300, 128, 0), (1000, 256, 1), (129, 256, 0)])
def test_gram_gemm(T, N, keep):
    from sgrl_b200._lib import lib, ptr, stream, check
    g = torch.Generator(device="cuda").manual_seed(T + N)
    Z = torch.randn(T, 3, 32, device="cuda", generator=g)
    Z[:, :, 30] = torch.tensor([0.0, 0.0, -9.81], device="cuda")
    W = torch.randn(N, 1024, device="cuda", generator=g) / 32.0
    b = torch.randn(N, device="cuda", generator=g)
    Wf = fold(W).contiguous()
    hi, lo = split(Wf)
    Cm = torch.full((T, N + 4), 7.0, device="cuda")
    F = torch.zeros(T, device="cuda")
    Gp = torch.full((T, 544), 3.0, device="cuda") if keep else None
    check(lib.sgrl_gemm_gram(ptr(Z), ptr(hi), ptr(lo), ptr(b), ptr(Cm), N + 4, ptr(F), ptr(Gp), T, N, 1, stream()), "gram")
    torch.cuda.synchronize()
    Zd = Z.double()
    G = Zd.transpose(1, 2) @ Zd                                  # (T,32,32)
    ref = torch.relu(G.reshape(T, 1024) @ W.double().T + b.double())
    assert rel(Cm[:, :N], ref) < TOL
    assert torch.all(Cm[:, N:] == 7.0)
    assert rel(F, G.reshape(T, -1).norm(dim=1) + 1.0) < 1e-6
    if keep:
        from sgrl_b200.packing import pack_indices, tri_table
        slots, ri, ci = pack_indices()
        pads = [p for p, e in enumerate(tri_table()) if e is None]
        assert rel(Gp[:, slots], G[:, ri, ci]) < 1e-6 and torch.all(Gp[:, pads] == 0)


@pytest.mark.parametrize("T,N", [(40000, 256), (38001, 128), (60000, 256)])
def test_gram_gemm_persistent(monkeypatch, T, N):
    """Persistent GRAM projection of inference passes (csrc/gemm_tc_persist.cuh: all N <= 256 columns in one tile, every packed Gram
    row generated once) against fp64 and against the per-tile kernel; row tail, sub-matrix ldc, F = ||G|| + 1."""
    from sgrl_b200._lib import lib, ptr, stream, check
    g = torch.Generator(device="cuda").manual_seed(T + N)
    Z = torch.randn(T, 3, 32, device="cuda", generator=g)
    Z[:, :, 30] = torch.tensor([0.0, 0.0, -9.81], device="cuda")
    W = torch.randn(N, 1024, device="cuda", generator=g) / 32.0
    b = torch.randn(N, device="cuda", generator=g)
    hi, lo = split(fold(W).contiguous())
    out = []
    for mode in ("2", "0"):
        monkeypatch.setenv("SGRL_TC_PERSIST", mode)
        Cm = torch.full((T, N + 4), 7.0, device="cuda")
        F = torch.zeros(T, device="cuda")
        check(lib.sgrl_gemm_gram(ptr(Z), ptr(hi), ptr(lo), ptr(b), ptr(Cm), N + 4, ptr(F), None, T, N, 1, stream()), "gram")
        torch.cuda.synchronize()
        out.append((Cm, F))
    Zd = Z.double()
    G = Zd.transpose(1, 2) @ Zd
    ref = torch.relu(G.reshape(T, 1024) @ W.double().T + b.double())
    for Cm, F in out:
        assert rel(Cm[:, :N], ref) < TOL
        assert torch.all(Cm[:, N:] == 7.0)
        assert rel(F, G.reshape(T, -1).norm(dim=1) + 1.0) < 1e-6
    assert torch.equal(out[0][1], out[1][1])                      # F: same products and summation order in both kernels


@pytest.mark.parametrize("T3,K,ldw", [(6912, 128, 128), (900, 128, 136), (601, 128, 128)])
def test_gd_projection(T3, K, ldw):
    from sgrl_b200._lib import lib, ptr, stream, check
    g = torch.Generator(device="cuda").manual_seed(T3)
    T = T3 // 3
    T3 = T * 3
    X = torch.randn(T3, K, device="cuda", generator=g)
    Wfull = torch.randn(40, ldw, device="cuda", generator=g) / K ** 0.5      # rows 30.. stand for the next tensors of the arena
    gd = torch.randn(T, 3, 2, device="cuda", generator=g)
    hi, lo = split(Wfull)
    Z = torch.full((T3, 32), 5.0, device="cuda")
    check(lib.sgrl_gemm_gd(ptr(X), K, ptr(hi), ptr(lo), ldw, ptr(gd), ptr(Z), T3, K, stream()), "gd")
    torch.cuda.synchronize()
    ref = torch.cat([X.double() @ Wfull[:30, :K].double().T, gd.reshape(T3, 2).double()], dim=1)
    assert rel(Z, ref) < TOL
    assert torch.equal(Z[:, 30:], gd.reshape(T3, 2))


@pytest.mark.parametrize("T,K,div,second", [(2304, 256, 0, 0), (2304, 256, 1, 1), (300, 256, 1, 0), (131, 128, 0, 1)])
def test_layernorm_epilogue(T, K, div, second):
    from sgrl_b200._lib import lib, ptr, stream, check
    g = torch.Generator(device="cuda").manual_seed(T + K + div)
    A = torch.randn(T, K, device="cuda", generator=g)
    W = torch.randn(128, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(128, device="cuda", generator=g)
    F = (torch.rand(T, device="cuda", generator=g) * 3 + 0.5) if div else None
    res = torch.randn(T, 256, device="cuda", generator=g)
    gam, bet = torch.rand(128, device="cuda", generator=g) + 0.5, torch.randn(128, device="cuda", generator=g)
    gam2, bet2 = (torch.rand(128, device="cuda", generator=g) + 0.5, torch.randn(128, device="cuda", generator=g)) if second else (None, None)
    hi, lo = split(W)
    y = torch.full((T, 256), 9.0, device="cuda")
    x, x0, st = torch.zeros(T, 128, device="cuda"), torch.zeros(T, 128, device="cuda"), torch.zeros(T, 2, device="cuda")
    y2 = torch.full((T, 148), 4.0, device="cuda") if second else None
    st2 = torch.zeros(T, 2, device="cuda") if second else None
    check(lib.sgrl_gemm_ln(ptr(A), K, ptr(hi), ptr(lo), ptr(b), ptr(F), ptr(res[:, 128:]), 256, ptr(gam), ptr(bet), ptr(gam2), ptr(bet2),
                           ptr(y[:, 128:]), 256, ptr(x), ptr(x0), ptr(st), ptr(y2[:, 20:]) if second else None, 148, ptr(st2), T, K, stream()), "ln")
    torch.cuda.synchronize()
    r0 = A.double() @ W.double().T + b.double()
    if div:
        r0 = r0 / F.double()[:, None]
    rx = r0 + res[:, 128:].double()
    ln = lambda v, ga, be: torch.nn.functional.layer_norm(v, (128,), ga.double(), be.double(), 1e-5)
    ry = ln(rx, gam, bet)
    assert rel(x0, r0) < TOL and rel(x, rx) < TOL and rel(y[:, 128:], ry) < TOL
    assert torch.all(y[:, :128] == 9.0)
    assert rel(st[:, 0], rx.mean(1)) < 1e-4 and rel(st[:, 1], 1.0 / torch.sqrt(rx.var(1, unbiased=False) + 1e-5)) < TOL
    if second:
        assert rel(y2[:, 20:], ln(ry, gam2, bet2)) < TOL and torch.all(y2[:, :20] == 4.0)
        assert rel(st2[:, 1], 1.0 / torch.sqrt(ry.var(1, unbiased=False) + 1e-5)) < TOL


@pytest.mark.parametrize("s0,s1", [((2304, 128, 256), (6912, 252, 128)), ((2304, 128, 256), (2304, 1024, 256)), ((300, 768, 256), (900, 30, 128))])
def test_grouped_pair(s0, s1):
    from sgrl_b200._lib import lib, ptr, stream, check
    g = torch.Generator(device="cuda").manual_seed(s0[0] + s1[1])
    ops = []
    for M, N, K in (s0, s1):
        A = torch.randn(M, K, device="cuda", generator=g)
        W = torch.randn(N + 2, K, device="cuda", generator=g) / K ** 0.5
        b = torch.randn(N, device="cuda", generator=g)
        ldc = (N + 7) // 4 * 4
        ops.append((A, W, b, torch.full((M, ldc), 2.0, device="cuda"), ldc, M, N, K) + split(W))
    a, c = ops
    check(lib.sgrl_gemm_pair(ptr(a[0]), a[7], ptr(a[8]), ptr(a[9]), ptr(a[2]), ptr(a[3]), a[4], a[5], a[6], a[7],
                             ptr(c[0]), c[7], ptr(c[8]), ptr(c[9]), ptr(c[2]), ptr(c[3]), c[4], c[5], c[6], c[7], 1, stream()), "pair")
    torch.cuda.synchronize()
    for A, W, b, Cm, ldc, M, N, K, _, _ in ops:
        ref = torch.relu(A.double() @ W[:N].double().T + b.double())
        assert rel(Cm[:, :N], ref) < TOL
        assert torch.all(Cm[:, N:] == 2.0)
