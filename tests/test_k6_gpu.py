"""K6 unit tests: sgrl_sumsq / sgrl_adam_clip / sgrl_polyak / sgrl_split_tf32 called directly through the C ABI on random
flat arenas, against torch.nn.utils.clip_grad_norm_ + torch.optim.Adam (src/agent.py:150-156,170-178) and the three
tensor ops of the reference's soft update (src/common/functional.py:7-10)."""
import pytest
import torch

from sgrl_b200._lib import check, lib, ptr, stream

pytestmark = pytest.mark.gpu


def _rand(n, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(n, device="cuda", generator=g) * scale


def _rna_tf32(x):
    """round-to-nearest, ties away, to 10 mantissa bits (what cvt.rna.tf32.f32 does) with integer ops"""
    b = x.view(torch.int32)
    return ((b + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("n", [4, 1024, 1 << 20, (1 << 20) + 12])
def test_sumsq_accumulates(n):
    g = _rand(n, n)
    out = torch.full((1,), 3.0, device="cuda")
    check(lib.sgrl_sumsq(ptr(g), n, ptr(out), stream()))
    want = 3.0 + (g.double() ** 2).sum().item()
    assert abs(out.item() - want) <= 2e-6 * want


@pytest.mark.parametrize("max_norm,grad_scale,gmag", [(0.1, 1.0, 1.0), (0.1, 0.25, 1e-3), (0.0, 1.0, 1.0), (1e3, 0.5, 1.0)])
def test_adam_clip_matches_torch(max_norm, grad_scale, gmag):
    """Five steps on one flat arena: bias correction (step 1..5), the clip coefficient max_norm / (norm + 1e-6) in both
    regimes (clipping / not clipping), grad_scale (the 1/world of the data-parallel sum), and the tf32 split refresh."""
    n, lr, b1, b2, eps = 300_000, 1e-4, 0.9, 0.999, 1e-8
    p = _rand(n, 1)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=lr, betas=(b1, b2), eps=eps)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    hi, lo = torch.empty_like(p), torch.empty_like(p)
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    ss = torch.zeros(1, device="cuda")
    for it in range(5):
        g = _rand(n, 10 + it, gmag)
        ref.grad = g.clone() * grad_scale
        if max_norm > 0:
            torch.nn.utils.clip_grad_norm_([ref], max_norm)
        opt.step()
        ss.zero_()
        check(lib.sgrl_sumsq(ptr(g), n, ptr(ss), stream()))
        check(lib.sgrl_bump_step(ptr(step), stream()))
        check(lib.sgrl_adam_clip(ptr(p), ptr(g), ptr(m), ptr(v), n, ptr(ss), ptr(step), lr, b1, b2, eps, max_norm, grad_scale,
                                 ptr(hi), ptr(lo), stream()))
        st = opt.state[ref]
        # the update is lr-sized (1e-4) on O(1) weights: compare the STEP, not the parameter, and the moments
        assert (p - ref.data).abs().max().item() <= 2e-7 * lr / 1e-4 + 1.2e-7 * p.abs().max().item(), it
        assert ((m - st["exp_avg"]).norm() / st["exp_avg"].norm()).item() < 1e-6, it
        assert ((v - st["exp_avg_sq"]).norm() / st["exp_avg_sq"].norm()).item() < 1e-6, it
    assert int(step.item()) == 5
    delta = (p - _rand(n, 1)).double()                      # five Adam steps moved every weight by about 5 lr
    dref = (ref.data - _rand(n, 1)).double()
    assert ((delta - dref).norm() / dref.norm()).item() < 1e-3
    # the split the tcgen05 GEMMs stream: hi = rna_tf32(p), lo = rna_tf32(p - hi)
    assert torch.equal(hi, _rna_tf32(p))
    assert torch.equal(lo, _rna_tf32(p - hi))
    assert ((hi.double() + lo.double() - p.double()).abs().max() / p.abs().max()).item() < 2.0 ** -21


def test_adam_without_split_pointers_leaves_the_same_parameters():
    n = 4096
    p1, p2 = _rand(n, 2), _rand(n, 2)
    g = _rand(n, 3)
    for p, with_split in ((p1, True), (p2, False)):
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        hi, lo = torch.empty_like(p), torch.empty_like(p)
        step = torch.ones(1, dtype=torch.int32, device="cuda")
        ss = (g.double() ** 2).sum().float().reshape(1)
        check(lib.sgrl_adam_clip(ptr(p), ptr(g), ptr(m), ptr(v), n, ptr(ss), ptr(step), 1e-4, 0.9, 0.999, 1e-8, 0.1, 1.0,
                                 ptr(hi) if with_split else None, ptr(lo) if with_split else None, stream()))
    assert torch.equal(p1, p2)


@pytest.mark.parametrize("tau", [0.005, 1.0, 0.3])
def test_polyak_is_bit_exact_with_the_reference_ops(tau):
    n, n_live = 200_000, 120_000
    s, t = _rand(n, 4), _rand(n, 5)
    want = tau * s + (1 - tau) * t                               # functional.py:9: three fp32 tensor ops
    hi, lo = torch.zeros(n_live, device="cuda"), torch.zeros(n_live, device="cuda")
    check(lib.sgrl_polyak(ptr(t), ptr(s), n, tau, ptr(hi), ptr(lo), n_live, stream()))
    assert torch.equal(t, want)
    assert torch.equal(hi, _rna_tf32(t[:n_live]))
    assert torch.equal(lo, _rna_tf32(t[:n_live] - hi))


def test_split_tf32_kernel():
    w = torch.cat([_rand(100_000, 6), _rand(100_000, 7, 1e-6), _rand(100_000, 8, 1e6)])
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    check(lib.sgrl_split_tf32(ptr(w), ptr(hi), ptr(lo), w.numel(), stream()))
    assert torch.equal(hi, _rna_tf32(w)) and torch.equal(lo, _rna_tf32(w - hi))
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0 and int((lo.view(torch.int32) & 0x1FFF).abs().max()) == 0
