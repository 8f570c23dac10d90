"""TD3 glue kernels (csrc/td3.cuh).  CPU part: the numpy restatement of the in-kernel noise stream (oracle/philox.py)
against the Random123 known-answer vectors of Philox4x32-10.  GPU part: sgrl_td3_smooth_action_rng against that
restatement, and the reward statistics of sgrl_td3_critic_loss against torch."""
import numpy as np
import pytest
import torch

from oracle import philox


def test_philox_restatement_matches_random123_known_answers():
    kat = [  # Random123 kat_vectors, philox4x32 10 rounds: counter, key -> output
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array(key, dtype=np.uint32))[0]
        assert [int(x) for x in got] == want


def test_restated_normals_are_standard_normal():
    z = philox.normals(400_000, seed=1234, draw=1)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
    assert abs((z ** 3).mean()) < 2e-2 and abs((z ** 4).mean() - 3) < 5e-2
    assert not np.allclose(z[:1000], philox.normals(1000, seed=1234, draw=2))


@pytest.mark.gpu
@pytest.mark.parametrize("n", [27, 6912, 100_003])
def test_smooth_action_rng_matches_restatement(n):
    from sgrl_b200._lib import check, lib, ptr, stream
    g = torch.Generator(device="cuda").manual_seed(n)
    a = torch.rand(n, device="cuda", generator=g) * 2 - 1
    out, nz = torch.empty_like(a), torch.empty_like(a)
    draw = torch.zeros(1, dtype=torch.int32, device="cuda")
    seed, sigma, clip, amax = 0x1234_5678_9ABC_DEF0, 0.2, 0.5, 1.0
    for d in (1, 2):
        check(lib.sgrl_bump_step(ptr(draw), stream()))
        check(lib.sgrl_td3_smooth_action_rng(ptr(a), ptr(out), ptr(nz), sigma, clip, amax, n, seed, ptr(draw), stream()))
        want = philox.normals(n, seed, d) * sigma
        assert np.abs(nz.cpu().numpy() - want).max() < 2e-5            # fp32 log / sincospi vs fp64
        ref = np.clip(a.cpu().numpy().astype(np.float64) + np.clip(want, -clip, clip), -amax, amax)
        assert np.abs(out.cpu().numpy() - ref).max() < 3e-5


@pytest.mark.gpu
def test_critic_loss_kernel_reports_reward_sums():
    from sgrl_b200._lib import check, lib, ptr, stream
    G, N = 100, 9
    T = G * N
    gen = torch.Generator(device="cuda").manual_seed(3)
    r = lambda *s: torch.randn(*s, device="cuda", generator=gen)
    q1, q2, t1, t2, rew = r(T), r(T), r(T), r(T), r(G) * 3 + 1
    done = (torch.rand(G, device="cuda", generator=gen) < 0.1).float()
    tokg = torch.arange(G, dtype=torch.int32, device="cuda").repeat_interleave(N).contiguous()
    target, d1, d2, loss = torch.empty(T, device="cuda"), torch.empty(T, device="cuda"), torch.empty(T, device="cuda"), torch.zeros(1, device="cuda")
    st = torch.zeros(2, dtype=torch.float64, device="cuda")
    scale, gamma = 0.5, 0.99
    check(lib.sgrl_td3_critic_loss(ptr(q1), ptr(q2), ptr(t1), ptr(t2), ptr(rew), ptr(done), ptr(tokg), None, ptr(target), ptr(d1), ptr(d2),
                                   ptr(loss), gamma, scale, T, ptr(st), G, stream()))
    x = (rew * scale).double()
    assert abs(st[0].item() - x.sum().item()) < 1e-9 * G and abs(st[1].item() - (x * x).sum().item()) < 1e-9 * G
    mean, var = st[0].item() / G, (st[1].item() - st[0].item() ** 2 / G) / (G - 1)
    assert abs(mean - x.mean().item()) < 1e-12 and abs(var - x.var().item()) < 1e-10
    y = rew.repeat_interleave(N) * scale + (1 - done.repeat_interleave(N)) * gamma * torch.minimum(t1, t2)
    assert torch.allclose(target, y, rtol=1e-6, atol=1e-6)
    want = ((q1 - y) ** 2).mean() + ((q2 - y) ** 2).mean()
    assert abs(loss.item() - want.item()) < 1e-5 * want.item()
    assert torch.allclose(d1, 2 * (q1 - y) / T, rtol=1e-5, atol=1e-8)
