"""CPU check of the algebra behind the packed Gram (csrc/layout.h, csrc/feature.cuh): for a symmetric G the contraction
vec(G) . W^T equals packed_triangle(G) . W'^T with W'[o][p(i,j)] = W[o][32i+j] + W[o][32j+i] (i<j), W[o][33i] (i=j); the
gradient w.r.t. W unfolds as dW[o][32a+b] = dW'[o][p(min,max)], and the backward's S = dG + dG^T reads 2*dG'[p] on the diagonal."""
import numpy as np
import torch


from sgrl_b200.packing import GP_K, pack_indices, tri_index, tri_table


def test_packing_is_a_bijection_with_the_documented_block_structure():
    tab = tri_table()
    assert len(tab) == GP_K and sum(e is not None for e in tab) == 528
    pads = [p for p, e in enumerate(tab) if e is None]
    assert pads == [478, 479, 510, 511] + list(range(532, 544))
    # k-blocks 0..13: two off-diagonal 4x4 blocks each, row-major inside a block
    for kb in range(14):
        for h in range(2):
            blk = tab[kb * 32 + h * 16: kb * 32 + h * 16 + 16]
            i0, j0 = blk[0]
            assert i0 % 4 == 0 and j0 % 4 == 0 and i0 < j0
            assert blk == [(i0 + a, j0 + b) for a in range(4) for b in range(4)]
    assert tri_index(0, 4) == 0 and tri_index(31, 31) == 448 + 64 + 19


def test_folded_contraction_and_gradients():
    g = torch.Generator().manual_seed(0)
    T, O_ = 7, 5
    Z = torch.randn(T, 3, 32, generator=g, dtype=torch.float64)
    W = torch.randn(O_, 1024, generator=g, dtype=torch.float64, requires_grad=True)
    G = torch.einsum("tri,trj->tij", Z, Z)
    y = G.reshape(T, 1024) @ W.T
    _, ri, ci = pack_indices()
    iu = (torch.tensor(ri), torch.tensor(ci))                           # the 528 real slots in packed order
    Gp = G[:, iu[0], iu[1]]                                             # (T, 528) packed triangle
    W3 = W.detach().reshape(O_, 32, 32)
    Wf = (W3 + W3.transpose(1, 2))[:, iu[0], iu[1]]
    Wf[:, iu[0] == iu[1]] = W3[:, iu[0][iu[0] == iu[1]], iu[1][iu[0] == iu[1]]]
    np.testing.assert_allclose((Gp @ Wf.T).numpy(), y.detach().numpy(), rtol=1e-12, atol=1e-12)
    # Frobenius norm from the triangle
    w = torch.where(iu[0] == iu[1], 1.0, 2.0).double()
    np.testing.assert_allclose(((Gp ** 2) * w).sum(1).sqrt().numpy(), G.reshape(T, -1).norm(dim=1).numpy(), rtol=1e-12)
    # gradient of the weight: unfold
    dy = torch.randn(T, O_, generator=g, dtype=torch.float64)
    (y * dy).sum().backward()
    dWf = dy.T @ Gp                                                     # (O, 528)
    dW = torch.zeros(O_, 32, 32, dtype=torch.float64)
    dW[:, iu[0], iu[1]] = dWf
    dW[:, iu[1], iu[0]] = dWf
    np.testing.assert_allclose(dW.reshape(O_, 1024).numpy(), W.grad.numpy(), rtol=1e-12, atol=1e-12)
    # gradient w.r.t. Z through the packed dG': S = dG + dG^T
    dGf = dy @ Wf                                                       # (T, 528) = dG[ij] + dG[ji] (i<j), dG[ii]
    S = torch.zeros(T, 32, 32, dtype=torch.float64)
    S[:, iu[0], iu[1]] = dGf
    S[:, iu[1], iu[0]] = dGf
    S[:, range(32), range(32)] *= 2.0
    dG = (dy @ W.detach()).reshape(T, 32, 32)
    np.testing.assert_allclose(S.numpy(), (dG + dG.transpose(1, 2)).numpy(), rtol=1e-12, atol=1e-12)
