"""TEST INFRASTRUCTURE (oracle): numpy restatement of the target-policy noise stream of sgrl_td3_smooth_action_rng
(csrc/td3.cuh): Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; the
Random123 reference implementation) + Box-Muller, four normals per call.  The reference draws this noise with
torch.randn_like(action) * policy_noise (src/agent.py:128); any N(0, sigma^2) stream is equivalent, so parity is defined
against this restatement, itself pinned to the Random123 known-answer vectors in tests/test_td3_glue.py."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr (n,4) uint32, key (2,) uint32 -> (n,4) uint32"""
    c = [ctr[:, i].astype(np.uint64) for i in range(4)]
    k0, k1 = int(key[0]), int(key[1])
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, 1).astype(np.uint32)


def normals(n: int, seed: int, draw: int) -> np.ndarray:
    """the first n standard normals of draw number `draw` under `seed`, in the kernel's element order"""
    nq = (n + 3) // 4
    q = np.arange(nq, dtype=np.uint64)
    ctr = np.stack([q & np.uint64(0xFFFFFFFF), q >> np.uint64(32), np.full(nq, draw, np.uint64), np.zeros(nq, np.uint64)], 1).astype(np.uint32)
    r = philox4x32_10(ctr, np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)).astype(np.float64)
    u = (r.astype(np.float32) + np.float32(0.5)).astype(np.float64) * 2.0 ** -32        # the kernel's fp32 (x + 0.5) rounding
    u = np.minimum(u, 0.99999994)
    r0, r1 = np.sqrt(-2 * np.log(u[:, 0])), np.sqrt(-2 * np.log(u[:, 2]))
    z = np.stack([r0 * np.cos(2 * np.pi * u[:, 1]), r0 * np.sin(2 * np.pi * u[:, 1]),
                  r1 * np.cos(2 * np.pi * u[:, 3]), r1 * np.sin(2 * np.pi * u[:, 3])], 1)
    return z.reshape(-1)[:n]
