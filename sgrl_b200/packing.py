"""Packed layout of the symmetric Gram G = Z^T Z (csrc/layout.h): the 528 entries (i <= j) of the 32x32 matrix in the
order the tcgen05 GEMM generates them (28 off-diagonal 4x4 blocks, then the 8 diagonal blocks), zero-padded to 544."""
GP_K = 544
GP_OFF = 448


def tri_index(i: int, j: int) -> int:
    """Slot of G[i][j], i <= j (csrc/layout.h tri_index)."""
    ib, jb, a, b = i >> 2, j >> 2, i & 3, j & 3
    if ib < jb:
        return (ib * 7 - (ib * (ib - 1)) // 2 + (jb - ib - 1)) * 16 + a * 4 + b
    return GP_OFF + (ib // 3) * 32 + (ib % 3) * 10 + (a * 4 - (a * (a - 1)) // 2 + (b - a))


def tri_table():
    """[(i, j) | None] * 544: which Gram entry each slot holds (None = zero padding)."""
    t = [None] * GP_K
    for i in range(32):
        for j in range(i, 32):
            p = tri_index(i, j)
            assert t[p] is None
            t[p] = (i, j)
    return t


def pack_indices():
    """(slots, rows, cols) index lists: packed[:, slots] = G[:, rows, cols]."""
    tab = tri_table()
    slots = [p for p, e in enumerate(tab) if e is not None]
    return slots, [tab[p][0] for p in slots], [tab[p][1] for p in slots]
