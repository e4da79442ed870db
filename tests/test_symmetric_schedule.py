"""CPU test of the work lists of the symmetric forward kernels, enumerated on the HOST with the kernels' own index
functions (sm3_debug_sym_enumerate): on one rank every unordered pair of 128-row tiles is computed exactly once, across
ranks (exchange mode 4) every ordered (row tile, column tile) pair of the global matrix is accounted for exactly once --
as the row sums of the rank that computed the tile or as the column sums it ships to the column owner."""
import ctypes as C

import numpy as np
import pytest

from skin_sm3_b200 import _lib


def enumerate_tiles(n_local, world, rank, tpc=0):
    lib = _lib.lib()
    cap = 1 << 20
    buf = (C.c_int * (5 * cap))()
    n = lib.sm3_debug_sym_enumerate(n_local, world, rank, tpc, buf, cap)
    assert 0 < n <= cap, (n, _lib.last_error())
    return np.frombuffer(buf, dtype=np.int32, count=5 * n).reshape(n, 5).copy()


def global_tile(owner, j_l, n_local, world):
    half = n_local // 128
    return owner * half + j_l if j_l < half else world * half + owner * half + (j_l - half)


@pytest.mark.parametrize("n_local,tpc", [(128, 0), (384, 1), (640, 3), (2304, 0), (4096, 0), (8192, 7)])
def test_single_rank_list_covers_the_upper_triangle_once(n_local, tpc):
    rec = enumerate_tiles(n_local, 1, 0, tpc)
    T = 2 * n_local // 128
    P = T // 2
    seen = {}
    for cta, R, J, docol, slab in rec:
        assert 2 * R <= J < T and slab == R and docol == int(J > 2 * R + 1)
        assert (R, J) not in seen
        seen[(R, J)] = cta
    assert len(seen) == P * (P + 1)
    cover = np.zeros((T, T), np.int32)          # (row tile, column tile) -> how often its sums reach the row tile's rows
    for (R, J) in seen:
        for rb in (0, 1):
            cover[2 * R + rb, J] += 1
            if J > 2 * R + 1:
                cover[J, 2 * R + rb] += 1
    assert (cover == 1).all()
    ctas = np.array(sorted(set(seen.values())))
    assert (ctas == np.arange(len(ctas))).all()                       # contiguous pieces, no empty CTA in between
    sizes = np.bincount(rec[:, 0])
    assert sizes.max() - sizes[:-1].min() <= 0 or sizes[:-1].min() == sizes.max()   # equal pieces, the last may be short


@pytest.mark.parametrize("world,n_local,tpc", [(2, 256, 0), (2, 384, 2), (3, 128, 0), (4, 384, 5), (5, 256, 0), (8, 256, 0),
                                               (8, 4096, 0), (16, 128, 1)])
def test_cross_rank_lists_cover_the_global_matrix_once(world, n_local, tpc):
    T_l = 2 * n_local // 128
    T_g = T_l * world
    cover = np.zeros((T_g, T_g), np.int32)
    work = []
    for rank in range(world):
        rec = enumerate_tiles(n_local, world, rank, tpc)
        work.append(len(rec))
        slabs = set()
        for cta, R, gJ, docol, slab in rec:
            for rb in (0, 1):
                gi = global_tile(rank, 2 * R + rb, n_local, world)
                cover[gi, gJ] += 1                                       # row sums stay with the rank that computed the tile
                if docol:
                    cover[gJ, gi] += 1                                   # column sums go to the column owner's rows
            if docol:
                slabs.add(slab)
        assert all(0 <= s < (1 + (world - 1) // 2 + (1 if world % 2 == 0 else 0)) * (T_l // 2) for s in slabs)
    assert (cover == 1).all(), np.argwhere(cover != 1)[:5]
    # the circulant assignment is balanced: every rank computes about half of its row block's tiles
    assert max(work) <= 1.15 * min(work) + T_l, work
    assert sum(work) <= 0.5 * (T_g // 2) * T_g + world * T_l


@pytest.mark.parametrize("world,n_local,d", [(2, 16384, 256), (2, 16384, 64), (4, 8192, 256), (8, 4096, 256), (8, 256, 64),
                                             (3, 384, 128), (16, 128, 64)])
def test_peer_step_scratch_covers_every_ranks_symmetric_workspace(world, n_local, d):
    """Whether exchange mode 4 applies must not depend on the rank (a rank that fell back to mode 2 alone would leave its
    partners waiting for column sums): the peer step's scratch is sized for the largest per-rank plan."""
    lib = _lib.lib()
    total = lib.sm3_infonce_step_peer_scratch_bytes(n_local, n_local * world, d)
    need = [lib.sm3_debug_mr_workspace(n_local, world, r) for r in range(world)]
    assert min(need) > 0 and total >= max(need), (total, need)
    assert max(need) - min(need) <= 0.2 * max(need)                 # the two antipodal classes differ by a few slabs at most
