"""Host side of the full rebuild of Q(k) as a matrix product (faunus_b200/csrc/device/fb_fullq.cuh): the tile
layout `fb_debug_fullq_layout` derives from the k-vectors of PolicyIonIon::updateBox / PolicyIonIonIPBC::updateBox
(src/energy.cpp:133-186, 356-412). No device needed: the layout is pure host code; the product itself is checked against numpy
on the GPU (tests/test_gpu_parity.py::test_full_q_matrix_product)."""
import ctypes as C

import numpy as np
import pytest

import faunus_b200.native as native


def layout(n_cutoff, policy=0, spherical=True, box=(50.0, 50.0, 50.0)):
    lib = native.load()
    cfg = native.FbEwaldConfig(alpha=0.3, n_cutoff=n_cutoff, kappa=0.0, surface_dielectric_constant=0.0,
                               bjerrum_length=7.0, spherical_sum=int(spherical), policy=policy)
    ncc = int(np.ceil(n_cutoff))
    max_k = (ncc + 1) * (2 * ncc + 1) ** 2
    max_tiles = 4 * ((ncc + 4) // 4) * ((2 * ncc + 8) // 8) * ((2 * ncc + 64) // 64)
    nxyz = np.zeros(3 * max_k, dtype=np.int32)
    index = np.zeros(max_k, dtype=np.int32)
    tiles = np.zeros(4 * max_tiles, dtype=np.int32)
    order = np.zeros(max_tiles, dtype=np.int32)
    cols = np.zeros(max_tiles + 1, dtype=np.int32)
    n_k, n_tiles, n_cols = C.c_int(), C.c_int(), C.c_int()
    p = lambda a: a.ctypes.data_as(native.c_int_p)
    rc = lib.fb_debug_fullq_layout(C.byref(cfg), (C.c_double * 3)(*box), max_k, max_tiles, p(nxyz), p(index), p(tiles), p(order),
                                   p(cols), C.byref(n_k), C.byref(n_tiles), C.byref(n_cols))
    assert rc == 0
    K, T = n_k.value, n_tiles.value
    return (ncc, nxyz[:3 * K].reshape(K, 3), index[:K], tiles[:4 * T].reshape(T, 4), order[:T], cols[:n_cols.value + 1])


@pytest.mark.parametrize("n_cutoff,policy,spherical,expect_k", [(11.0, 0, True, 2975), (30.0, 0, True, 57950), (6.0, 1, True, None),
                                                               (7.0, 0, False, 8 * 15 * 15 - 1), (40.0, 0, True, None),
                                                               (13.5, 0, True, None), (9.0, 2, True, None)])
def test_every_kvector_has_its_own_slot(n_cutoff, policy, spherical, expect_k):
    """index[k] decodes to the k-vector's own integer triplet, no two k-vectors share a slot, every column lies inside the
    groups its tile computes, windows are at most 64 wide, and the tiles of a column follow each other"""
    ncc, n, index, tiles, order, cols = layout(n_cutoff, policy, spherical)
    K, T = len(n), len(tiles)
    if expect_k is not None:
        assert K == expect_k          # (2975: the reference's doctest count for n_cutoff 11, src/energy.cpp:672)
    assert len(np.unique(index)) == K
    t, row, col = index // 2048, (index % 2048) // 64, index % 64
    assert t.min() == 0 and t.max() == T - 1
    assert np.array_equal(tiles[t, 0] + row // 8, n[:, 0])
    assert np.array_equal(tiles[t, 1] + row % 8 - ncc, n[:, 1])
    assert np.array_equal(tiles[t, 2] + col - ncc, n[:, 2])
    assert np.all(col < 8 * tiles[t, 3]) and np.all(tiles[:, 3] >= 1) and np.all(tiles[:, 3] <= 8)
    assert np.all(tiles[:, 0] % 4 == 0) and np.all(tiles[:, 1] % 8 == 0)
    # storage order: the tile numbers never decrease by more than the windows of one column, columns are contiguous
    assert cols[0] == 0 and cols[-1] == T and np.all(np.diff(cols) >= 1)
    column_of_tile = np.searchsorted(cols, np.arange(T), side="right") - 1
    assert np.all(np.diff(column_of_tile[t]) >= 0)    # a range of columns is a range of stored k-vectors (the slabs)
    for c0, c1 in zip(cols[:-1], cols[1:]):
        assert np.all(tiles[c0:c1, 0] == tiles[c0, 0]) and np.all(tiles[c0:c1, 1] == tiles[c0, 1])
        assert np.all(np.diff(tiles[c0:c1, 2]) == 64)  # windows of one column, 64 columns apart
        assert np.all(tiles[c0:c1 - 1, 3] == 8)
    # heaviest tiles first
    assert sorted(order.tolist()) == list(range(T))
    assert np.all(np.diff(tiles[order, 3]) <= 0)


def test_slot_efficiency_of_the_headline_configuration():
    """S1 (n_cutoff 30, spherical): 55 tiles, 71 % of the computed slots hold a k-vector"""
    _, n, _, tiles, _, _ = layout(30.0)
    slots = int((32 * 8 * tiles[:, 3]).sum())
    assert len(tiles) == 55
    assert 0.70 < len(n) / slots < 0.73


def test_two_windows_beyond_64_columns():
    """n_cutoff 40: 81 nz values per column in the middle of the sphere — two windows; the IPBC octant (nz ≥ 0) needs one"""
    _, _, _, tiles, _, cols = layout(40.0)
    assert np.diff(cols).max() == 2
    _, n, _, tiles, _, cols = layout(40.0, policy=2)
    assert n.min() == 0 and np.diff(cols).max() == 1
