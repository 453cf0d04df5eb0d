"""Host-side callers of the hot path (one rank): kernel tables, Hilbert keys, SFC order and octree view.

Thin wrappers over the C ABI (sphx_make_tables_host, sphx_hilbert_keys_host, sphx_host_tree_*); all work happens in
libsphx.so (csrc/host_domain.cpp).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _cabi


def make_box(lim, boundary) -> _cabi.SphxBox:
    b = _cabi.SphxBox()
    for i in range(6):
        b.lim[i] = float(lim[i])
    for i in range(3):
        b.boundary[i] = int(boundary[i])
    return b


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def make_tables(sinc_index: float = 6.0):
    """(wh, whd, K): ParticlesData::createTables of the reference (particles_data.hpp:380-387)."""
    L = _cabi.load()
    wh = np.zeros(20000, np.float32)
    whd = np.zeros(20000, np.float32)
    K = C.c_double(0)
    _cabi.check(L.sphx_make_tables_host(C.c_double(sinc_index), _p(wh), _p(whd), C.byref(K)))
    return wh, whd, K.value


def hilbert_keys(x, y, z, box_lim, boundary=(0, 0, 0)) -> np.ndarray:
    L = _cabi.load()
    x, y, z = (np.ascontiguousarray(a, np.float64) for a in (x, y, z))
    keys = np.zeros(x.size, np.uint64)
    b = make_box(box_lim, boundary)
    L.sphx_hilbert_keys_host(_p(x), _p(y), _p(z), x.size, C.byref(b), _p(keys))
    return keys


@dataclass
class HostTree:
    """Octree view arrays (cstone::OctreeNsView layout) + SFC permutation, all host numpy arrays."""
    order: np.ndarray
    keys: np.ndarray
    prefixes: np.ndarray
    childOffsets: np.ndarray
    internalToLeaf: np.ndarray
    levelRange: np.ndarray
    leaves: np.ndarray
    layout: np.ndarray
    centers: np.ndarray
    sizes: np.ndarray

    @property
    def num_nodes(self):
        return self.childOffsets.size

    @property
    def num_leaves(self):
        return self.layout.size - 1

    def as_dump_dict(self) -> dict:
        """same names as the reference-harness dumps (tests/refdata.py)"""
        return dict(numLeafNodes=np.array([self.num_leaves], np.int32), numNodes=np.array([self.num_nodes], np.int32),
                    tree_prefixes=self.prefixes, tree_childOffsets=self.childOffsets,
                    tree_internalToLeaf=self.internalToLeaf, tree_levelRange=self.levelRange, tree_leaves=self.leaves,
                    tree_layout=self.layout, tree_centers=self.centers, tree_sizes=self.sizes)


def build_tree(x, y, z, box_lim, boundary=(0, 0, 0), bucket_size: int = 64) -> HostTree:
    """Keys, SFC order and the converged bucket_size octree for particles given in arbitrary order.

    Tree arrays refer to the particles AFTER reordering with `order`."""
    L = _cabi.load()
    x, y, z = (np.ascontiguousarray(a, np.float64) for a in (x, y, z))
    n = x.size
    b = make_box(box_lim, boundary)
    h = L.sphx_host_tree_build(_p(x), _p(y), _p(z), n, C.byref(b), bucket_size)
    try:
        sz = np.zeros(2, np.int32)
        L.sphx_host_tree_sizes(h, _p(sz))
        nn, nl = int(sz[0]), int(sz[1])
        t = HostTree(order=np.zeros(n, np.uint32), keys=np.zeros(n, np.uint64), prefixes=np.zeros(nn, np.uint64),
                     childOffsets=np.zeros(nn, np.int32), internalToLeaf=np.zeros(nn, np.int32),
                     levelRange=np.zeros(23, np.int32), leaves=np.zeros(nl + 1, np.uint64),
                     layout=np.zeros(nl + 1, np.uint32), centers=np.zeros(nn * 3, np.float64),
                     sizes=np.zeros(nn * 3, np.float64))
        L.sphx_host_tree_get(h, _p(t.order), _p(t.keys), _p(t.prefixes), _p(t.childOffsets), _p(t.internalToLeaf),
                             _p(t.levelRange), _p(t.leaves), _p(t.layout), _p(t.centers), _p(t.sizes))
    finally:
        L.sphx_host_tree_free(h)
    return t
