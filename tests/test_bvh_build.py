"""BuildBVH's object order on the GPU (grt_bvh_order) against the CPU restatement of bvhHelper."""
import ctypes as C
import os

import numpy as np
import pytest

import go_raytracer_b200 as g
from go_raytracer_b200 import _native as N
import parity_util as PU


def _as_np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.frombuffer((C.c_uint8 * (n * dtype.itemsize)).from_address(ptr), dtype=dtype).copy()


def _boxes(n, seed, ties=False):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-50, 50, size=(n, 3))
    if ties:   # many equal minima and equal (min, max) pairs: tie order must be list order
        c = np.round(c / 10.0) * 10.0
    h = rng.uniform(0.0, 3.0, size=(n, 3))
    if ties:
        h = np.round(h)
    return np.concatenate([c - h, c + h], axis=1)


def test_ref_order_small_cases_by_hand():
    # three boxes spread along y: sorted by y-min; spans of one and two objects are not sorted (bvh.go:44-49)
    b = np.array([[0, 5, 0, 1, 6, 1], [0, 1, 0, 1, 2, 1], [0, 3, 0, 1, 4, 1]], dtype=np.float64)
    assert PU.ref_bvh_order(b).tolist() == [1, 2, 0]
    assert PU.ref_bvh_order(b[:2]).tolist() == [0, 1]
    # equal minima: the maximum decides (bvh.go:28-31)
    b = np.array([[0, 0, 0, 9, 1, 1], [0, 0, 0, 5, 1, 1], [0, 0, 0, 7, 1, 1]], dtype=np.float64)
    assert PU.ref_bvh_order(b).tolist() == [1, 2, 0]


def test_bvh_order_fails_loudly_without_a_device():
    if N.lib().grt_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(N.GrtError):
        g.bvh_order(_boxes(10, 1))


@pytest.mark.gpu
@pytest.mark.parametrize("n,ties", [(1, False), (2, False), (3, False), (4, True), (7, False), (33, True), (1000, False),
                                    (1000, True), (4097, True), (200_000, False), (131_072, True)])
def test_gpu_order_equals_reference_order(n, ties):
    b = _boxes(n, 100 + n, ties)
    got = g.bvh_order(b)
    assert np.array_equal(got, PU.ref_bvh_order(b))          # index work: bit-exact
    assert np.array_equal(np.sort(got), np.arange(n))         # a permutation


@pytest.mark.gpu
def test_gpu_order_degenerate_boxes():
    """Flat and coincident boxes (padded axes, -0.0 vs +0.0 minima) still give the reference order."""
    rng = np.random.default_rng(3)
    n = 5000
    c = rng.uniform(-1, 1, size=(n, 3))
    c[:, 1] = 0.0
    c[::2, 1] = -0.0
    b = np.concatenate([c, c], axis=1)           # zero-size boxes on a plane
    assert np.array_equal(g.bvh_order(b), PU.ref_bvh_order(b))


@pytest.mark.gpu
def test_flatten_with_gpu_build_equals_host_build():
    """The flattened scene (nodes, triangles, list entries) is identical whichever side did BuildBVH's sorts."""
    flats = {}
    for mode in ("cpu", "gpu"):
        os.environ["GRT_BVH_BUILD"] = mode
        try:
            s, _ = g.builtin_scene(8, width=32, spp=4, mesh_segments=120)
            f = s.flatten()
            flats[mode] = (f.n_nodes, f.n_tris, f.n_items,
                           {"nodes": _as_np(f.nodes, f.n_nodes, N.NODE_DTYPE), "tris": _as_np(f.tris, f.n_tris, N.TRI_DTYPE),
                            "items": _as_np(f.items, f.n_items, np.dtype("<u4"))})
        finally:
            os.environ.pop("GRT_BVH_BUILD", None)
    a, b = flats["cpu"], flats["gpu"]
    assert a[:3] == b[:3] and a[1] > 20000
    for k in a[3]:
        assert a[3][k].tobytes() == b[3][k].tobytes(), k
