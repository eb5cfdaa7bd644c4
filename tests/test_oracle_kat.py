"""Pins the oracle's base layer against the reference's OWN unit-test vectors
(internal/vec/vec_test.go:24-154, internal/interval/interval_test.go:9-72, internal/ray/ray_test.go:11-19)
and the RNG against the published Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
import ctypes as C
import math
import numpy as np
from oracle import oracle_py as O

L = O.lib()


def d3(v):
    return (C.c_double * 3)(*v)


def vec_op(op, a, b=(0, 0, 0), c=0.0):
    out = (C.c_double * 3)()
    L.orc_vec_op(op, d3(a), d3(b), float(c), out)
    return tuple(out)


def vec_scalar(op, a, b=(0, 0, 0)):
    return L.orc_vec_scalar(op, d3(a), d3(b))


def test_vec_algebra_reference_vectors():
    a, b = (1, 2, 3), (4, 5, 6)
    assert vec_op(0, a, b) == (5, 7, 9)                 # TestVecAdd
    assert vec_op(1, a, b) == (-3, -3, -3)              # TestVecSub
    assert vec_op(2, a, b) == (4, 10, 18)               # TestVecMul
    assert vec_op(3, a, b) == (0.25, 0.4, 0.5)          # TestVecDiv
    assert vec_op(4, a) == (-1, -2, -3)                 # TestVecNegate
    assert vec_scalar(4, a, (1, 2, 3)) == 1 and vec_scalar(4, a, (1, 2, 4)) == 0   # TestVecEquals
    assert vec_scalar(0, a, b) == 32.0                  # TestVecDot
    assert vec_op(5, a, b) == (-3, 6, -3)               # TestVecCross
    assert vec_op(6, a, (2, 0, 0)) == (2, 4, 6)         # TestVecScale / ScaleInplace
    assert vec_scalar(1, a) == 14.0                     # TestVecLengthSquared
    assert vec_scalar(2, a) == math.sqrt(14)            # TestVecLength
    # TestVecUnitVector: the reference expects exact equality with x/sqrt(14); UnitVector is Scale(1/Length)
    # (vec.go:125), so compare against that expression
    inv = 1 / math.sqrt(14)
    assert vec_op(7, a) == (1 * inv, 2 * inv, 3 * inv)
    assert vec_scalar(3, (1e-9, 1e-9, 1e-9)) == 1 and vec_scalar(3, (1e-7, 1e-7, 1e-7)) == 0   # TestVecNearZero


def test_write_color_reference_vector():
    out = (C.c_int * 3)()
    L.orc_color_bytes(d3((0, 128, 255)), out)           # TestWriteColor: "0 255 255"
    assert tuple(out) == (0, 255, 255)
    L.orc_color_bytes(d3((float("nan"), 0.25, -1.0)), out)   # color.go:28-36 NaN -> 0; sqrt(.25)*256 = 128
    assert tuple(out) == (0, 128, 0)


def test_interval_reference_vectors():
    iv = lambda op, x: L.orc_interval(op, 0.0, 1.0, x)
    assert iv(0, 0.5) == 1 and iv(0, 0) == 1 and iv(0, 1) == 1 and iv(0, -1) == 0 and iv(0, 2) == 0   # Contains (closed)
    assert iv(1, 0.5) == 1 and iv(1, 0.001) == 1 and iv(1, 0.99) == 1 and iv(1, -1) == 0 and iv(1, 1.1) == 0
    assert iv(1, 0.0) == 0 and iv(1, 1.0) == 0          # Surrounds is open (interval.go:33)
    assert iv(2, 0.5) == 0.5 and iv(2, -1) == 0 and iv(2, 2) == 1 and iv(2, 0) == 0 and iv(2, 1) == 1


def test_ray_at_reference_vector():
    out = (C.c_double * 3)()
    L.orc_ray_at(d3((0, 0, 0)), d3((0, 1, 0)), 1.0, out)
    assert tuple(out) == (0, 1, 0)


def test_philox_known_answers():
    def ph(ctr, key):
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        L.orc_philox(c, k, o)
        return tuple(o)
    assert ph([0] * 4, [0] * 2) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert ph([0xffffffff] * 4, [0xffffffff] * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_uniform_is_open_interval_and_fp32_exact():
    us = np.array([L.orc_uniform(0xC0FFEE, p, 3, 1, 0, i) for p in range(50) for i in range(9)])
    assert (us > 0).all() and (us < 1).all()
    assert (us.astype(np.float32).astype(np.float64) == us).all()      # odd multiples of 2^-24
    assert abs(us.mean() - 0.5) < 0.05


def test_aabb_slab_semantics():
    hit = lambda bmin, bmax, o, d, t0=0.001, t1=float("inf"): L.orc_aabb_hit(d3(bmin), d3(bmax), d3(o), d3(d), t0, t1)
    assert hit((0, 0, 0), (1, 1, 1), (-1, .5, .5), (1, 0, 0)) == 1
    assert hit((0, 0, 0), (1, 1, 1), (-1, 2, .5), (1, 0, 0)) == 0
    assert hit((0, 0, 0), (1, 1, 1), (2, .5, .5), (1, 0, 0)) == 0        # behind the ray
    # Go's builtin min/max propagate NaN (aabb.go:104-105): origin on a slab plane with a zero direction
    # component gives 0*Inf = NaN and the box is ACCEPTED even though the ray misses in y
    assert hit((0, 0, 0), (1, 1, 1), (0, 5, .5), (0, 0, 1), -10, 10) == 1


# imageLoader_test.go:64-90: the 5x5 PNG fixture's exact pixel data (row-major, idx = y*Width + x)
REF_IMG_DATA = [(209, 226, 249), (161, 176, 189), (126, 146, 139), (146, 162, 191), (214, 230, 254),
                (124, 132, 172), (53, 64, 122), (63, 80, 105), (26, 34, 112), (154, 165, 198),
                (83, 88, 143), (4, 13, 116), (19, 35, 120), (93, 119, 64), (94, 108, 131),
                (138, 146, 181), (0, 0, 113), (0, 12, 114), (110, 122, 112), (140, 149, 164),
                (220, 231, 249), (111, 122, 166), (109, 122, 166), (140, 152, 175), (215, 227, 246)]


def test_image_texture_lookup_on_the_reference_fixture():
    """imageTexture.Value (texture.go:70-86) over the pixel data the reference's own imageLoader test pins
    (imageLoader_test.go:64-90): u -> |fmod(u,1)|, v -> 1 - |fmod(v,1)|, i = int(u (W-1)), j = int(v (H-1)),
    colour = byte * (1/255).  (With this rule the last column is never sampled and the last row only at integer v.)"""
    import go_raytracer_b200 as g
    img = np.array(REF_IMG_DATA, dtype=np.uint8).reshape(5, 5, 3)
    sc = g.Scene()
    tex = sc.NewImageTextureFromArray(img)
    lam = sc.NewTexturedLambertian(tex)
    light = sc.NewQuad((-1, 5, -1), (2, 0, 0), (0, 0, 2), sc.NewDiffuseLight((4, 4, 4)))
    sc.set_world(sc.NewHittableList([sc.NewQuad((0, 0, 0), (1, 0, 0), (0, 1, 0), lam), light]))
    sc.set_lights(sc.NewHittableList([light]))
    ow = O.OracleWorld(sc)
    scale = 1.0 / 255.0
    for j in range(4):
        for i in range(4):
            u, v = (i + 0.5) / 4.0, 1.0 - (j + 0.5) / 4.0
            exp = np.array(REF_IMG_DATA[j * 5 + i], dtype=np.float64) * scale
            # the texture repeats with period 1 in u and v, and a negative u mirrors: |fmod(-x, 1)| = x
            for uu, vv in ((u, v), (u + 3.0, v), (u, v + 2.0), (-u, v)):
                got = ow.texture_value(tex, uu, vv, (0, 0, 0))
                assert np.array_equal(np.asarray(got), exp), (i, j, uu, vv)
    # integer v selects the last row (v' = 1), u = 0 the first column
    assert np.array_equal(np.asarray(ow.texture_value(tex, 0.0, 2.0, (0, 0, 0))), np.array(REF_IMG_DATA[20], dtype=np.float64) * scale)
