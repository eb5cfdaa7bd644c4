"""Hand-derived known-answer cases for the oracle's render-path restatement (the reference has no tests for
camera/hittable/aabb: PARITY UNPINNED by the reference, pinned here by arithmetic one can do on paper)."""
import math
import numpy as np
import pytest
import go_raytracer_b200 as g
from oracle import oracle_py as O
import parity_util as PU


def trace(scene, o, d, **kw):
    ow = O.OracleWorld(scene)
    return ow.trace_batch(PU.make_rays(np.atleast_2d(o), np.atleast_2d(d), **kw), audit_eps=0)


def cornell():
    return g.builtin_scene(6)


def test_cornell_axis_ray_hits_back_wall():
    s, cfg = cornell()
    # y = 500 is above both boxes (heights 330 and 165), so the first surface along +z is the back wall z = 555
    h = trace(s, (278, 500, -800), (0, 0, 1))[0]
    assert h["t"] == pytest.approx(1355.0, abs=1e-9)
    # the back wall's u x v = (555,0,0) x (0,555,0) points to +z: the camera sees its BACK face, and
    # setFaceNormal (hittable.go:27-34) flips the normal against the ray
    assert h["front_face"] == 0 and np.allclose(h["n"], (0, 0, -1))
    assert np.allclose(h["p"], (278, 500, 555))
    # alpha/beta of the back wall quad Q=(0,0,555) u=(555,0,0) v=(0,555,0)
    assert h["u"] == pytest.approx(278 / 555) and h["v"] == pytest.approx(500 / 555)


def test_cornell_light_and_ceiling():
    s, _ = cornell()
    up = trace(s, (278, 100, 279.5), (0, 1, 0))[0]        # light quad at y=550 (main.go:295) spans x 213..343, z 227..332
    assert up["t"] == pytest.approx(450.0) and up["front_face"] == 1    # light normal (0,-1,0)... u x v = (-130,0,0)x(0,0,-105) = (0,-13650,0)
    side = trace(s, (100, 100, 279.5), (0, 1, 0))[0]      # misses the light, hits the ceiling y=555
    assert side["t"] == pytest.approx(455.0)
    assert up["id"] != side["id"]


def test_quad_closed_edges_and_tie_order():
    sc = g.Scene()
    m = sc.NewLambertian((.5, .5, .5))
    q1 = sc.NewQuad((0, 0, 0), (1, 0, 0), (0, 1, 0), m)
    q2 = sc.NewQuad((0, 0, 0), (1, 0, 0), (0, 1, 0), m)   # coincident: equal t
    sc.set_world(sc.NewHittableList([q1, q2])); sc.set_lights(sc.NewHittableList())
    h = trace(sc, [(0, 0, -1), (1, 1, -1), (1.0000001, .5, -1), (.5, .5, -1)], [(0, 0, 1)] * 4)
    assert h["id"][0] == q2 and h["id"][1] == q2      # closed [0,1]^2 (objects.go:199); later equal-t quad REPLACES (closed Contains)
    assert h["id"][2] == -1
    assert h["id"][3] == q2 and h["t"][3] == 1.0


def test_sphere_open_interval_first_wins_and_inside():
    sc = g.Scene()
    m = sc.NewLambertian((.5, .5, .5))
    s1 = sc.NewSphere((0, 0, 0), 1, m)
    s2 = sc.NewSphere((0, 0, 0), 1, m)
    sc.set_world(sc.NewHittableList([s1, s2])); sc.set_lights(sc.NewHittableList())
    h = trace(sc, [(0, 0, -3), (0, 0, 0), (0, 2, -3)], [(0, 0, 1)] * 3)
    assert h["id"][0] == s1 and h["t"][0] == pytest.approx(2.0)     # Surrounds is open: the equal-t second sphere does not replace
    assert h["front_face"][0] == 1 and np.allclose(h["n"][0], (0, 0, -1))
    assert h["t"][1] == pytest.approx(1.0) and h["front_face"][1] == 0 and np.allclose(h["n"][1], (0, 0, -1))   # from inside
    assert h["id"][2] == -1
    # sphere UV (objects.go:44-50): p=(0,0,-1): theta=acos(0)=pi/2, phi=atan2(1,0)+pi=3pi/2 -> u=.75 v=.5
    assert h["u"][0] == pytest.approx(0.75) and h["v"][0] == pytest.approx(0.5)


def test_unnormalised_direction_scales_t():
    sc = g.Scene()
    m = sc.NewLambertian((.5, .5, .5))
    sc.set_world(sc.NewHittableList([sc.NewQuad((-1, -1, 5), (2, 0, 0), (0, 2, 0), m)])); sc.set_lights(sc.NewHittableList())
    h = trace(sc, [(0, 0, 0), (0, 0, 0), (0, 0, 4.9995)], [(0, 0, 1), (0, 0, 10), (0, 0, 1)])
    assert h["t"][0] == pytest.approx(5.0) and h["t"][1] == pytest.approx(0.5)
    assert h["id"][2] == -1                              # t = 0.0005 < tmin 0.001 (camera.go:300)


def test_triangle_moller_trumbore_and_vertex_normals():
    sc = g.Scene()
    m = sc.NewLambertian((.5, .5, .5))
    t1 = sc.NewTriangle([(0, 0, 0), (1, 0, 0), (0, 1, 0)], m)
    n = [(0, 0, 1), (1, 0, 1), (0, 1, 1)]
    t2 = sc.NewTriangleWithNormals([(0, 0, 2), (1, 0, 2), (0, 1, 2)], n, m)
    sc.set_world(sc.NewHittableList([t1, t2])); sc.set_lights(sc.NewHittableList())
    h = trace(sc, [(.25, .25, 5), (.25, .25, 1), (.6, .6, 5)], [(0, 0, -1), (0, 0, 1), (0, 0, -1)])
    assert h["id"][0] == t2 and h["t"][0] == pytest.approx(3.0)
    assert h["u"][0] == pytest.approx(.25) and h["v"][0] == pytest.approx(.25)     # default UVs = barycentrics (objects.go:443-445)
    exp = np.array([.25, .25, 1.0]); exp /= np.linalg.norm(exp)                     # w n0 + u n1 + v n2, normalised
    assert np.allclose(h["n"][0], exp) and h["front_face"][0] == 1
    assert h["id"][1] == t2 and h["front_face"][1] == 0 and np.allclose(h["n"][1], -exp)
    assert h["id"][2] == -1                                                        # u+v > 1


def test_translate_rotate_instances():
    sc = g.Scene()
    m = sc.NewLambertian((.5, .5, .5))
    box = sc.NewBox((0, 0, 0), (1, 1, 1), m)
    inst = sc.Translate(sc.RotateY(box, 90), (10, 0, 0))
    sc.set_world(sc.NewHittableList([inst])); sc.set_lights(sc.NewHittableList())
    # rotateY(+90): object (x,y,z) -> world (z, y, -x) (transformation.go:87-93), so the unit cube occupies
    # world x in [10,11], z in [-1,0]
    h = trace(sc, [(10.5, .5, 5), (10.5, .5, -5), (9.5, .5, 5)], [(0, 0, -1), (0, 0, 1), (0, 0, -1)])
    assert h["t"][0] == pytest.approx(5.0) and np.allclose(h["n"][0], (0, 0, 1), atol=1e-12)
    assert h["t"][1] == pytest.approx(4.0) and np.allclose(h["n"][1], (0, 0, -1), atol=1e-12)
    assert h["id"][2] == -1


def test_motion_sphere_uses_ray_time():
    sc = g.Scene()
    m = sc.NewLambertian((.5, .5, .5))
    sc.set_world(sc.NewHittableList([sc.NewMotionSphere((0, 0, 0), (0, 4, 0), 1, m)])); sc.set_lights(sc.NewHittableList())
    h = trace(sc, [(0, 0, -5), (0, 2, -5), (0, 2, -5)], [(0, 0, 1)] * 3, time=[0.0, 0.5, 0.0])
    assert h["t"][0] == pytest.approx(4.0) and h["t"][1] == pytest.approx(4.0) and h["id"][2] == -1


def test_bvh_span1_duplicates_and_span3_split():
    # 3 objects: span 3 -> sort on the longest axis, mid = 1: left = span-1 node (object duplicated, bvh.go:44-46)
    sc = g.Scene()
    m = sc.NewLambertian((.5, .5, .5))
    qs = [sc.NewQuad((x, 0, 0), (1, 0, 0), (0, 1, 0), m) for x in (4, 0, 2)]
    sc.set_world(sc.BuildBVH(sc.NewHittableList(qs))); sc.set_lights(sc.NewHittableList())
    ow = O.OracleWorld(sc)
    rays = PU.make_rays([(x + .5, .5, -1) for x in (0, 2, 4)], [(0, 0, 1)] * 3)
    h = ow.trace_batch(rays)
    assert list(h["id"]) == [qs[1], qs[2], qs[0]]
    # event counts through the render-loop stats: a ray at x=0.5 tests root box, left node box, the left quad TWICE
    flat = sc.flatten(collapse_whole=0, collapse_leaf=0)  # the reference's tree, node for node
    assert flat.n_nodes == 3 and flat.n_quads == 3       # the duplicated leaf is flattened once and referenced twice
    flat = sc.flatten()                                   # default: a 3-leaf tree is one ordered run
    assert flat.n_nodes == 0 and flat.n_items == 3


def test_constant_medium_scatter_statistics():
    # a slab of density rho crossed over length Lm scatters with probability 1 - exp(-rho Lm) (medium.go:46-50)
    sc = g.Scene()
    white = sc.NewLambertian((.5, .5, .5))
    box = sc.NewBox((0, 0, 0), (1, 1, 2), white)
    med = sc.ConstantMedium(box, 0.7, (1, 1, 1))
    sc.set_world(sc.NewHittableList([med])); sc.set_lights(sc.NewHittableList())
    n = 40000
    rays = PU.make_rays(np.tile((.5, .5, -1.0), (n, 1)), np.tile((0, 0, 1.0), (n, 1)))
    h = O.OracleWorld(sc).trace_batch(rays)
    frac = (h["id"] == med).mean()
    expect = 1 - math.exp(-0.7 * 2)
    assert abs(frac - expect) < 4 * math.sqrt(expect * (1 - expect) / n)
    t = h["t"][h["id"] == med]
    assert t.min() >= 1.0 and t.max() <= 3.0
    assert np.allclose(h["n"][h["id"] == med], (1, 0, 0)) and (h["front_face"][h["id"] == med] == 1).all()


def test_camera_initialize_cornell():
    _, cfg = cornell()
    d = O.derived_camera(cfg)
    assert (d.width, d.height, d.spp_sqrt, d.max_depth) == (600, 600, 10, 50)
    # viewport height 2*tan(20 deg)*10, pixel delta = that / 600; camera looks down +z so u = (-1,0,0)... vup x w with w = (0,0,-1)
    vh = 2 * math.tan(math.radians(20)) * 10
    assert d.delta_u[0] == pytest.approx(-vh / 600) and d.delta_v[1] == pytest.approx(-vh / 600)
    assert d.pixel00[2] == pytest.approx(-790.0)
    assert d.max_contribution == 1.5 and d.defocus_angle == 0
    # sppSqrt floors: 10 spp -> 3 (camera.go:211)
    _, cfg7 = g.builtin_scene(7)
    assert O.derived_camera(cfg7).spp_sqrt == 3


def test_clamp_and_direct_light_pixel():
    # one sample per pixel: a camera-visible light returns its emission UNCLAMPED (camera.go:313), every other
    # sample went through clampContribution and has R+G+B <= MaxContribution = 1.5 (camera.go:334-341, :205-207)
    s, cfg = g.builtin_scene(6, width=64, spp=1)
    ow = O.OracleWorld(s)
    sums, _, _, _ = ow.render(cfg)
    tot = sums.sum(axis=2)
    direct = np.isclose(tot, 45.0)
    assert direct.sum() > 10 and np.allclose(sums[direct], 15.0)
    assert tot[~direct].max() <= 1.5 + 1e-9
    assert (tot[~direct] > 1.4999).sum() > 20            # the clamp is active on a visible share of paths
