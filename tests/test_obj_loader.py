"""internal/objLoader mirror (csrc/obj_loader.hpp): hand-checked cases for the parsing rules of objLoader.go and
the material heuristics of mtlLoader.go.  CPU only."""
import ctypes as C
import numpy as np
import pytest
import go_raytracer_b200 as g
from go_raytracer_b200 import _native as N


def as_np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.frombuffer((C.c_uint8 * (n * dtype.itemsize)).from_address(ptr), dtype=dtype).copy()


def load(obj, mtl=None, **kw):
    sc = g.Scene()
    model, lights, ntri = sc.LoadObjWithOptions(obj, mtl, **kw)
    sc.set_world(sc.NewHittableList([model]))
    sc.set_lights(lights)
    flat = sc.flatten(0, 0)
    tris = as_np(flat.tris, flat.n_tris, N.TRI_DTYPE)
    mats = as_np(flat.materials, flat.n_materials, N.MATERIAL_DTYPE)
    texs = as_np(flat.textures, flat.n_textures, N.TEXTURE_DTYPE)
    order = np.argsort(tris["id"])          # creation order of the triangles
    return sc, flat, tris[order], mats, texs, ntri


def verts(t):
    return np.stack([t["v0"], t["v0"] + t["e0"], t["v0"] + t["e1"]])


QUAD = "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\n"


def test_fan_triangulation_and_index_forms():
    # objLoader.go:396-398: (v[0], v[i-1], v[i]); negative indices count from the end (fixIndex :47-61)
    _, flat, tris, _, _, ntri = load(QUAD + "f 1 2 3 4\n", Center=False)
    assert ntri == 2
    assert np.allclose(verts(tris[0]), [(0, 0, 0), (1, 0, 0), (1, 1, 0)])
    assert np.allclose(verts(tris[1]), [(0, 0, 0), (1, 1, 0), (0, 1, 0)])
    _, _, t2, _, _, _ = load(QUAD + "f -4 -3 -2 -1\n", Center=False)
    assert np.allclose(verts(t2[0]), verts(tris[0])) and np.allclose(verts(t2[1]), verts(tris[1]))
    # out-of-range indices are clamped, not fatal
    _, _, t3, _, _, _ = load(QUAD + "f 1 2 99\n", Center=False)
    assert np.allclose(verts(t3[0])[2], (0, 1, 0))


def test_scale_flip_center_position():
    # scale first, then flipYZ, bounds on the scaled values, then -centre +Position (objLoader.go:188-250)
    obj = "v 0 0 0\nv 2 0 0\nv 0 4 0\nf 1 2 3\n"
    _, _, t, _, _, _ = load(obj, ScaleFactor=0.5, Center=True, Position=(10, 20, 30))
    assert np.allclose(verts(t[0]), np.array([(0, 0, 0), (1, 0, 0), (0, 2, 0)]) - (0.5, 1, 0) + (10, 20, 30))
    _, _, t, _, _, _ = load(obj, FlipYZ=True, Center=False)
    assert np.allclose(verts(t[0]), [(0, 0, 0), (2, 0, 0), (0, 0, 4)])
    _, _, t, _, _, _ = load(obj, FlipFaces=True, Center=False)
    assert np.allclose(verts(t[0]), [(0, 0, 0), (0, 4, 0), (2, 0, 0)])
    # Position is ignored without Center (objLoader.go:241-247)
    _, _, t, _, _, _ = load(obj, Center=False, Position=(5, 5, 5))
    assert np.allclose(verts(t[0])[0], (0, 0, 0))


def test_normals_and_texcoords():
    obj = QUAD + "vn 0 0 2\nvn 0 3 0\nvt 0 0\nvt 1 0\nvt 1 1\nf 1/1/1 2/2/1 3/3/2\nf 1//1 2//1 3//2\nf 1/1 2/2 3/3\nf 1 2 3\n"
    keep, flat, tris, _, _, _ = load(obj, Center=False)     # `keep` owns the arrays `flat` points into
    assert list(tris["flags"]) == [3, 1, 2, 0]            # uv+normals, normals, uv, plain (objLoader.go:406-465)
    sh = as_np(flat.tri_shade, flat.n_tris, np.dtype([("n", "<f4", 9), ("uv", "<f4", 6), ("pad", "<f4")]))
    first = int(np.argmin(as_np(flat.tris, flat.n_tris, N.TRI_DTYPE)["id"]))
    assert np.allclose(sh[first]["n"], (0, 0, 1, 0, 0, 1, 0, 1, 0))    # vn is normalised (objLoader.go:312-316)
    assert np.allclose(sh[first]["uv"], (0, 0, 1, 0, 1, 1))
    _, _, t2, _, _, _ = load(obj, Center=False, IgnoreNormals=True)
    assert list(t2["flags"]) == [2, 0, 2, 0]


MTL = """
newmtl lamp
Ke 4 4 4
newmtl glass
d 0.2
Ni 1.45
newmtl fog
d 0.5
Kd 0.1 0.2 0.3
newmtl steel
Ks 0.9 0.9 0.9
Kd 0.1 0.1 0.1
Ns 500
newmtl paint
Kd 0.6 0.5 0.4
newmtl mirror
illum 3
Ks 0.05 0.05 0.05
"""


def test_mtl_material_heuristics_and_light_list():
    obj = "mtllib scene.mtl\n" + QUAD + "".join(f"usemtl {m}\nf 1 2 3\n" for m in ("lamp", "glass", "fog", "steel", "paint", "mirror", "nosuch"))
    sc, flat, tris, mats, texs, ntri = load(obj, MTL, Center=False)
    assert ntri == 7
    m = mats[tris["mat"]]
    assert list(m["type"]) == [N.GRT_OK + 3, 2, 4, 1, 0, 1, 0]      # light, dielectric, isotropic, metal, lambertian, metal, default
    assert m["ior"][1] == pytest.approx(1.45)
    assert np.allclose(texs[m["tex"][2]]["color"], (0.1, 0.2, 0.3))
    assert m["fuzz"][3] == pytest.approx((1 - 500 / 1000) ** 2) and np.allclose(m["albedo"][3], 0.9)   # mtlLoader.go:274-296
    assert np.allclose(texs[m["tex"][4]]["color"], (0.6, 0.5, 0.4))
    assert m["fuzz"][5] == pytest.approx(0.3)                         # illum 3 (mtlLoader.go:311-313)
    assert np.allclose(texs[m["tex"][6]]["color"], 0.8)               # unknown material -> default Lambertian(0.8)
    assert flat.n_lights == 1 and flat.lights_mode == 0               # only the emissive triangle (objLoader.go:492-510)
    # FindWindows adds the dielectric triangle too
    sc2 = g.Scene()
    _, lights, _ = sc2.LoadObjWithOptions(obj, MTL, Center=False, FindWindows=True)
    sc2.set_world(sc2.NewHittableList()); sc2.set_lights(lights)
    assert sc2.flatten().n_lights == 2
    # the library is only consulted when the OBJ names one (objLoader.go:105-133), and IgnoreMtl skips it
    _, _, t3, m3, _, _ = load(obj.replace("mtllib scene.mtl\n", ""), MTL, Center=False)
    assert (m3[t3["mat"]]["type"] == 0).all()
    _, _, t4, m4, _, _ = load(obj, MTL, Center=False, IgnoreMtl=True)
    assert (m4[t4["mat"]]["type"] == 0).all()


def test_errors_are_codes():
    sc = g.Scene()
    with pytest.raises(g.GrtError, match="No triangles"):
        sc.LoadObjWithOptions("v 0 0 0\nv 1 0 0\n")                  # objLoader.go:485-487 (log.Fatalf)
    with pytest.raises(g.GrtError, match="Could not open"):
        sc.LoadObjWithOptions("mtllib m.mtl\n" + QUAD + "usemtl t\nf 1 2 3\n", "newmtl t\nmap_Kd wood.jpg\n")


def test_model_example_goes_through_the_loader():
    # config C5's mesh is OBJ text run through the loader: triangle count = 2 per quad face, vertex normals kept
    s, cfg = g.builtin_scene(8, mesh_segments=16)
    flat = s.flatten()
    assert flat.n_tris == 2 * 16 * 16 and flat.tri_shade
    tris = as_np(flat.tris, flat.n_tris, N.TRI_DTYPE)
    assert (tris["flags"] == 1).all()
    # centred on Position (0, 1.8, 0), then RotateY(180) about the origin (main.go:380-384)
    p = np.concatenate([tris["v0"], tris["v0"] + tris["e0"], tris["v0"] + tris["e1"]])
    assert np.allclose((p.min(0) + p.max(0)) / 2, (0, 1.8, 0), atol=1e-5)


def test_load_from_file_resolves_mtllib_next_to_the_obj(tmp_path):
    """objLoader.go:84-140: the file variant reads the OBJ, takes the first `mtllib` (fields joined by one space) from
    the OBJ's directory, and carries on with the default material when that file is missing."""
    (tmp_path / "sub").mkdir()
    (tmp_path / "sub" / "my lib.mtl").write_text("newmtl red\nKd 1 0 0\n")
    objp = tmp_path / "sub" / "m.obj"
    objp.write_text("# comment\n  mtllib   my   lib.mtl  \n" + QUAD + "usemtl red\nf 1 2 3\n")
    sc = g.Scene()
    model, lights, ntri = sc.LoadObj(objp, Center=False)
    sc.set_world(sc.NewHittableList([model])); sc.set_lights(lights)
    flat = sc.flatten(0, 0)
    tris = as_np(flat.tris, flat.n_tris, N.TRI_DTYPE)
    mats = as_np(flat.materials, flat.n_materials, N.MATERIAL_DTYPE)
    texs = as_np(flat.textures, flat.n_textures, N.TEXTURE_DTYPE)
    assert ntri == 1
    m = mats[tris[0]["mat"]]
    assert np.allclose(texs[m["tex"]]["color"], (1, 0, 0))          # Kd of the library next to the OBJ
    # same result as the text entry point
    sc2, flat2, tris2, mats2, texs2, _ = load("mtllib x.mtl\n" + QUAD + "usemtl red\nf 1 2 3\n", "newmtl red\nKd 1 0 0\n", Center=False)
    assert np.allclose(verts(tris[0]), verts(tris2[0]))
    # a missing library is not an error (objLoader.go:136-139)
    obj2 = tmp_path / "n.obj"
    obj2.write_text("mtllib nowhere.mtl\n" + QUAD + "f 1 2 3\n")
    sc3 = g.Scene()
    assert sc3.LoadObj(obj2)[2] == 1
    # a missing OBJ is (objLoader.go:84-87)
    with pytest.raises(N.GrtError):
        g.Scene().LoadObj(tmp_path / "absent.obj")
