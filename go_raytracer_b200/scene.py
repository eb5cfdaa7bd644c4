"""Host-side mirror of package `hittable` and of main.go's scene functions.

Names and argument meaning follow the reference (internal/hittable/*.go,
main.go:19-409); objects are integer handles into a host scene owned by
libgrt_cuda's host layer (csrc/scene_ir.hpp), which also does BuildBVH, instance
baking and flattening (csrc/flatten.hpp).
"""
import ctypes as C
import numpy as np
from . import _native as N

PERLIN, MARBLE, TURBULENT = 1, 2, 3   # texture.go:93-96


def _d3(v):
    a = (C.c_double * 3)(*[float(x) for x in v])
    return a


class Scene:
    """Owns one host scene (the arguments of Camera.Render: world + lights)."""

    def __init__(self):
        self._L = N.lib()
        self._h = self._L.grt_host_scene_new()
        self._keep = []
        self.world = None
        self.lights = None

    def __del__(self):
        try:
            if self._h:
                self._L.grt_host_scene_free(self._h)
                self._h = None
        except Exception:
            pass

    # ---- textures (texture.go) ------------------------------------------
    def NewSolidColor(self, albedo):
        return N.host_check(self._L.grt_host_solid_color(self._h, *[float(x) for x in albedo]))

    def NewCheckerboard(self, scale, even, odd):
        return N.host_check(self._L.grt_host_checkerboard(self._h, float(scale), even, odd))

    def NewCheckerboardColors(self, scale, even, odd):
        return self.NewCheckerboard(scale, self.NewSolidColor(even), self.NewSolidColor(odd))

    def NewImageTextureFromArray(self, rgb):
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        h, w, _ = rgb.shape
        img = N.host_check(self._L.grt_host_image(self._h, w, h, rgb.ctypes.data))
        return N.host_check(self._L.grt_host_image_texture(self._h, img))

    def NewImageTexture(self, filename):
        """hittable.NewImageTexture(filename) (texture.go:66-68 -> imageLoader.LoadImage): JPEG files are decoded with
        Go's image/jpeg arithmetic (csrc/jpeg_go.hpp), so the texels are the reference's."""
        return self.NewImageTextureFromArray(load_image(filename))

    def NewNoiseTextureWithType(self, scale, variant, seed=1):
        return N.host_check(self._L.grt_host_noise_texture(self._h, float(scale), int(variant), int(seed)))

    # ---- materials (materials.go) -----------------------------------------
    def NewTexturedLambertian(self, tex):
        return N.host_check(self._L.grt_host_lambertian(self._h, tex))

    def NewLambertian(self, albedo):
        return self.NewTexturedLambertian(self.NewSolidColor(albedo))

    def NewMetal(self, albedo, fuzz):
        return N.host_check(self._L.grt_host_metal(self._h, *[float(x) for x in albedo], float(fuzz)))

    def NewDielectric(self, ior):
        return N.host_check(self._L.grt_host_dielectric(self._h, float(ior)))

    def NewDiffuseLightTextured(self, tex):
        return N.host_check(self._L.grt_host_diffuse_light(self._h, tex))

    def NewDiffuseLight(self, color):
        return self.NewDiffuseLightTextured(self.NewSolidColor(color))

    def NewIsotropicTexture(self, tex):
        return N.host_check(self._L.grt_host_isotropic(self._h, tex))

    def NewIsotropic(self, albedo):
        return self.NewIsotropicTexture(self.NewSolidColor(albedo))

    # ---- hittables ------------------------------------------------------------
    def NewSphere(self, center, radius, mat):
        return N.host_check(self._L.grt_host_sphere(self._h, _d3(center), float(radius), mat))

    def NewMotionSphere(self, c1, c2, radius, mat):
        return N.host_check(self._L.grt_host_motion_sphere(self._h, _d3(c1), _d3(c2), float(radius), mat))

    def NewQuad(self, Q, u, v, mat):
        return N.host_check(self._L.grt_host_quad(self._h, _d3(Q), _d3(u), _d3(v), mat))

    def NewBox(self, a, b, mat):
        return N.host_check(self._L.grt_host_box(self._h, _d3(a), _d3(b), mat))

    def _tri(self, verts, normals, uvs, mat):
        v = (C.c_double * 9)(*[float(x) for p in verts for x in p])
        n = (C.c_double * 9)(*[float(x) for p in normals for x in p]) if normals is not None else None
        t = (C.c_double * 6)(*[float(x) for p in uvs for x in p]) if uvs is not None else None
        return N.host_check(self._L.grt_host_triangle(self._h, v, n, t, mat))

    def NewTriangle(self, verts, mat):
        return self._tri(verts, None, None, mat)

    def NewTriangleWithNormals(self, verts, normals, mat):
        return self._tri(verts, normals, None, mat)

    def NewTexturedTriangle(self, verts, uvs, mat):
        return self._tri(verts, None, uvs, mat)

    def NewTexturedTriangleWithNormals(self, verts, normals, uvs, mat):
        return self._tri(verts, normals, uvs, mat)

    def NewHittableList(self, objs=()):
        l = N.host_check(self._L.grt_host_list(self._h))
        for o in objs:
            self.Add(l, o)
        return l

    def Add(self, lst, obj):
        N.host_check(self._L.grt_host_list_add(self._h, lst, obj))

    def BuildBVH(self, lst):
        return N.host_check(self._L.grt_host_bvh(self._h, lst))

    def Translate(self, obj, offset):
        return N.host_check(self._L.grt_host_translate(self._h, obj, _d3(offset)))

    def RotateY(self, obj, degrees):
        return N.host_check(self._L.grt_host_rotate_y(self._h, obj, float(degrees)))

    def ConstantMediumTexture(self, boundary, density, tex):
        return N.host_check(self._L.grt_host_constant_medium(self._h, boundary, float(density), tex))

    def ConstantMedium(self, boundary, density, albedo):
        return self.ConstantMediumTexture(boundary, density, self.NewSolidColor(albedo))

    def LoadObjWithOptions(self, obj_text, mtl_text=None, ScaleFactor=1.0, FlipYZ=False, IgnoreNormals=False, Center=True,
                           FlipFaces=False, Position=(0, 0, 0), DefaultMaterial=-1, IgnoreMtl=False, FindWindows=False):
        """objLoader.LoadObjWithOptions on text (objLoader.go:72).  Returns (model BVH, lights list, n triangles)."""
        o = N.GrtObjOptions()
        o.ScaleFactor = float(ScaleFactor)
        o.FlipYZ, o.IgnoreNormals, o.Center, o.FlipFaces = int(FlipYZ), int(IgnoreNormals), int(Center), int(FlipFaces)
        o.IgnoreMtl, o.FindWindows, o.DefaultMaterial = int(IgnoreMtl), int(FindWindows), int(DefaultMaterial)
        for i in range(3):
            o.Position[i] = float(Position[i])
        model, lights, ntri = C.c_int(-1), C.c_int(-1), C.c_int(0)
        N.host_check(self._L.grt_host_load_obj(self._h, obj_text.encode(), mtl_text.encode() if mtl_text else None, C.byref(o),
                                               C.byref(model), C.byref(lights), C.byref(ntri)))
        return model.value, lights.value, ntri.value

    def LoadObj(self, path, **options):
        """objLoader.LoadObjWithOptions(filename, options) on a file (objLoader.go:72): the `mtllib` the OBJ names is
        read from the same directory.  Returns (model BVH, lights list, n triangles)."""
        o = N.GrtObjOptions()
        o.ScaleFactor = float(options.get("ScaleFactor", 1.0))
        for k in ("FlipYZ", "IgnoreNormals", "FlipFaces", "IgnoreMtl", "FindWindows"):
            setattr(o, k, int(options.get(k, False)))
        o.Center = int(options.get("Center", True))
        o.DefaultMaterial = int(options.get("DefaultMaterial", -1))
        for i, v in enumerate(options.get("Position", (0, 0, 0))):
            o.Position[i] = float(v)
        model, lights, ntri = C.c_int(-1), C.c_int(-1), C.c_int(0)
        N.host_check(self._L.grt_host_load_obj_file(self._h, str(path).encode(), C.byref(o), C.byref(model), C.byref(lights), C.byref(ntri)))
        return model.value, lights.value, ntri.value

    def set_world(self, obj):
        N.host_check(self._L.grt_host_set_world(self._h, obj))
        self.world = obj

    def set_lights(self, obj):
        N.host_check(self._L.grt_host_set_lights(self._h, obj))
        self.lights = obj

    # ---- flatten / description ---------------------------------------------------
    def flatten(self, collapse_whole=None, collapse_leaf=None):
        """hittable.Flatten: returns the GrtScene view (valid while this Scene lives, until the next flatten).
        collapse_* tune how much of each BVH is emitted as ordered leaf runs (results do not depend on them)."""
        s = N.GrtScene()
        if collapse_whole is None and collapse_leaf is None:
            N.host_check(self._L.grt_host_flatten(self._h, C.byref(s)))
        else:
            N.host_check(self._L.grt_host_flatten_opts(self._h, int(collapse_whole or 0), int(collapse_leaf or 0), C.byref(s)))
        return s

    def description_ptr(self):
        """Opaque pointer for the test oracle (tests / bench cpu baseline only)."""
        return self._L.grt_host_scene_description(self._h)


def decode_jpeg(data):
    """imageLoader.LoadImage on JPEG bytes -> uint8 [H, W, 3], texel for texel what Go's image/jpeg + color.YCbCr give."""
    L = N.lib()
    buf = (C.c_ubyte * len(data)).from_buffer_copy(data)
    w, h = C.c_int(0), C.c_int(0)
    N.host_check(L.grt_host_decode_jpeg(buf, len(data), C.byref(w), C.byref(h), None, 0))
    out = np.zeros((h.value, w.value, 3), dtype=np.uint8)
    N.host_check(L.grt_host_decode_jpeg(buf, len(data), C.byref(w), C.byref(h), out.ctypes.data, out.nbytes))
    return out


def load_image(filename):
    with open(filename, "rb") as f:
        return decode_jpeg(f.read())


def builtin_scene(scene_id, width=0, spp=0, aspect=0.0, seed=0, mesh_segments=0, image=None):
    """main.go's scene functions (-S 1..8).  Returns (Scene, GrtCameraConfig)."""
    s = Scene()
    opt = N.GrtSceneOptions()
    opt.width, opt.spp, opt.aspect, opt.seed, opt.mesh_segments = int(width), int(spp), float(aspect), int(seed), int(mesh_segments)
    if image is not None:
        image = np.ascontiguousarray(image, dtype=np.uint8)
        s._keep.append(image)
        opt.image_h, opt.image_w = image.shape[0], image.shape[1]
        opt.image_rgb = image.ctypes.data
    cfg = N.GrtCameraConfig()
    N.host_check(s._L.grt_host_builtin_scene(s._h, int(scene_id), C.byref(opt), C.byref(cfg)))
    return s, cfg


SCENE_NAMES = {1: "book1", 2: "book2", 3: "book3", 4: "simpleLight", 5: "quads", 6: "cornellBox", 7: "cornellSmoke",
               8: "modelExample"}
