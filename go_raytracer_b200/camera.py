"""Mirror of internal/camera: Camera's public fields, PositionCamera and
Render(world, lights), with the row goroutines replaced by the CUDA backend."""
import ctypes as C
import io
import numpy as np
from . import _native as N


def derive_camera(cfg):
    """Camera.initialize (camera.go:179-253): public fields -> derived GrtCamera."""
    cam = N.GrtCamera()
    N.host_check(N.lib().grt_host_camera_derive(C.byref(cfg), C.byref(cam)))
    return cam


class DeviceScene:
    """A flattened scene resident in HBM on one device."""

    def __init__(self, scene, device=0, collapse_whole=None, collapse_leaf=None):
        self._L = N.lib()
        self.scene = scene
        flat = scene.flatten(collapse_whole, collapse_leaf)
        h = C.c_void_p()
        N.check(self._L.grt_scene_upload(C.byref(flat), int(device), C.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.grt_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trace_batch(self, rays):
        """World.Hit for a batch of rays (numpy RAY_DTYPE) -> numpy HIT_DTYPE."""
        rays = np.ascontiguousarray(rays, dtype=N.RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=N.HIT_DTYPE)
        N.check(self._L.grt_trace_batch(self._h, rays.ctypes.data, len(rays), hits.ctypes.data))
        return hits

    def render(self, cam, seed=0xC0FFEE, variant=N.GRT_VARIANT_MEGAKERNEL, sample_first=0, sample_stride=1,
               window=None, want_rgb8=False, want_stats=False):
        """grt_render with host buffers.  Returns (rgb_sum[H,W,3] float32, rgb8 or None, stats or None)."""
        opt = N.GrtOptions()
        opt.seed, opt.variant, opt.device = int(seed), int(variant), int(self.device)
        opt.sample_first, opt.sample_stride = int(sample_first), int(sample_stride)
        if window is not None:
            opt.x0, opt.y0, opt.x1, opt.y1 = [int(v) for v in window]
        sums = np.zeros((cam.height, cam.width, 3), dtype=np.float32)
        rgb8 = np.zeros((cam.height, cam.width, 3), dtype=np.uint8) if want_rgb8 else None
        stats = N.GrtStats() if want_stats else None
        N.check(self._L.grt_render(self._h, C.byref(cam), C.byref(opt), sums.ctypes.data,
                                   rgb8.ctypes.data if want_rgb8 else None, C.byref(stats) if want_stats else None))
        return sums, rgb8, (stats.as_dict() if want_stats else None)

    def render_device(self, cam, d_rgb_sum_ptr, stream_ptr=0, seed=0xC0FFEE, variant=N.GRT_VARIANT_MEGAKERNEL,
                      sample_first=0, sample_stride=1, d_stats_ptr=0, flags=0):
        """grt_render_device: accumulate into a device buffer, asynchronously on `stream_ptr`."""
        opt = N.GrtOptions()
        opt.seed, opt.variant, opt.device = int(seed), int(variant), int(self.device)
        opt.sample_first, opt.sample_stride = int(sample_first), int(sample_stride)
        opt.flags = int(flags)
        if d_stats_ptr:
            opt.flags |= N.GRT_OPT_STATS
        N.check(self._L.grt_render_device(self._h, C.byref(cam), C.byref(opt), C.c_void_p(d_rgb_sum_ptr),
                                          C.c_void_p(stream_ptr), C.c_void_p(d_stats_ptr)))

    def tonemap_device(self, d_rgb_sum_ptr, d_rgb8_ptr, n_values, scale, stream_ptr=0):
        N.check(self._L.grt_tonemap_device(C.c_void_p(d_rgb_sum_ptr), C.c_void_p(d_rgb8_ptr), int(n_values),
                                           float(scale), C.c_void_p(stream_ptr)))


def write_p6(rgb8):
    """Binary PPM of the same pixels."""
    rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
    h, w, _ = rgb8.shape
    buf = C.create_string_buffer(32 + w * h * 3)
    n = N.lib().grt_host_write_p6(rgb8.ctypes.data, w, h, buf, len(buf))
    if n < 0:
        raise ValueError("grt_host_write_p6 failed")
    return buf.raw[:n]


def write_png(rgb8):
    """PNG (8-bit RGB, stored deflate blocks) of the same pixels."""
    rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
    h, w, _ = rgb8.shape
    buf = C.create_string_buffer(256 + (3 * w + 1) * h + 5 * ((3 * w + 1) * h // 65535 + 2))
    n = N.lib().grt_host_write_png(rgb8.ctypes.data, w, h, buf, len(buf))
    if n < 0:
        raise ValueError("grt_host_write_png failed")
    return buf.raw[:n]


def write_ppm(rgb8):
    """P3 text exactly as camera.go:160 + color.go:45 produce it."""
    rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
    h, w, _ = rgb8.shape
    cap = 32 + w * h * 12
    buf = C.create_string_buffer(cap)
    n = N.lib().grt_host_write_ppm(rgb8.ctypes.data, w, h, buf, cap)
    if n < 0:
        raise N.GrtError(-1, "ppm buffer too small")
    return buf.raw[:n]


class Camera:
    """camera.Camera: same public fields, PositionCamera and Render(world, lights)."""

    def __init__(self):
        self.AspectRatio = 0.0
        self.Width = 0
        self.Out = None
        self.SamplesPerPixel = 0
        self.MaxDepth = 0
        self.MaxThreads = 1          # -N: meaningless for the GPU backend, kept for drop-in
        self.VerticalFOV = 0.0
        self.DefocusAngle = 0.0
        self.FocusDistance = 0.0
        self.Background = (0.0, 0.0, 0.0)
        self.MaxContribution = 0.0
        self._from, self._at, self._up = (0, 0, 0), (0, 0, -1), (0, 1, 0)
        # backend knobs (not in the reference)
        self.Seed = 0xC0FFEE
        self.Gpus = 1
        self.Variant = N.GRT_VARIANT_AUTO

    def PositionCamera(self, lookFrom=None, lookAt=None, vup=None):   # camera.go:65-81
        self._from = tuple(lookFrom) if lookFrom is not None else (0, 0, 0)
        self._at = tuple(lookAt) if lookAt is not None else (0, 0, -1)
        self._up = tuple(vup) if vup is not None else (0, 1, 0)

    @classmethod
    def from_config(cls, cfg):
        c = cls()
        for f in ("AspectRatio", "Width", "SamplesPerPixel", "MaxDepth", "MaxThreads", "VerticalFOV", "DefocusAngle",
                  "FocusDistance", "MaxContribution"):
            setattr(c, f, getattr(cfg, f))
        c.Background = tuple(cfg.Background)
        c.PositionCamera(tuple(cfg.lookFrom), tuple(cfg.lookAt), tuple(cfg.vup))
        return c

    def config(self):
        cfg = N.GrtCameraConfig()
        cfg.AspectRatio, cfg.Width, cfg.SamplesPerPixel = float(self.AspectRatio), int(self.Width), int(self.SamplesPerPixel)
        cfg.MaxDepth, cfg.MaxThreads = int(self.MaxDepth), int(self.MaxThreads)
        cfg.VerticalFOV, cfg.DefocusAngle, cfg.FocusDistance = float(self.VerticalFOV), float(self.DefocusAngle), float(self.FocusDistance)
        cfg.MaxContribution = float(self.MaxContribution)
        for i in range(3):
            cfg.Background[i] = float(self.Background[i])
            cfg.lookFrom[i], cfg.lookAt[i], cfg.vup[i] = float(self._from[i]), float(self._at[i]), float(self._up[i])
        return cfg

    def Render(self, scene, world=None, lights=None):
        """camera.go:156: renders `world` lit by `lights` and writes P3 text to self.Out."""
        if self.Out is None:
            raise ValueError("Must specify an output")          # camera.go:187-189 (log.Fatal)
        if world is not None:
            scene.set_world(world)
        if lights is not None:
            scene.set_lights(lights)
        cfg = self.config()
        cam = derive_camera(cfg)
        cap = 32 + cam.width * cam.height * 12
        buf = C.create_string_buffer(cap)
        n = C.c_long(0)
        ms = C.c_double(0)
        L = N.lib()
        rc = L.grt_host_camera_render(scene._h, C.byref(cfg), int(self.Seed), int(self.Variant), int(self.Gpus), None, buf, cap,
                                      C.byref(n), C.byref(ms))
        N.check(rc)
        data = buf.raw[:n.value]
        if isinstance(self.Out, io.TextIOBase):
            self.Out.write(data.decode())
        else:
            self.Out.write(data)
