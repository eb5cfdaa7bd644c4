"""Generates tests/golden/earthmap_rgb8.npz from the reference's earthmap.jpg (main.go:143,223).
Run in the build container (the GPU box has no /root/reference).

Decoder: go_raytracer_b200.decode_jpeg (csrc/jpeg_go.hpp), which restates Go's image/jpeg + color.YCbCr arithmetic and
reproduces the reference's own golden texels for test.jpg exactly (tests/test_jpeg_go.py) — so these ARE the texels the
Go program renders with (round 1 used Pillow/libjpeg, within 3/255 of them)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import go_raytracer_b200 as g
im = g.load_image("/root/reference/earthmap.jpg")
np.savez_compressed("tests/golden/earthmap_rgb8.npz", rgb=im)
print(im.shape, im.mean())
try:
    from PIL import Image
    pil = np.asarray(Image.open("/root/reference/earthmap.jpg").convert("RGB"))
    d = np.abs(im.astype(int) - pil.astype(int))
    print("vs Pillow: max |diff|", d.max(), "texel-channels differing", (d > 0).mean())
except ImportError:
    pass
