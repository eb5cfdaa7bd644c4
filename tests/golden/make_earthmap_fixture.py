"""Generates tests/golden/earthmap_rgb8.npz from the reference's earthmap.jpg (main.go:143,223).
Run in the build container (the GPU box has no /root/reference).  Decoder: Pillow/libjpeg — Go's image/jpeg
differs from it by <= 3/255 per channel, so these texels are "Go's texture within 3 LSB"."""
import numpy as np
from PIL import Image
im = np.asarray(Image.open("/root/reference/earthmap.jpg").convert("RGB"), dtype=np.uint8)
np.savez_compressed("tests/golden/earthmap_rgb8.npz", rgb=im)
print(im.shape, im.mean())
