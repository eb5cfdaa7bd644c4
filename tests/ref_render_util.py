"""Image-level parity against the REFERENCE'S OWN renders (tests/golden/ref_renders/*.jpg, copied from
/root/reference/readmeImgs — README.md:39-49; the Go program's output at main.go's shipped resolution).

These JPEGs are the only output of the real Go renderer that exists in this environment (no Go toolchain), so they
are what pins the oracle — and the CUDA path — to the reference rather than to each other.  A render is compared
after the reference's own tone mapping (color.go:14-46: sqrt gamma, clamp, x256) at the reference's resolution and
at a MATCHED sample count: the gamma curve is concave, so the mean of a noisy image sits below the mean of a
converged one (2-4/255 between 9 and 400 spp) and only equal spp gives comparable images.

  scene            -S  size      spp (set -> used)   source of the spp
  cornellBox        6  600x600   400 -> 400          readme render is smoother than the shipped 100 spp; adjacent-pixel
                                                      roughness 2.87/255 identifies ~400 (100 spp: 5.3, 400 spp: 2.8)
  cornellSmoke      7  600x600   10 -> 9             main.go:357 as shipped (roughness 15.4 vs 16.1)
  book3             3  600x600   10 -> 9             main.go:209 as shipped (roughness 22.7 vs 22.1)
  quads             5  400x400   100 -> 100          main.go:239 as shipped; only pixels whose content is deterministic
                                                      (sky, the light, the teal quad) — the Perlin tables are drawn
                                                      from Go's unseeded math/rand and the earth texels from Go's JPEG
                                                      decoder

Bars (0-255 scale): per-channel image mean within MEAN_BAR, mean |difference| of 30x30-pixel block means within
BLOCK_BAR.  JPEG quantisation noise averages out inside a block; what is left is the renderer.

The Go image went through a JPEG encoder after rendering; a noisy 9-spp image does not survive that unchanged
(ringing around isolated bright pixels is clipped at 0, which lifts the mean of the darkest channel: blue reads
0.5-0.9/255 high).  So the comparison is made twice: on the raw render with the bars above, and after passing OUR
render through the same codec settings (the quantisation tables and chroma subsampling read from the reference
JPEG itself), where the channel means agree to <= 0.1/255 and the bars are JPEG_MEAN_BAR / JPEG_BLOCK_BAR.
"""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "golden", "ref_renders")

MEAN_BAR = 1.0      # /255, per channel
BLOCK_BAR = 1.5     # /255, mean over 30x30 blocks and channels
BLOCK = 30
JPEG_MEAN_BAR = 0.5    # /255, after the same JPEG quantisation as the reference image
JPEG_BLOCK_BAR = 1.0

# name -> (scene id, width, SamplesPerPixel to set)
CASES = {"cornellBox": (6, 600, 400), "cornellSmoke": (7, 600, 10), "book3": (3, 600, 10)}


def load_reference(name):
    from PIL import Image
    return np.asarray(Image.open(os.path.join(REF_DIR, name + ".jpg")).convert("RGB")).astype(np.float64)


def through_reference_codec(img8, name):
    """Encode img8 with the quantisation tables and chroma subsampling of the reference JPEG, and decode it again."""
    import io
    from PIL import Image, JpegImagePlugin
    ref = Image.open(os.path.join(REF_DIR, name + ".jpg"))
    buf = io.BytesIO()
    Image.fromarray(np.asarray(img8, dtype=np.uint8)).save(buf, format="JPEG", qtables=ref.quantization,
                                                           subsampling=JpegImagePlugin.get_sampling(ref))
    return np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB")).astype(np.float64)


def print_color(sums, samples):
    """Vec3.PrintColor (color.go:23-46) on per-pixel sums: scale, NaN -> 0, sqrt gamma, clamp to [0, 0.99999], x256."""
    c = np.asarray(sums, dtype=np.float64) / float(samples)
    c = np.where(np.isnan(c), 0.0, c)
    c = np.sqrt(np.maximum(c, 0.0))
    c = np.clip(c, 0.0, 0.99999)
    return np.floor(c * 256.0)


def block_means(img, b=BLOCK):
    h, w, _ = img.shape
    return img[:h // b * b, :w // b * b].reshape(h // b, b, w // b, b, 3).mean(axis=(1, 3))


def roughness(img):
    return float(np.abs(np.diff(img, axis=1)).mean())


def compare(img8, ref8, mask=None):
    """img8, ref8: [H,W,3] arrays on the 0-255 scale.  Returns (per-channel mean difference, block mean |diff|)."""
    assert img8.shape == ref8.shape, (img8.shape, ref8.shape)
    if mask is None:
        dmean = np.abs(img8.mean(axis=(0, 1)) - ref8.mean(axis=(0, 1)))
        dblock = float(np.abs(block_means(img8) - block_means(ref8)).mean())
        return dmean, dblock
    m = mask[..., None].astype(np.float64)
    dmean = np.abs((img8 * m).sum(axis=(0, 1)) - (ref8 * m).sum(axis=(0, 1))) / max(1.0, mask.sum())
    # block means over the masked pixels of blocks that are entirely inside the mask
    full = block_means(np.repeat(m, 3, axis=2))[..., 0] == 1.0
    d = np.abs(block_means(img8) - block_means(ref8))[full]
    return dmean, float(d.mean()) if d.size else 0.0


def assert_matches_reference(img8, name, what, mask=None):
    ref = load_reference(name)
    dmean, dblock = compare(img8, ref, mask)
    assert (dmean <= MEAN_BAR).all(), f"{what} vs readmeImgs/{name}.jpg: channel means differ by {dmean} (/255), bar {MEAN_BAR}"
    assert dblock <= BLOCK_BAR, f"{what} vs readmeImgs/{name}.jpg: 30x30-block mean |diff| {dblock:.3f} (/255), bar {BLOCK_BAR}"
    if mask is None:
        jm, jb = compare(through_reference_codec(img8, name), ref)
        assert (jm <= JPEG_MEAN_BAR).all(), f"{what} vs readmeImgs/{name}.jpg through the same JPEG tables: channel means differ by {jm}, bar {JPEG_MEAN_BAR}"
        assert jb <= JPEG_BLOCK_BAR, f"{what} vs readmeImgs/{name}.jpg through the same JPEG tables: block mean |diff| {jb:.3f}, bar {JPEG_BLOCK_BAR}"
        return dmean, dblock, jm, jb
    return dmean, dblock, None, None
