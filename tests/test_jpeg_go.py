"""The Go-exact JPEG texel import (csrc/jpeg_go.hpp) against the reference's own golden vector:
internal/imageloader/imageLoader_test.go:33-62 lists the 25 RGB texels Go's image/jpeg + color.YCbCr produce for
test.jpg (5x5, 4:2:0).  tests/golden/imageloader/test.jpg is that file."""
import os
import numpy as np
import pytest
import go_raytracer_b200 as g

HERE = os.path.dirname(os.path.abspath(__file__))

# imageLoader_test.go:33-62 (LOSSY_IMG_DATA), row-major, idx = y * Width + x
LOSSY_IMG_DATA = [(216, 226, 255), (160, 170, 208), (133, 143, 180), (132, 142, 179), (226, 237, 255),
                  (114, 123, 162), (60, 69, 108), (85, 95, 131), (47, 57, 93), (136, 147, 177),
                  (90, 99, 138), (0, 9, 48), (26, 37, 71), (76, 86, 121), (99, 110, 140),
                  (130, 140, 178), (3, 13, 52), (9, 20, 54), (124, 134, 169), (147, 158, 188),
                  (218, 228, 255), (108, 117, 156), (102, 112, 147), (134, 144, 179), (222, 234, 255)]


def test_decoder_reproduces_gos_texels_for_the_reference_fixture():
    img = g.load_image(os.path.join(HERE, "golden", "imageloader", "test.jpg"))
    assert img.shape == (5, 5, 3)                                   # imageLoader_test.go:23 "5, 5, jpeg, 25"
    assert np.array_equal(img.reshape(25, 3), np.array(LOSSY_IMG_DATA, dtype=np.uint8))


def test_libjpeg_does_not(_=None):
    """Why the decoder exists: Pillow/libjpeg (fancy chroma upsampling, a different IDCT and colour matrix) is off by up
    to 3/255 on the same file — the +-3 LSB the round-1 earth texels carried."""
    Image = pytest.importorskip("PIL.Image")
    pil = np.asarray(Image.open(os.path.join(HERE, "golden", "imageloader", "test.jpg")).convert("RGB"))
    d = np.abs(pil.astype(int).reshape(25, 3) - np.array(LOSSY_IMG_DATA))
    assert 0 < d.max() <= 3


def test_ycbcr_known_answer():
    """image/color/ycbcr.go's own worked example: YCbCr{0x7f, 0x7f, 0x7f}.RGBA() = 0x7e18 0x808d 0x7db9, i.e. 7e 80 7d
    after imageLoader.go:66-68's >> 8.  A grey 8x8 JPEG with those planes is not at hand, so the example is checked
    through a synthetic single-MCU file built here: DC-only blocks, quantiser 1."""
    import struct
    # DC-only baseline JPEG, 8x8, 3 components 1x1, all three planes = 0x7f: DC coefficient = (0x7f - 128) * 8 = -8
    def seg(m, payload):
        return bytes([0xFF, m]) + struct.pack(">H", len(payload) + 2) + payload
    dqt = seg(0xDB, bytes([0]) + bytes([1] * 64))
    sof = seg(0xC0, bytes([8, 0, 8, 0, 8, 3, 1, 0x11, 0, 2, 0x11, 0, 3, 0x11, 0]))
    # one DC table: category 4 has the 1-bit code '0'; one AC table: EOB (0x00) has the 1-bit code '0'
    dht_dc = seg(0xC4, bytes([0x00, 1] + [0] * 15 + [4]))
    dht_ac = seg(0xC4, bytes([0x10, 1] + [0] * 15 + [0]))
    sos = seg(0xDA, bytes([3, 1, 0x00, 2, 0x00, 3, 0x00, 0, 63, 0]))
    # component 1: DC diff -8 -> category 4, bits = -8 + 15 = 7 = 0111; EOB '0'.  components 2, 3: the predictor is per
    # component, so the same again: '0' '0111' '0' three times = 18 bits, padded with ones
    bits = "001110" * 3
    bits += "1" * (-len(bits) % 8)
    data = bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8))
    jpg = b"\xFF\xD8" + dqt + sof + dht_dc + dht_ac + sos + data + b"\xFF\xD9"
    img = g.decode_jpeg(jpg)
    assert img.shape == (8, 8, 3)
    assert (img == np.array([0x7e, 0x80, 0x7d], dtype=np.uint8)).all()


def test_earth_fixture_is_the_go_decode():
    """tests/golden/earthmap_rgb8.npz (the C4 / quads texture) was produced by this decoder from the reference's
    earthmap.jpg (tests/golden/make_earthmap_fixture.py): 1024 x 512, 4:4:4 baseline."""
    rgb = np.load(os.path.join(HERE, "golden", "earthmap_rgb8.npz"))["rgb"]
    assert rgb.shape == (512, 1024, 3) and rgb.dtype == np.uint8
    ref = "/root/reference/earthmap.jpg"
    if os.path.exists(ref):                     # only in the build container
        assert np.array_equal(g.load_image(ref), rgb)


def test_refuses_what_it_does_not_restate():
    with pytest.raises(g.GrtError):
        g.decode_jpeg(b"\x89PNG\r\n\x1a\n" + bytes(32))
    with pytest.raises(g.GrtError):
        g.decode_jpeg(b"\xFF\xD8\xFF\xC2\x00\x0b\x08\x00\x08\x00\x08\x01\x01\x11\x00\xFF\xD9")    # progressive SOF2
