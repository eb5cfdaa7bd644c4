"""Output containers: P3 text as camera.go:160 + color.go:45 write it, and the same pixels as
binary P6 and PNG.  CPU only: pure host code."""
import struct
import zlib

import numpy as np

import go_raytracer_b200 as g


def _img(h, w, seed=0):
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, 3), dtype=np.uint8)


def test_p3_text_is_the_reference_format():
    img = _img(3, 5)
    txt = g.write_ppm(img).decode()
    lines = txt.split("\n")
    assert lines[0] == "P3" and lines[1] == "5 3" and lines[2] == "255" and lines[-1] == ""
    body = np.array([[int(x) for x in l.split(" ")] for l in lines[3:-1]], dtype=np.uint8)
    assert np.array_equal(body.reshape(3, 5, 3), img)


def test_p6_round_trip():
    img = _img(7, 9, 1)
    raw = g.write_p6(img)
    head = b"P6\n9 7\n255\n"
    assert raw.startswith(head) and len(raw) == len(head) + img.size
    assert np.array_equal(np.frombuffer(raw[len(head):], dtype=np.uint8).reshape(7, 9, 3), img)


def _decode_png(raw):
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(raw):
        (n,) = struct.unpack(">I", raw[pos:pos + 4])
        typ, data = raw[pos + 4:pos + 8], raw[pos + 8:pos + 8 + n]
        (crc,) = struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(typ + data) == crc, typ
        chunks.append((typ, data))
        pos += 12 + n
    assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]
    w, h, depth, ctype, comp, filt, inter = struct.unpack(">IIBBBBB", chunks[0][1])
    assert (depth, ctype, comp, filt, inter) == (8, 2, 0, 0, 0)
    rows = np.frombuffer(zlib.decompress(chunks[1][1]), dtype=np.uint8).reshape(h, 3 * w + 1)   # checks Adler-32 too
    assert (rows[:, 0] == 0).all()
    return rows[:, 1:].reshape(h, w, 3)


def test_png_decodes_with_zlib_small_and_multi_block():
    for h, w, seed in [(1, 1, 2), (5, 8, 3), (200, 300, 4)]:          # the last one spans several 64 KiB stored blocks
        img = _img(h, w, seed)
        assert np.array_equal(_decode_png(g.write_png(img)), img)
