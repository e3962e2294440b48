"""Host-side HDR image files (kiraray_b200/host/image.cpp; reference src/core/texture.cpp:27-118,
src/util/image.cpp): EXR and PFM round trips, the file layout against an independent decoder written
from the OpenEXR specification, OpenCV's reader/writer as a third party, and the channel permutation of
the reference's EXR writer that ErrorMeasurePass undoes.  No GPU."""
import os
import struct

import numpy as np
import pytest

import kiraray_b200 as krr


def sample_image(w=37, h=23, seed=7272):
    rng = np.random.Generator(np.random.PCG64(seed))
    img = rng.uniform(0, 4, (h, w, 4)).astype(np.float32)
    img[0, 0] = (0, 1e-6, 65000.0, 1)  # small / large values of the half range
    img[1, 2] = (0.5, 0.25, 0.125, 1)
    return img


def decode_exr_uncompressed(path):
    """Minimal scanline / NO_COMPRESSION decoder following the OpenEXR file layout document."""
    b = open(path, "rb").read()
    assert b[:4] == bytes([0x76, 0x2F, 0x31, 0x01]) and b[4] == 2
    pos, attrs = 8, {}
    while b[pos] != 0:
        e = b.index(0, pos); name = b[pos:e].decode(); pos = e + 1
        e = b.index(0, pos); typ = b[pos:e].decode(); pos = e + 1
        size = struct.unpack_from("<i", b, pos)[0]; pos += 4
        attrs[name] = (typ, b[pos:pos + size]); pos += size
    pos += 1
    chans, c = [], attrs["channels"][1]
    q = 0
    while c[q] != 0:
        e = c.index(0, q); nm = c[q:e].decode(); q = e + 1
        ptype, _, xs, ys = struct.unpack_from("<iiii", c, q); q += 16
        chans.append((nm, ptype))
    assert attrs["compression"][1][0] == 0
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    offs = struct.unpack_from("<%dQ" % h, b, pos)
    out = {}
    for nm, _ in chans:
        out[nm] = np.zeros((h, w), np.float32)
    for o in offs:
        y, size = struct.unpack_from("<ii", b, o)
        p = o + 8
        for nm, ptype in chans:
            dt = np.float16 if ptype == 1 else np.float32
            out[nm][y - y0] = np.frombuffer(b, dt, w, p).astype(np.float32)
            p += w * np.dtype(dt).itemsize
    return [nm for nm, _ in chans], out


def test_exr_float_roundtrip_is_bit_exact(tmp_path):
    img = sample_image()
    for zip_ in (False, True):
        p = tmp_path / f"f{int(zip_)}.exr"
        krr.save_exr(p, img, half=False, zip=zip_)
        back = krr.load_image(p)
        assert np.array_equal(back.view(np.uint32), img.view(np.uint32))
    assert os.path.getsize(tmp_path / "f1.exr") < os.path.getsize(tmp_path / "f0.exr")


def test_exr_half_roundtrip_matches_numpy_float16(tmp_path):
    img = sample_image()
    for zip_ in (False, True):
        p = tmp_path / f"h{int(zip_)}.exr"
        krr.save_exr(p, img, half=True, zip=zip_)
        back = krr.load_image(p)
        assert np.array_equal(back, img.astype(np.float16).astype(np.float32))  # round to nearest even, like numpy


def test_exr_file_layout_against_the_specification(tmp_path):
    img = sample_image()
    p = tmp_path / "spec.exr"
    krr.save_exr(p, img, half=True, zip=False)
    names, planes = decode_exr_uncompressed(p)
    assert names == ["A", "B", "G", "R"], "channel lists are sorted by name"
    for k, nm in enumerate("RGBA"):
        assert np.array_equal(planes[nm], img[..., k].astype(np.float16).astype(np.float32))


def test_flip_and_reference_channel_order(tmp_path):
    """saveImage(path, flip=True) as AccumulatePass::saveImage calls it, read back the way
    ErrorMeasurePass::loadReferenceImage does (flip=True, then the {3,0,1,2} permutation)."""
    img = sample_image()
    p = tmp_path / "ref.exr"
    krr.save_image(p, img, flip=True, reference_channel_order=True)
    names, planes = decode_exr_uncompressed(p)
    half = lambda a: a.astype(np.float16).astype(np.float32)
    # file plane X of the reference's writer: A <- image R, R <- image G, G <- image B, B <- image A; rows flipped
    assert np.array_equal(planes["A"], half(img[::-1, :, 0])) and np.array_equal(planes["R"], half(img[::-1, :, 1]))
    assert np.array_equal(planes["G"], half(img[::-1, :, 2])) and np.array_equal(planes["B"], half(img[::-1, :, 3]))
    back = krr.load_image(p, flip=True)
    undone = back[..., [3, 0, 1, 2]]
    assert np.array_equal(undone, half(img))
    # and without the quirk the file is a plain RGBA image
    krr.save_image(tmp_path / "plain.exr", img, flip=False, reference_channel_order=False)
    assert np.array_equal(krr.load_image(tmp_path / "plain.exr"), half(img))


def test_pfm_roundtrip_and_big_endian(tmp_path):
    img = sample_image()
    p = tmp_path / "a.pfm"
    krr.save_image(p, img)
    back = krr.load_image(p)
    assert np.array_equal(back[..., :3].view(np.uint32), img[..., :3].view(np.uint32)) and np.all(back[..., 3] == 1)
    # hand-written big-endian grey PFM, scale 2: bottom row first
    g = np.arange(6, dtype=np.float32).reshape(2, 3)
    with open(tmp_path / "g.pfm", "wb") as f:
        f.write(b"Pf\n3 2\n2.0\n")
        f.write(g[::-1].astype(">f4").tobytes())
    back = krr.load_image(tmp_path / "g.pfm")
    assert np.array_equal(back[..., 0], 2 * g) and np.array_equal(back[..., 3], 2 * g)


def test_errors_are_reported(tmp_path):
    with pytest.raises(RuntimeError):
        krr.load_image(tmp_path / "missing.exr")
    (tmp_path / "bad.exr").write_bytes(b"not an exr file at all")
    with pytest.raises(RuntimeError, match="not an OpenEXR"):
        krr.load_image(tmp_path / "bad.exr")
    with pytest.raises(RuntimeError, match="unsupported image format"):
        krr.save_image(tmp_path / "x.png", sample_image())


def test_opencv_interoperability(tmp_path):
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    img = sample_image()
    p = tmp_path / "ours.exr"
    krr.save_exr(p, img, half=False, zip=True)
    theirs = cv2.imread(str(p), cv2.IMREAD_UNCHANGED)
    if theirs is None:
        pytest.skip("this OpenCV build has no OpenEXR codec")
    assert np.array_equal(theirs[..., [2, 1, 0, 3]], img)  # OpenCV returns BGRA
    q = tmp_path / "theirs.exr"
    assert cv2.imwrite(str(q), img[..., [2, 1, 0, 3]], [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_FLOAT])
    assert np.array_equal(krr.load_image(q), img)


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["piz_half_37x45", "piz_float_70x33", "piz_flat_40x40"])
def test_piz_fixtures_decode_bit_exactly(name):
    """PIZ (wavelet + Huffman) files written by OpenCV's OpenEXR codec (tools/make_golden_piz.py): 14-bit and
    16-bit wavelet butterflies, odd sizes, a short last block, the Huffman run-length symbol."""
    mine = krr.load_image(os.path.join(GOLDEN, name + ".exr"))
    want = np.load(os.path.join(GOLDEN, name + ".npy"))
    assert mine.shape == want.shape
    assert np.array_equal(mine.view(np.uint32), want.view(np.uint32))


def test_corrupt_piz_data_is_an_error(tmp_path):
    data = bytearray(open(os.path.join(GOLDEN, "piz_half_37x45.exr"), "rb").read())
    for k in range(len(data) - 600, len(data) - 200):  # inside the Huffman stream of the last chunk
        data[k] ^= 0x5A
    (tmp_path / "bad_piz.exr").write_bytes(bytes(data))
    try:
        img = krr.load_image(tmp_path / "bad_piz.exr")  # a damaged stream may still decode to a full block
        assert img.shape == (45, 37, 4)
    except RuntimeError as e:
        assert "PIZ" in str(e)
    (tmp_path / "short_piz.exr").write_bytes(bytes(data[: len(data) // 2]))
    with pytest.raises(RuntimeError):
        krr.load_image(tmp_path / "short_piz.exr")


@pytest.mark.skipif(not os.path.exists("/root/reference/common/assets/textures/sky.exr"), reason="reference checkout not present")
def test_the_reference_sky_texture_decodes_like_opencv():
    """The one EXR the reference ships (the environment map of its example configs) is PIZ-compressed FLOAT."""
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    path = "/root/reference/common/assets/textures/sky.exr"
    mine = krr.load_image(path)
    theirs = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if theirs is None:
        pytest.skip("this OpenCV build has no OpenEXR codec")
    assert mine.shape == (512, 1024, 4)
    assert np.array_equal(mine.view(np.uint32), theirs[..., [2, 1, 0, 3]].view(np.uint32))
