"""Host-side glTF 2.0 importer and PNG decoder (kiraray_b200/host/gltf.cpp; reference: Assimp import +
createMaterial(..., GLTF2), src/scene/assimp.cpp:93-226, animation src/core/animation.cpp).  The fixture is
written by the test itself (JSON + .bin + a PNG encoded with zlib), so every number has a known answer; the
reference's AnimatedCube asset is loaded too when the checkout is present.  No GPU."""
import json
import os
import struct
import zlib

import numpy as np
import pytest

import kiraray_b200 as krr


def write_png(path, rgba):
    h, w, c = rgba.shape
    raw = bytearray()
    prev = np.zeros((w, c), np.int32)
    for y in range(h):
        row = rgba[y].astype(np.int32)
        ft = y % 3  # filter types 0 (none), 1 (sub), 2 (up) in turn
        if ft == 0:
            enc = row
        elif ft == 1:
            left = np.vstack([np.zeros((1, c), np.int32), row[:-1]])
            enc = row - left
        else:
            enc = row - prev
        raw.append(ft)
        raw += (enc % 256).astype(np.uint8).tobytes()
        prev = row

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, {3: 2, 4: 6}[c], 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(bytes(raw))) + chunk(b"IEND", b""))


def make_fixture(d):
    pos = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (4, 1))
    uv = np.array([[0, 0], [65535, 0], [65535, 65535], [0, 65535]], np.uint16)  # normalised UNSIGNED_SHORT
    idx = np.array([0, 1, 2, 0, 2, 3], np.uint16)
    times = np.array([0.0, 1.0, 3.0], np.float32)
    trans = np.array([[0, 0, 0], [2, 0, 0], [2, 4, 0]], np.float32)
    h = np.sqrt(0.5)
    rots = np.array([[0, 0, 0, 1], [0, 0, h, h], [0, 0, 1, 0]], np.float32)  # 0, 90, 180 degrees about z
    blobs, views, acc = [], [], []
    off = 0

    def add(arr, ctype, typ, normalized=False):
        nonlocal off
        b = arr.tobytes()
        pad = (-len(b)) % 4
        views.append({"buffer": 0, "byteOffset": off, "byteLength": len(b)})
        a = {"bufferView": len(views) - 1, "componentType": ctype, "count": len(arr), "type": typ}
        if normalized:
            a["normalized"] = True
        acc.append(a)
        blobs.append(b + b"\0" * pad)
        off += len(b) + pad
        return len(acc) - 1
    a_pos, a_nrm = add(pos, 5126, "VEC3"), add(nrm, 5126, "VEC3")
    a_uv, a_idx = add(uv, 5123, "VEC2", True), add(idx, 5123, "SCALAR")
    a_t, a_tr, a_rot = add(times, 5126, "SCALAR"), add(trans, 5126, "VEC3"), add(rots, 5126, "VEC4")
    open(os.path.join(d, "quad.bin"), "wb").write(b"".join(blobs))
    tex = np.zeros((3, 2, 4), np.uint8)
    tex[0, 0], tex[0, 1], tex[1, 0], tex[1, 1], tex[2, 0], tex[2, 1] = (255, 0, 0, 255), (0, 255, 0, 255), (0, 0, 255, 128), (188, 188, 188, 255), (10, 20, 30, 255), (1, 2, 3, 4)
    write_png(os.path.join(d, "base.png"), tex)
    doc = {
        "asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}],
        "nodes": [{"name": "root", "translation": [0, 0, 10], "scale": [2, 2, 2], "children": [1, 2]},
                  {"name": "moving", "mesh": 0, "translation": [9, 9, 9]},
                  {"name": "static", "mesh": 0, "matrix": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 5, 6, 7, 1]}],
        "meshes": [{"name": "quad", "primitives": [{"attributes": {"POSITION": a_pos, "NORMAL": a_nrm, "TEXCOORD_0": a_uv}, "indices": a_idx, "material": 0}]}],
        "materials": [{"name": "painted", "pbrMetallicRoughness": {"baseColorFactor": [0.8, 0.6, 0.4, 1.0], "metallicFactor": 0.25, "roughnessFactor": 0.5,
                                                                  "baseColorTexture": {"index": 0}},
                       "emissiveFactor": [1.0, 0.5, 0.25], "extensions": {"KHR_materials_emissive_strength": {"emissiveStrength": 4.0}}}],
        "textures": [{"source": 0}], "images": [{"uri": "base.png"}],
        "buffers": [{"uri": "quad.bin", "byteLength": off}], "bufferViews": views, "accessors": acc,
        "animations": [{"channels": [{"sampler": 0, "target": {"node": 1, "path": "translation"}}, {"sampler": 1, "target": {"node": 1, "path": "rotation"}}],
                        "samplers": [{"input": a_t, "output": a_tr, "interpolation": "LINEAR"}, {"input": a_t, "output": a_rot, "interpolation": "LINEAR"}]}],
    }
    json.dump(doc, open(os.path.join(d, "quad.gltf"), "w"))
    return tex


def app_for(model_path, asset_root):
    cfg = {"resolution": [32, 32], "passes": [{"enable": True, "name": "WavefrontPathTracer", "params": {}}],
           "scene": {"model": [{"model": model_path}]}}
    return krr.HostApp(cfg, asset_root=str(asset_root))


def test_gltf_fixture(tmp_path):
    tex = make_fixture(str(tmp_path))
    app = app_for("quad.gltf", tmp_path)
    d = app.scene_desc().contents
    assert d.n_meshes == 1 and d.n_instances == 2 and d.n_materials == 1
    m = d.meshes[0]
    assert m.n_vertices == 4 and m.n_triangles == 2 and m.material == 0
    assert np.allclose(np.ctypeslib.as_array(m.positions, (12,)), [0, 0, 0, 1, 0, 0, 1, 1, 0, 0, 1, 0])
    assert list(np.ctypeslib.as_array(m.indices, (6,))) == [0, 1, 2, 0, 2, 3]
    assert np.allclose(np.ctypeslib.as_array(m.texcoords, (8,)), [0, 0, 1, 0, 1, 1, 0, 1]), "normalised UNSIGNED_SHORT texcoords"
    # material mapping of createMaterial(GLTF2): base colour -> diffuse, roughness -> specular.g, metallic -> specular.b
    mat = d.materials[0]
    assert np.allclose(list(mat.diffuse), [0.8, 0.6, 0.4, 1.0]) and mat.specular[1] == pytest.approx(0.5) and mat.specular[2] == pytest.approx(0.25)
    assert mat.shading_model == 0 and mat.bsdf_type == 4
    em = mat.textures[2]
    assert em.valid == 1 and np.allclose(list(em.value)[:3], [4.0, 2.0, 1.0]), "emissiveFactor * emissiveStrength as a constant texture"
    dt = mat.textures[0]
    assert dt.valid == 1 and (dt.width, dt.height) == (2, 3)
    texels = np.ctypeslib.as_array(dt.image, (3 * 2 * 4,)).reshape(3, 2, 4)
    lin = lambda c: np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)
    want = tex.astype(np.float64) / 255
    want[..., :3] = lin(want[..., :3])
    assert np.allclose(texels, want, atol=1e-6), "PNG filters 0/1/2 decoded, sRGB -> linear on colour, alpha untouched"
    # static child: root (T(0,0,10) * S(2)) * matrix(translate 5,6,7)
    xs = np.array(list(d.instances[1].transform)).reshape(3, 4)
    assert np.allclose(xs, [[2, 0, 0, 10], [0, 2, 0, 12], [0, 0, 2, 24]])
    # animated child at t = 0.5: translation (1,0,0), rotation 45 degrees about z (nlerp of the keys), under the root
    app.camera(0.5)
    d = app.scene_desc().contents
    xm = np.array(list(d.instances[0].transform)).reshape(3, 4)
    c = np.sqrt(0.5)
    assert np.allclose(xm, [[2 * c, -2 * c, 0, 2], [2 * c, 2 * c, 0, 0], [0, 0, 2, 10]], atol=1e-6)
    app.camera(2.0)  # second segment: translation (2,2,0), rotation 135 degrees
    xm = np.array(list(app.scene_desc().contents.instances[0].transform)).reshape(3, 4)
    assert np.allclose(xm, [[-2 * c, -2 * c, 0, 4], [2 * c, -2 * c, 0, 4], [0, 0, 2, 10]], atol=1e-6)


def test_animated_ancestors(tmp_path):
    """An animated root above a static node above an animated mesh node and a static mesh node: every instance
    follows the chain  root(t) * static * [node(t)]  (the reference animates scene-graph nodes, animation.cpp)."""
    make_fixture(str(tmp_path))
    doc = json.load(open(tmp_path / "quad.gltf"))
    a_t, a_tr, a_rot = 4, 5, 6  # accessors of make_fixture: times [0, 1, 3], translations, rotations 0 / 90 / 180 degrees about z
    doc["nodes"] = [{"name": "R", "children": [1]},
                    {"name": "M", "translation": [0, 1, 0], "scale": [2, 2, 2], "children": [2, 3]},
                    {"name": "C", "mesh": 0, "translation": [1, 0, 0]},
                    {"name": "S", "mesh": 0, "translation": [0, 0, 3]}]
    doc["animations"] = [{"channels": [{"sampler": 0, "target": {"node": 0, "path": "translation"}}, {"sampler": 1, "target": {"node": 2, "path": "rotation"}}],
                          "samplers": [{"input": a_t, "output": a_tr, "interpolation": "LINEAR"}, {"input": a_t, "output": a_rot, "interpolation": "LINEAR"}]}]
    json.dump(doc, open(tmp_path / "chain.gltf", "w"))
    app = app_for("chain.gltf", tmp_path)
    c = np.sqrt(0.5)

    def xf(i):
        return np.array(list(app.scene_desc().contents.instances[i].transform)).reshape(3, 4)
    app.camera(0.5)
    assert np.allclose(xf(0), [[2 * c, -2 * c, 0, 3], [2 * c, 2 * c, 0, 1], [0, 0, 2, 0]], atol=1e-6)
    assert np.allclose(xf(1), [[2, 0, 0, 1], [0, 2, 0, 1], [0, 0, 2, 6]], atol=1e-6)
    app.camera(2.0)
    assert np.allclose(xf(0), [[-2 * c, -2 * c, 0, 4], [2 * c, -2 * c, 0, 3], [0, 0, 2, 0]], atol=1e-6)
    assert np.allclose(xf(1), [[2, 0, 0, 2], [0, 2, 0, 3], [0, 0, 2, 6]], atol=1e-6)
    app.camera(10.0)  # clamped at the last key: root at (2, 4, 0), 180 degrees
    assert np.allclose(xf(0), [[-2, 0, 0, 4], [0, -2, 0, 5], [0, 0, 2, 0]], atol=1e-6)


def to_glb(d):
    """Packs quad.gltf + quad.bin + base.png into quad.glb: JSON chunk + one BIN chunk, the image in a buffer view."""
    doc = json.load(open(os.path.join(d, "quad.gltf")))
    blob = open(os.path.join(d, "quad.bin"), "rb").read()
    png = open(os.path.join(d, "base.png"), "rb").read()
    doc["bufferViews"].append({"buffer": 0, "byteOffset": len(blob), "byteLength": len(png)})
    doc["images"] = [{"bufferView": len(doc["bufferViews"]) - 1, "mimeType": "image/png"}]
    blob += png + b"\0" * ((-len(png)) % 4)
    doc["buffers"] = [{"byteLength": len(blob)}]
    text = json.dumps(doc).encode()
    text += b" " * ((-len(text)) % 4)
    body = struct.pack("<II", len(text), 0x4E4F534A) + text + struct.pack("<II", len(blob), 0x004E4942) + blob
    open(os.path.join(d, "quad.glb"), "wb").write(struct.pack("<III", 0x46546C67, 2, 12 + len(body)) + body)


def test_binary_gltf_gives_the_same_scene(tmp_path):
    make_fixture(str(tmp_path))
    to_glb(str(tmp_path))
    os.remove(tmp_path / "quad.bin"), os.remove(tmp_path / "base.png")  # the .glb is self-contained
    a = app_for("quad.glb", tmp_path)
    make_fixture(str(tmp_path))
    b = app_for("quad.gltf", tmp_path)
    da, db = a.scene_desc().contents, b.scene_desc().contents
    assert (da.n_meshes, da.n_instances, da.n_materials) == (db.n_meshes, db.n_instances, db.n_materials) == (1, 2, 1)
    for name, n in (("positions", 12), ("normals", 12), ("texcoords", 8)):
        assert np.array_equal(np.ctypeslib.as_array(getattr(da.meshes[0], name), (n,)), np.ctypeslib.as_array(getattr(db.meshes[0], name), (n,)))
    assert list(np.ctypeslib.as_array(da.meshes[0].indices, (6,))) == [0, 1, 2, 0, 2, 3]
    ta, tb = da.materials[0].textures[0], db.materials[0].textures[0]
    assert ta.valid == 1 and (ta.width, ta.height) == (2, 3)
    assert np.array_equal(np.ctypeslib.as_array(ta.image, (24,)), np.ctypeslib.as_array(tb.image, (24,)))
    for t in (0.5, 2.0):  # the animation samplers read the BIN chunk too
        a.camera(t), b.camera(t)
        assert list(a.scene_desc().contents.instances[0].transform) == list(b.scene_desc().contents.instances[0].transform)
    bad = bytearray(open(tmp_path / "quad.glb", "rb").read())
    bad[4] = 1  # version
    open(tmp_path / "old.glb", "wb").write(bytes(bad))
    with pytest.raises(RuntimeError, match="glb version"):
        app_for("old.glb", tmp_path)


@pytest.mark.skipif(not os.path.exists("/root/reference/common/assets/scenes/anime-cube/AnimatedCube.gltf"), reason="reference checkout not present")
def test_reference_animated_cube():
    app = app_for("common/assets/scenes/anime-cube/AnimatedCube.gltf", "/root/reference")
    d = app.scene_desc().contents
    assert d.n_meshes == 1 and d.n_instances == 1 and d.meshes[0].n_triangles == 12 and d.meshes[0].n_vertices == 36
    assert d.meshes[0].tangents and d.meshes[0].texcoords and d.meshes[0].normals
    mat = d.materials[0]
    assert mat.textures[0].valid and mat.textures[0].width == mat.textures[0].height > 16, "base colour PNG decoded"
    assert mat.textures[1].valid, "metallic-roughness PNG decoded"
    try:  # the 892 KB base-colour PNG (all five filter types) against OpenCV's decoder
        import cv2
        ref = cv2.imread("/root/reference/common/assets/scenes/anime-cube/AnimatedCube_BaseColor.png", cv2.IMREAD_UNCHANGED)
    except ImportError:
        ref = None
    if ref is not None:
        t = mat.textures[0]
        ours = np.ctypeslib.as_array(t.image, (t.height * t.width * 4,)).reshape(t.height, t.width, 4)
        rgb = ref[..., [2, 1, 0]].astype(np.float64) / 255
        lin = np.where(rgb <= 0.04045, rgb / 12.92, ((rgb + 0.055) / 1.055) ** 2.4)
        assert np.allclose(ours[..., :3], lin, atol=1e-6)
    # rotation keys: (0,-1,0,0) -> ... about y; the node transform changes with time and stays a rotation
    x0 = np.array(list(d.instances[0].transform)).reshape(3, 4).copy()
    app.camera(1.0)
    x1 = np.array(list(app.scene_desc().contents.instances[0].transform)).reshape(3, 4)
    assert not np.allclose(x0, x1) and np.allclose(x1[:, :3] @ x1[:, :3].T, np.eye(3), atol=1e-5)
