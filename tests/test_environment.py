"""Image-based infinite lights: the `"texture"` parameter of a JSON light and the `"environment"` key of a
scene / application config (reference src/scene/krrscene.cpp:44-48, 267-274; src/main/renderer.cpp:295-302;
InfiniteLight::Li, src/core/light.h:242-246), the `$name` path rule of Image::loadImage (texture.cpp:34-38),
and the oracle's lat-long lookup.  The GPU comparison is in test_gpu_environment.py."""
import copy
import json
import os

import numpy as np
import pytest

import kiraray_b200 as krr
import oracle_binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CBOX = os.path.join(ROOT, "assets", "configs", "cbox.json")
ENV = "tests/golden/piz_float_70x33.exr"  # 70 x 33 FLOAT, PIZ-compressed (tools/make_golden_piz.py)
KIND = "reference"
KRR_LIGHT_INFINITE = 4  # include/krr_wfpt.h


def config(**scene_extra):
    cfg = json.load(open(CBOX))
    cfg["scene"].update(copy.deepcopy(scene_extra))
    return cfg


def lights_of(app):
    d = app.scene_desc().contents
    return [d.lights[i] for i in range(d.n_lights)]


def test_environment_key_adds_an_infinite_light_with_the_image():
    app = krr.HostApp(config(environment=ENV), asset_root=ROOT)
    (l,) = lights_of(app)
    assert l.type == KRR_LIGHT_INFINITE and l.scale == 1.0 and list(l.color) == [1.0, 1.0, 1.0]
    assert l.texture.valid == 1 and (l.texture.width, l.texture.height) == (70, 33)
    want = np.load(os.path.join(ROOT, "tests", "golden", "piz_float_70x33.npy"))
    got = np.ctypeslib.as_array(l.texture.image, shape=(33, 70, 4))
    assert np.array_equal(got, want)  # row 0 first (Texture::createFromFile: flip = false)
    assert l.scene_radius > 0


def test_application_level_environment_and_light_texture_parameter():
    cfg = json.load(open(CBOX))
    cfg["environment"] = ENV  # renderer.cpp:295-302 requires a model before it
    cfg["model"] = "assets/cbox/cbox.obj"
    scene = cfg.pop("scene")
    app = krr.HostApp(cfg, asset_root=ROOT)
    assert [l.texture.width for l in lights_of(app)] == [70]
    cfg = {"passes": json.load(open(CBOX))["passes"], "scene": scene}
    cfg["scene"]["model"].append({"type": "light", "name": "environment", "params": {"type": "infinite", "texture": ENV, "scale": 2.5}})
    app = krr.HostApp(cfg, asset_root=ROOT)
    (l,) = lights_of(app)
    assert l.scale == 2.5 and l.texture.valid == 1 and l.texture.height == 33


def test_dollar_names_resolve_in_the_texture_directory(tmp_path, monkeypatch):
    monkeypatch.setenv("KRR_TEXTURE_DIR", os.path.join(ROOT, "tests", "golden"))
    app = krr.HostApp(config(environment="$piz_half_37x45.exr"), asset_root=ROOT)
    assert lights_of(app)[0].texture.width == 37


def test_missing_texture_leaves_a_constant_light(capfd):
    app = krr.HostApp(config(environment="no/such/file.exr"), asset_root=ROOT)
    (l,) = lights_of(app)
    assert l.type == KRR_LIGHT_INFINITE and not l.texture.image  # the reference logs and carries on
    assert "failed to load texture" in capfd.readouterr().err


def test_oracle_constant_image_equals_constant_tint(tmp_path):
    """An image whose texels are all (0.5, 0.25, 1) must give the film of the constant texture of that value,
    bit for bit: it exercises the lat-long lookup in handleMiss and in sampleLi without depending on it."""
    img = np.empty((8, 16, 4), np.float32)
    img[...] = (0.5, 0.25, 1.0, 1.0)
    krr.save_exr(tmp_path / "flat.exr", img, half=False, zip=True)
    films = []
    for tex in (str(tmp_path / "flat.exr"), None):
        app = krr.HostApp(config(environment=str(tmp_path / "flat.exr")), asset_root=ROOT)
        app.set_resolution(48, 48)
        if tex is None:  # same light, image dropped, constant value kept
            l = app.scene_desc().contents.lights[0]
            l.texture.value[0], l.texture.value[1], l.texture.value[2] = 0.5, 0.25, 1.0
            l.texture.image = None
        orc = ob.Oracle(app.scene_desc(), KIND)
        films.append(orc.render(app.camera(), 48, 48, frame_index=1, spp=2, max_depth=3, use_bvh=True)["film"])
        orc.close()
    assert np.array_equal(films[0].view(np.uint32), films[1].view(np.uint32))
    assert films[0][..., :3].max() > 0
