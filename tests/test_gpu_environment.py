"""GPU parity of image-based infinite lights (InfiniteLight::Li / sampleLi with a lat-long image,
src/core/light.h:222-246): the open Cornell box lit by an environment map, loaded through the scene JSON's
"environment" key.  Escaping rays look the image up in handleMiss, NEE samples the sphere uniformly."""
import json
import os

import numpy as np
import pytest

import kiraray_b200 as krr
import oracle_binding as ob
from __graft_entry__ import relmse

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CBOX = os.path.join(ROOT, "assets", "configs", "cbox.json")
KIND = "reference"


def sky(path, w=64, h=32):
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.empty((h, w, 4), np.float32)
    img[..., 0] = 0.2 + 2.5 * np.exp(-((x - 40) ** 2 + (y - 6) ** 2) / 30.0)  # a soft sun
    img[..., 1] = 0.3 + 0.5 * (1 - y / h)
    img[..., 2] = 0.4 + 1.0 * (1 - y / h) + 0.2 * np.sin(x / w * 2 * np.pi)
    img[..., 3] = 1
    krr.save_exr(path, img, half=False, zip=True)
    return img


def render(cfg, w, h, spp, depth):
    app = krr.HostApp(cfg, asset_root=ROOT)
    app.set_resolution(w, h)
    app.set_wfpt_params(spp=spp, max_depth=depth)
    gpu = krr.Wfpt(params=dict(app.wfpt_params()))
    gpu.set_scene(app.scene_desc())
    gpu.resize(w, h)
    gpu.begin_frame(1, app.camera())
    film = gpu.render_to_host().copy()
    orc = ob.Oracle(app.scene_desc(), KIND)
    ref = orc.render(app.camera(), w, h, frame_index=1, spp=spp, max_depth=depth, use_bvh=True)
    orc.close()
    return gpu, film, ref


def test_environment_map_matches_the_oracle(tmp_path):
    sky(tmp_path / "sky.exr")
    cfg = json.load(open(CBOX))
    cfg["scene"]["environment"] = str(tmp_path / "sky.exr")
    # pull the camera back so that part of the frame sees the sky directly
    cfg["scene"]["cameraController"]["mData"]["radius"] = 6.5
    w = h = 80
    gpu, film, ref = render(cfg, w, h, spp=4, depth=4)
    inst, prim = gpu.first_hits()
    assert np.array_equal(inst, ref["first_hits"][:, 0]) and np.array_equal(prim, ref["first_hits"][:, 1])
    assert (inst < 0).sum() > 0.1 * w * h  # sky pixels
    st, rs = gpu.stats(), ref["stats"]
    for k in ("closest_rays", "shadow_rays", "miss_items", "scatter_items"):
        assert abs(st[k] - rs[k]) <= 0.002 * rs[k] + 4, (k, st[k], rs[k])
    assert np.isfinite(film).all()
    err = relmse(film, ref["film"])
    assert err <= 2e-3, err
    # the image is what lights the scene: the constant-tint light gives a different film
    cfg2 = json.load(open(CBOX))
    cfg2["scene"]["model"].append({"type": "light", "name": "env", "params": {"type": "infinite"}})
    cfg2["scene"]["cameraController"]["mData"]["radius"] = 6.5
    _, film2, ref2 = render(cfg2, w, h, spp=4, depth=4)
    assert relmse(film2, ref2["film"]) <= 2e-3
    assert relmse(film, film2) > 0.05


def test_rotated_environment_light(tmp_path):
    sky(tmp_path / "sky.exr")
    cfg = json.load(open(CBOX))
    cfg["scene"]["cameraController"]["mData"]["radius"] = 6.5
    cfg["scene"]["model"].append({"type": "light", "name": "environment", "rotate": [0.8, 0.2, 0.5, 0.2645751],
                                  "params": {"type": "infinite", "texture": str(tmp_path / "sky.exr"), "scale": 1.7}})
    gpu, film, ref = render(cfg, 64, 64, spp=4, depth=3)
    assert relmse(film, ref["film"]) <= 2e-3
