"""Host layer (C++17, libkrr_host.so): the reference's JSON config / scene schema / pass factory
surface, exercised without a GPU (loading a config touches no device)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import kiraray_b200 as krr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CBOX = os.path.join(ROOT, "assets", "configs", "cbox.json")


def test_cbox_config_loads_with_reference_defaults():
    app = krr.HostApp(CBOX, asset_root=ROOT)
    assert app.resolution == (750, 750)  # common/configs/example_cbox.json
    p = app.wfpt_params()
    # integrator.h:96-103 keys; values from the config, the rest the reference defaults
    assert p["nee"] is True and p["max_depth"] == 10 and abs(p["rr"] - 0.8) < 1e-7
    assert p["enable_medium"] is True and p["enable_clamp"] is False and p["clamp_max"] == 1000.0
    assert p["spp"] == 1


def test_cbox_scene_matches_the_asset():
    app = krr.HostApp(CBOX, asset_root=ROOT)
    d = app.scene_desc().contents
    tris = sum(d.meshes[i].n_triangles for i in range(d.n_meshes))
    assert tris == 36 and d.n_instances == d.n_meshes == 8 and d.n_materials == 8  # 36 triangles, 8 materials
    emissive = [i for i in range(d.n_materials) if d.materials[i].textures[2].valid]
    assert len(emissive) == 1
    assert list(d.materials[emissive[0]].textures[2].value)[:3] == [17.0, 12.0, 4.0]  # Ke 17 12 4
    for i in range(d.n_materials):
        assert d.materials[i].bsdf_type == 4      # Material default mBsdfType = Disney (core/texture.h:165)
        assert d.materials[i].shading_model == 1  # OBJ -> SpecularGlossiness (scene/assimp.cpp:222-224)
    for i in range(d.n_meshes):
        m = d.meshes[i]
        idx = np.ctypeslib.as_array(m.indices, (m.n_triangles * 3,))
        assert idx.min() >= 0 and idx.max() < m.n_vertices
        assert bool(m.normals)


def test_camera_follows_orbit_controller_and_aspect():
    app = krr.HostApp(CBOX, asset_root=ROOT)
    app.set_resolution(1920, 1080)
    cam = app.camera()
    assert abs(cam.aspect_ratio - 1920 / 1080) < 1e-6
    assert abs(cam.focal_length - 21.0) < 1e-6 and cam.lens_radius == 0.0
    # film height 24 mm, width = aspect * height (Camera::update, core/camera.cpp:6-17)
    assert abs(cam.film_size[1] - 24.0) < 1e-5 and abs(cam.film_size[0] - 24.0 * 1920 / 1080) < 1e-4
    t = np.array(list(cam.transform)).reshape(3, 4)
    R = t[:, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-5)
    target = np.array([-0.011140584014356136, 1.0211291313171387, -0.2837386131286621])
    assert abs(np.linalg.norm(t[:, 3] - target) - 3.1123313903808594) < 1e-4
    assert cam.medium == -1


def test_params_round_trip_and_validation():
    app = krr.HostApp(CBOX, asset_root=ROOT)
    app.set_wfpt_params(spp=4, max_depth=3, rr=0.5, nee=False, enable_clamp=True, clamp_max=10.0)
    p = app.wfpt_params()
    assert (p["spp"], p["max_depth"], p["nee"], p["enable_clamp"], p["clamp_max"]) == (4, 3, False, True, 10.0)
    lib = krr.load_wfpt()
    h = C.c_void_p()
    assert lib.krr_wfpt_create(b'{"max_depth": -1}', C.byref(h)) == -1
    assert lib.krr_wfpt_create(b'{"rr": 0}', C.byref(h)) == -1
    assert lib.krr_wfpt_create(b'{"spp": 0}', C.byref(h)) == -1
    assert lib.krr_wfpt_create(b"not json", C.byref(h)) == -1
    assert b"JSON" in lib.krr_wfpt_last_error()


def test_bad_configs_return_errors():
    with pytest.raises(RuntimeError):
        krr.HostApp("/nonexistent/config.json")
    with pytest.raises(RuntimeError):
        krr.HostApp({"passes": [{"name": "WavefrontPathTracer"}], "scene": {"model": [{"model": "missing.obj"}]}}, asset_root=ROOT)
    with pytest.raises(RuntimeError):  # unknown pass names are fatal in the reference (renderpass.h:216-220)
        krr.HostApp({"passes": [{"name": "NoSuchPass", "enable": True}], "scene": json.load(open(CBOX))["scene"]}, asset_root=ROOT)


def test_inline_scene_with_node_transform():
    cfg = json.load(open(CBOX))
    cfg["scene"]["model"][0].update({"translate": [1.0, 2.0, 3.0], "scale": [2.0, 2.0, 2.0]})
    app_a, app_b = krr.HostApp(cfg, asset_root=ROOT), krr.HostApp(CBOX, asset_root=ROOT)  # keep alive: the descs point into them
    a, b = app_a.scene_desc().contents, app_b.scene_desc().contents
    ta = np.array(list(a.instances[0].transform)).reshape(3, 4)
    tb = np.array(list(b.instances[0].transform)).reshape(3, 4)
    assert np.allclose(ta[:, :3], 2 * tb[:, :3]) and np.allclose(ta[:, 3], 2 * tb[:, 3] + [1, 2, 3])


def test_grid_medium_with_an_albedo_gradient():
    """{"type": "grid", ..., "albedo_gradient": [[r,g,b],[r,g,b]]}: the host layer hands the pass a density grid AND an RGB
    albedo grid on the same lattice (KrrMediumDesc::albedo_grid = NanoVDBMedium::albedoGrid, media.h:168-170)."""
    plain = krr.HostApp(os.path.join(ROOT, "assets", "configs", "cbox_smoke.json"), asset_root=ROOT)
    app = krr.HostApp(os.path.join(ROOT, "assets", "configs", "cbox_smoke_albedo.json"), asset_root=ROOT)
    d0, d = plain.scene_desc().contents, app.scene_desc().contents
    assert d0.n_media == d.n_media == 1
    assert not d0.media[0].albedo_grid, "no gradient asked for: constant albedo"
    m = d.media[0]
    res = tuple(m.res)
    assert res == (96, 96, 96) and m.density and m.albedo_grid
    n = res[0] * res[1] * res[2]
    a = np.ctypeslib.as_array(m.albedo_grid, shape=(res[2], res[1], res[0], 3))
    assert np.allclose(a[:, :, 0], [0.95, 0.25, 0.2]) and np.allclose(a[:, :, -1], [0.2, 0.35, 0.95])  # x = 0 / x = res - 1
    assert np.allclose(a[5, 7, 48], np.array([0.95, 0.25, 0.2]) + 48 / 95 * (np.array([0.2, 0.35, 0.95]) - [0.95, 0.25, 0.2]), atol=1e-6)
    den = np.ctypeslib.as_array(m.density, shape=(n,))
    den0 = np.ctypeslib.as_array(d0.media[0].density, shape=(n,))
    assert np.array_equal(den, den0), "the density is the same procedural grid"
