"""The tail kernel (k_tail: every remaining bounce of a path in one launch from loop depth `tail_depth` on) against
the per-depth stage launches it replaces: the film must be BIT-IDENTICAL (per pixel the order of the random draws
and of the additions to L is the reference's depth loop, integrator.cpp:232-256) and every per-depth counter equal."""
import os

import numpy as np
import pytest

import kiraray_b200 as krr
from kiraray_b200 import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def render(desc, cam, w, h, **params):
    gpu = krr.Wfpt(params=dict(debug_taps=False, **params))
    gpu.set_scene(desc)
    gpu.resize(w, h)
    films = []
    for f in (1, 2):
        gpu.begin_frame(f, cam)
        films.append(gpu.render_to_host().copy())
    return films, gpu.stats()


def check(desc, cam, w, h, tails, **params):
    ref, rs = render(desc, cam, w, h, tail_depth=0, **params)
    for t in tails:
        got, st = render(desc, cam, w, h, tail_depth=t, **params)
        for a, b in zip(got, ref):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"tail_depth={t}: film differs in {int((a != b).any(-1).sum())} pixels"
        for k in ("closest_rays", "shadow_rays", "scatter_items", "hit_light_items", "miss_items"):
            assert st[k] == rs[k], (t, k, st[k], rs[k])
        assert list(st["closest_by_depth"]) == list(rs["closest_by_depth"]) and list(st["shadow_by_depth"]) == list(rs["shadow_by_depth"]), t
        assert st["kernel_launches"] <= rs["kernel_launches"]


def test_tail_cornell_box_flat_list():
    app = krr.HostApp(os.path.join(ROOT, "assets", "configs", "cbox.json"), asset_root=ROOT)
    app.set_resolution(160, 120)
    app.set_wfpt_params(spp=3, max_depth=10)
    check(app.scene_desc(), app.camera(), 160, 120, (1, 2, 5, 10), **dict(app.wfpt_params()))


def test_tail_tree_scene_all_material_types():
    """merged / flattened tree + TLAS instances, Disney + diffuse + dielectric + conductor materials, a null-material
    surface the rays pass through, an environment light (miss items carry radiance)"""
    rng = np.random.Generator(np.random.PCG64(scenes.SEED))
    b = scenes.SceneBuilder()
    kinds = [dict(diffuse=(0.7, 0.4, 0.3), bsdf_type=1), dict(diffuse=(1, 1, 1), roughness=0.0, bsdf_type=2, ior=1.5),
             dict(diffuse=(0.9, 0.7, 0.3), roughness=0.3, bsdf_type=3, ior=0.4), dict(diffuse=(0.3, 0.5, 0.8), roughness=0.5, bsdf_type=4)]
    sph = scenes.displaced_sphere(40, 24, rng, amplitude=0.1)
    shared = b.add_mesh(sph[0], sph[2], sph[1], b.add_material(**kinds[3]))
    for k in range(6):
        m = b.add_material(**kinds[k % 4])
        s = scenes.displaced_sphere(32, 20, rng, amplitude=0.1)
        b.add_instance(b.add_mesh(s[0], s[2], s[1], m), scenes.translation((-2.5 + k, 0.1 * k, 0.3 * (k % 2)), 0.45))  # single use: flattened
        b.add_instance(shared, scenes.translation((-2.5 + k, 1.3, -0.8), 0.4))                                        # shared mesh: TLAS instance
    p, n, idx = scenes.quad((-6, -0.6, -6), (0, 0, 12), (12, 0, 0))
    b.add_instance(b.add_mesh(p, idx, n, b.add_material(diffuse=(0.5, 0.5, 0.5), bsdf_type=1)))
    p, n, idx = scenes.quad((-3, -0.6, 1.2), (6, 0, 0), (0, 3, 0))
    b.add_instance(b.add_mesh(p, idx, n, material=-1))  # null-material sheet in front of the camera
    p, n, idx = scenes.quad((-1, 3.0, -1), (2, 0, 0), (0, 0, 2))
    b.add_instance(b.add_mesh(p, idx, n, b.add_material(diffuse=(0, 0, 0), emissive=(17, 12, 4))))
    b.add_light(4, color=(0.4, 0.5, 0.7), scale=1.0, scene_radius=12.0)
    cam = scenes.look_at_camera((0, 1.0, 6.5), (0, 0.3, 0), 4 / 3)
    check(b.build(), cam, 128, 96, (1, 3, 6), spp=2, max_depth=6)


def test_tail_moving_instances():
    b, info = scenes.instanced_scene(n_blas=3, tris_per_blas=300, n_groups=4, per_group=6, motion=True)
    cam = scenes.look_at_camera((0.5, 3.0, 9.0), (0, 0, 0), 1.0, shutter_open=0.5, shutter_time=0.05)
    check(b.build(), cam, 96, 96, (2, 4), spp=2, max_depth=5)
