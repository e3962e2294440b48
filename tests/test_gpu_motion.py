"""Motion blur over a multi-level instanced scene graph (BASELINE.json config 5):
SRT motion transforms on group AND instance nodes, evaluated per ray at the ray's time; TLAS boxes of
the moving instances re-fitted to the camera's shutter interval every frame.

OptiX (closed) evaluates the reference's motion transforms, so the evaluation is specified by this
build (kiraray_b200/csrc/motion.cuh == oracle/driver.cpp nodeXf/chainXf): the transforms must agree
BIT FOR BIT, and the first-hit ids must equal the oracle's, which brute-forces every moving instance
(no bounds involved), so conservative motion bounds are verified by the same comparison."""
import numpy as np
import pytest

import kiraray_b200 as krr
import oracle_binding as ob
from __graft_entry__ import relmse
from kiraray_b200 import scenes

pytestmark = pytest.mark.gpu
KIND = "reference"


def small_scene(**kw):  # (n_keys, spin_scale, drift_scale: see scenes.instanced_scene)
    return scenes.instanced_scene(n_blas=3, tris_per_blas=300, n_groups=4, per_group=6, motion=True, **kw)


def camera(shutter_open, shutter_time):
    return scenes.look_at_camera((0.5, 3.0, 9.0), (0, 0, 0), 1.0, shutter_open=shutter_open, shutter_time=shutter_time)


def test_instance_transforms_at_ray_time_are_bit_exact():
    b, info = small_scene()
    desc = b.build()
    gpu = krr.Wfpt(params=dict(spp=1, max_depth=1))
    gpu.set_scene(desc)
    orc = ob.Oracle(desc, KIND)
    rng = np.random.Generator(np.random.PCG64(scenes.SEED))
    n = info["n_moving"]
    ids = np.concatenate([rng.integers(0, n, 200), [n, n + 1]]).astype(np.int32)  # + the static floor and light
    times = np.concatenate([rng.uniform(-0.2, 1.2, 196), [0.0, 1.0, 0.5, 0.25], [0.3, 0.7]]).astype(np.float32)
    got = gpu.instance_xf(ids, times)
    for k, (i, t) in enumerate(zip(ids, times)):
        m, inv = orc.instance_xf(int(i), float(t))
        assert np.array_equal(got[k, 0].view(np.uint32), m.view(np.uint32)), (k, i, t)
        assert np.array_equal(got[k, 1].view(np.uint32), inv.view(np.uint32)), (k, i, t)
        if i < n:  # against a float64 restatement of the chain (group node x instance node)
            assert np.allclose(got[k, 0], info["world"](int(i), float(t)), atol=2e-5)
            full = np.vstack([got[k, 0].reshape(3, 4), [0, 0, 0, 1]]) @ np.vstack([got[k, 1].reshape(3, 4), [0, 0, 0, 1]])
            assert np.allclose(full, np.eye(4), atol=1e-4)
    orc.close()


@pytest.mark.parametrize("shutter", [(0.5, 0.05), (0.0, 1.0), (0.9, 0.4)])
def test_motion_blur_first_hits_counts_and_radiance(shutter):
    b, info = small_scene()
    desc = b.build()
    w = h = 80
    cam = camera(*shutter)
    gpu = krr.Wfpt(params=dict(spp=2, max_depth=3))
    gpu.set_scene(desc)
    gpu.resize(w, h)
    gpu.begin_frame(1, cam)
    film = gpu.render_to_host()
    orc = ob.Oracle(desc, KIND)
    ref = orc.render(cam, w, h, frame_index=1, spp=2, max_depth=3, use_bvh=True)
    orc.close()
    inst, prim = gpu.first_hits()
    assert np.array_equal(inst, ref["first_hits"][:, 0]) and np.array_equal(prim, ref["first_hits"][:, 1])
    assert (inst < info["n_moving"]).sum() > 0.1 * w * h, "the moving instances must cover a good part of the frame"
    st, rs = gpu.stats(), ref["stats"]
    assert st["closest_by_depth"][0] == rs["closest_by_depth"][0] == 2 * w * h
    for d in range(1, 3):
        a, c = st["closest_by_depth"][d], rs["closest_by_depth"][d]
        assert abs(a - c) <= max(8, 0.02 * c), (d, a, c)
    assert abs(st["shadow_rays"] - rs["shadow_rays"]) <= 0.02 * rs["shadow_rays"]
    assert np.isfinite(film).all()
    assert relmse(film, ref["film"]) <= 0.1


@pytest.mark.parametrize("n_keys,spin,drift,shutter", [(2, 30.0, 6.0, (0.0, 1.0)), (5, 30.0, 4.0, (0.1, 0.8)), (3, 12.0, 10.0, (0.45, 0.3))])
def test_fast_rotations_stay_inside_the_motion_boxes(n_keys, spin, drift, shutter):
    """The TLAS boxes of moving instances are padded by a proven bound (bvh_build.cu chainMotionBound: speed bound from the
    keys x half the sample spacing), not by a curvature estimate.  Keys that turn an instance by up to half a revolution
    per segment and move it by several of its diameters inside ONE shutter interval: every ray the oracle's exhaustive
    loop over the moving instances (no bounds at all) finds must be found through the boxes, id for id."""
    b, info = small_scene(n_keys=n_keys, spin_scale=spin, drift_scale=drift)
    desc = b.build()
    w = h = 72
    cam = camera(*shutter)
    gpu = krr.Wfpt(params=dict(spp=2, max_depth=2))
    gpu.set_scene(desc)
    gpu.resize(w, h)
    gpu.begin_frame(3, cam)
    gpu.render_to_host()
    orc = ob.Oracle(desc, KIND)
    ref = orc.render(cam, w, h, frame_index=3, spp=2, max_depth=2, use_bvh=True)
    orc.close()
    inst, prim = gpu.first_hits()
    assert np.array_equal(inst, ref["first_hits"][:, 0]) and np.array_equal(prim, ref["first_hits"][:, 1])
    assert (inst < info["n_moving"]).sum() > 0.03 * w * h
    st, rs = gpu.stats(), ref["stats"]
    assert abs(st["closest_by_depth"][1] - rs["closest_by_depth"][1]) <= max(8, 0.02 * rs["closest_by_depth"][1])
    assert abs(st["shadow_rays"] - rs["shadow_rays"]) <= max(8, 0.02 * rs["shadow_rays"])


def test_shutter_window_change_refits_the_tlas():
    """Frames with different shutter intervals through ONE handle: begin_frame re-fits the moving instances'
    boxes to each interval; hits stay equal to a fresh oracle render, and going back reproduces frame 1."""
    b, info = small_scene()
    desc = b.build()
    w = h = 64
    gpu = krr.Wfpt(params=dict(spp=1, max_depth=1))
    gpu.set_scene(desc)
    gpu.resize(w, h)
    orc = ob.Oracle(desc, KIND)
    seen = []
    for shutter in [(0.1, 0.02), (0.8, 0.1), (0.1, 0.02)]:
        cam = camera(*shutter)
        gpu.begin_frame(1, cam)
        gpu.render_to_host()
        inst, prim = gpu.first_hits()
        ref = orc.render(cam, w, h, frame_index=1, spp=1, max_depth=1, use_bvh=True)
        assert np.array_equal(inst, ref["first_hits"][:, 0]) and np.array_equal(prim, ref["first_hits"][:, 1]), shutter
        seen.append((inst.copy(), prim.copy()))
    orc.close()
    assert not np.array_equal(seen[0][0], seen[1][0]), "the scene must look different at another time"
    assert np.array_equal(seen[0][0], seen[2][0]) and np.array_equal(seen[0][1], seen[2][1])


def test_instant_shutter_equals_the_static_scene_at_that_time():
    """shutterTime = 0 at time t: every ray sees the chain at t, i.e. the static scene whose instance
    transforms are the chains evaluated at t (tiny differences only where a silhouette is rounded differently:
    the static path inverts the composed matrix, the motion path composes the node inverses)."""
    t = 0.37
    b, info = small_scene()
    w = h = 64
    cam = camera(t, 0.0)
    gpu = krr.Wfpt(params=dict(spp=1, max_depth=1))
    gpu.set_scene(b.build())
    gpu.resize(w, h)
    gpu.begin_frame(1, cam)
    gpu.render_to_host()
    inst, prim = gpu.first_hits()
    bs, _ = scenes.instanced_scene(n_blas=3, tris_per_blas=300, n_groups=4, per_group=6, motion=False, time=t)
    gs = krr.Wfpt(params=dict(spp=1, max_depth=1))
    gs.set_scene(bs.build())
    gs.resize(w, h)
    gs.begin_frame(1, cam)
    gs.render_to_host()
    inst_s, prim_s = gs.first_hits()
    assert (inst != inst_s).mean() < 0.003 and (prim != prim_s).mean() < 0.003


def test_single_level_motion_keys_on_the_instance():
    """KrrInstanceDesc::motion_keys shorthand: one motion node per instance over [starttime, endtime], 3 keys."""
    rng = np.random.Generator(np.random.PCG64(scenes.SEED))
    b = scenes.SceneBuilder()
    mat = b.add_material(diffuse=(0.6, 0.4, 0.3))
    p, n, idx = scenes.displaced_sphere(16, 12, rng)
    mesh = b.add_mesh(p, idx, n, mat)
    for k in range(5):
        keys = []
        for a in range(3):
            q = np.array([0.1 * k, 0.3 * a, 0.2, 1.0])
            keys.append(np.concatenate([[0.5 + 0.1 * a] * 3, q / np.linalg.norm(q), [-3 + 1.5 * k + 0.4 * a, 0.3 * a * (k - 2), 0.2 * a]]))
        keys = np.array(keys, np.float32)
        b.add_instance(mesh, scenes.srt_to_mat(keys[0]), motion_keys=keys)
    p, n, idx = scenes.quad((-6, -1.2, -6), (0, 0, 12), (12, 0, 0))
    b.add_instance(b.add_mesh(p, idx, n, b.add_material(diffuse=(0.5, 0.5, 0.5))))
    b.add_light(0, color=(1, 1, 1), scale=60.0, transform=scenes.translation((0, 5, 3)))
    b.options.update(motionblur=1, starttime=2.0, endtime=4.0)
    desc = b.build()
    w = h = 64
    cam = scenes.look_at_camera((0, 2, 8), (0, 0, 0), 1.0, shutter_open=2.5, shutter_time=1.0)
    gpu = krr.Wfpt(params=dict(spp=2, max_depth=2))
    gpu.set_scene(desc)
    gpu.resize(w, h)
    gpu.begin_frame(3, cam)
    film = gpu.render_to_host()
    orc = ob.Oracle(desc, KIND)
    ref = orc.render(cam, w, h, frame_index=3, spp=2, max_depth=2, use_bvh=True)
    orc.close()
    inst, prim = gpu.first_hits()
    assert np.array_equal(inst, ref["first_hits"][:, 0]) and np.array_equal(prim, ref["first_hits"][:, 1])
    assert (inst >= 0).sum() > 0 and (inst < 5).sum() > 50
    assert relmse(film, ref["film"]) <= 0.1


def test_motion_keys_are_ignored_without_the_motionblur_option():
    """getMotionKeyframes returns nothing unless enableMotionBlur (optix.cpp:402): same film as the static scene."""
    b, _ = small_scene()
    b.options.update(motionblur=0)
    bs, _ = scenes.instanced_scene(n_blas=3, tris_per_blas=300, n_groups=4, per_group=6, motion=False)
    w = h = 48
    cam = camera(0.5, 0.2)
    films = []
    for sb in (b, bs):
        g = krr.Wfpt(params=dict(spp=1, max_depth=2))
        g.set_scene(sb.build())
        g.resize(w, h)
        g.begin_frame(1, cam)
        films.append(g.render_to_host())
    assert np.array_equal(films[0], films[1])
