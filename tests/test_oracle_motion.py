"""CPU checks of the oracle's motion-blur restatement (transform chains with SRT keys, oracle/driver.cpp
nodeXf/chainXf; reference call sites src/core/device/optix.cpp:400-563, shading.h:70-76): the chain
evaluation against an independent float64 restatement, clamping outside the key range, and the render
loop against static scenes posed at the same time."""
import numpy as np

import oracle_binding as ob
from kiraray_b200 import scenes

KIND = "reference"


def small_scene(motion=True, **kw):
    return scenes.instanced_scene(n_blas=2, tris_per_blas=120, n_groups=3, per_group=4, motion=motion, **kw)


def test_chain_transform_matches_float64_restatement_and_clamps():
    b, info = small_scene()
    orc = ob.Oracle(b.build(), KIND)
    n = info["n_moving"]
    for i in range(n):
        for t in (0.0, 0.13, 0.5, 0.77, 1.0):
            m, inv = orc.instance_xf(i, t)
            assert np.allclose(m, info["world"](i, t), atol=2e-5), (i, t)
            full = np.vstack([m.reshape(3, 4), [0, 0, 0, 1]]) @ np.vstack([inv.reshape(3, 4), [0, 0, 0, 1]])
            assert np.allclose(full, np.eye(4), atol=1e-4)
        # OptiX clamps the motion outside [timeBegin, timeEnd]
        assert np.array_equal(orc.instance_xf(i, -3.0)[0], orc.instance_xf(i, 0.0)[0])
        assert np.array_equal(orc.instance_xf(i, 7.0)[0], orc.instance_xf(i, 1.0)[0])
    # static instances (floor, light) keep the uploaded matrix at any time
    m, _ = orc.instance_xf(n, 0.4)
    assert np.array_equal(m, np.array(b.instances[n].transform, np.float32))
    orc.close()


def test_instant_shutter_equals_static_pose_and_wide_shutter_blurs():
    w = h = 40
    t = 0.6
    b, info = small_scene()
    orc = ob.Oracle(b.build(), KIND)
    cam = scenes.look_at_camera((0.5, 3.0, 9.0), (0, 0, 0), 1.0, shutter_open=t, shutter_time=0.0)
    moving = orc.render(cam, w, h, spp=1, max_depth=2)
    bs, _ = small_scene(motion=False, time=t)
    ors = ob.Oracle(bs.build(), KIND)
    static = ors.render(cam, w, h, spp=1, max_depth=2)
    assert (moving["first_hits"] != static["first_hits"]).mean() < 0.004
    assert moving["stats"]["closest_by_depth"][0] == w * h
    # a wide shutter mixes poses: the hits differ from every instantaneous pose
    wide = orc.render(scenes.look_at_camera((0.5, 3.0, 9.0), (0, 0, 0), 1.0, shutter_open=0.0, shutter_time=1.0), w, h, spp=1, max_depth=2)
    assert (wide["first_hits"] != moving["first_hits"]).mean() > 0.02
    # brute force and the BVH path (moving instances are always brute-forced) agree exactly
    brute = orc.render(cam, w, h, spp=1, max_depth=2, use_bvh=False)
    assert np.array_equal(brute["first_hits"], moving["first_hits"])
    orc.close(), ors.close()


def test_reciprocal_srt_evaluation_stays_within_ulps_of_the_division_form():
    """The kernels and the oracle evaluate an SRT node with one reciprocal + products (motion.cuh srtNodeXf, driver.cpp
    nodeXf) instead of the original per-entry divisions (driver.cpp nodeXfDiv, kept as a second statement).  Both
    round the same real number: a product with a correctly rounded reciprocal is within 1.5 ulp of the correctly
    rounded quotient, so chain entries may differ by a few ulp of the largest entry involved, never more; transformed
    points stay within 1e-6 relative."""
    b, info = scenes.instanced_scene(n_blas=2, tris_per_blas=120, n_groups=6, per_group=8, motion=True)
    orc = ob.Oracle(b.build(), KIND)
    rng = np.random.Generator(np.random.PCG64(scenes.SEED))
    worst = 0.0
    for i in range(info["n_moving"]):
        for t in rng.uniform(-0.1, 1.1, 6):
            m, inv = orc.instance_xf(i, float(t))
            md, invd = orc.instance_xf(i, float(t), division_form=True)
            for a, d in ((m, md), (inv, invd)):
                A, D = a.reshape(3, 4), d.reshape(3, 4)
                # linear part: ulp of the largest entry of the row (entries of a row are sums of products of that size)
                scale = np.abs(D[:, :3]).max(axis=1, keepdims=True)
                ulps = np.abs(A[:, :3] - D[:, :3]) / (np.spacing(scale.astype(np.float32)))
                worst = max(worst, float(ulps.max()))
                assert ulps.max() <= 16, (i, t, ulps.max())
                # translation column: relative to the size of the terms that are summed into it
                tscale = np.float32(np.abs(D[:, :3]).max() * 12.0 + np.abs(D[:, 3]).max())
                assert np.abs(A[:, 3] - D[:, 3]).max() <= 16 * np.spacing(tscale), (i, t)
            p = rng.uniform(-1, 1, 3).astype(np.float32)
            w1 = m.reshape(3, 4)[:, :3] @ p + m.reshape(3, 4)[:, 3]
            w2 = md.reshape(3, 4)[:, :3] @ p + md.reshape(3, 4)[:, 3]
            assert np.abs(w1 - w2).max() <= 1e-6 * max(1.0, float(np.abs(w2).max()))
    assert worst > 0, "the two forms are expected to differ in the last bits somewhere (else this test pins nothing)"
    orc.close()
