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


def test_moving_instance_tlas_returns_the_exhaustive_result():
    """use_bvh = 2 walks a BVH over the moving instances' boxes for the render's ray-time window -- sample positions
    padded by the proven speed bound (driver.cpp chainSpeedBoundHost, the host restatement of bvh_build.cu
    chainMotionBound) -- instead of visiting every moving instance.  Same rays, same hits, same film: on gentle keys,
    on keys that spin the instances by up to half a turn per segment and throw them several diameters inside one
    shutter interval, with 2 and with 5 keys per node."""
    import time
    w = h = 56
    for kw, shutter in ((dict(), (0.5, 0.05)), (dict(n_keys=2, spin_scale=30.0, drift_scale=6.0), (0.0, 1.0)),
                        (dict(n_keys=5, spin_scale=30.0, drift_scale=4.0), (0.1, 0.8)), (dict(n_keys=3, spin_scale=12.0, drift_scale=10.0), (0.45, 0.3))):
        b, info = scenes.instanced_scene(n_blas=2, tris_per_blas=120, n_groups=4, per_group=8, motion=True, **kw)
        orc = ob.Oracle(b.build(), KIND)
        cam = scenes.look_at_camera((0.5, 3.0, 9.0), (0, 0, 0), 1.0, shutter_open=shutter[0], shutter_time=shutter[1])
        t0 = time.time()
        full = orc.render(cam, w, h, frame_index=2, spp=2, max_depth=3, use_bvh=1)
        t1 = time.time()
        fast = orc.render(cam, w, h, frame_index=2, spp=2, max_depth=3, use_bvh=2)
        t2 = time.time()
        assert (full["first_hits"][:, 0] < info["n_moving"]).sum() > 0.02 * w * h, "moving instances must be in view"
        assert np.array_equal(full["first_hits"], fast["first_hits"]), (kw, shutter)
        assert np.array_equal(full["film"].view(np.uint32), fast["film"].view(np.uint32)), (kw, shutter)
        assert full["stats"] == fast["stats"]
        orc.close()
        print(f"moving TLAS {kw} shutter {shutter}: exhaustive {t1 - t0:.2f} s, bounded {t2 - t1:.2f} s")
