"""Frame batch ("frame_batch": F): F consecutive frame indices rendered by the same launches must give exactly the
mean of the F frames rendered one by one (each frame has its own PCG sequence, sampleIndex = frameIndex * spp,
integrator.cpp:217; pixels and frames are independent), and the ray counters must add up."""
import os

import numpy as np
import pytest

import kiraray_b200 as krr
from kiraray_b200 import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(desc, cam, w, h, first, n_frames, batch, **params):
    gpu = krr.Wfpt(params=dict(debug_taps=False, frame_batch=batch, **params))
    gpu.set_scene(desc)
    gpu.resize(w, h)
    films, rays = [], 0
    for f in range(first, first + n_frames, batch):
        gpu.begin_frame(f, cam)
        films.append(gpu.render_to_host().copy())
        st = gpu.stats()
        rays += st["closest_rays"] + st["shadow_rays"]
    return films, rays


def mean_in_frame_order(films):
    acc = films[0][..., :3].copy()
    for f in films[1:]:
        acc = acc + f[..., :3]
    return acc / np.float32(len(films))


@pytest.mark.parametrize("bands", [1, 2])
def test_batch_equals_mean_of_single_frames_cbox(bands):
    app = krr.HostApp(os.path.join(ROOT, "assets", "configs", "cbox.json"), asset_root=ROOT)
    app.set_resolution(96, 64)
    app.set_wfpt_params(spp=2, max_depth=6)
    p = dict(app.wfpt_params(), bands=bands)
    single, rays1 = run(app.scene_desc(), app.camera(), 96, 64, 3, 4, 1, **p)
    batch, rays4 = run(app.scene_desc(), app.camera(), 96, 64, 3, 4, 4, **p)
    assert rays1 == rays4
    want = mean_in_frame_order(single)
    assert np.array_equal(batch[0][..., :3].view(np.uint32), want.view(np.uint32))
    assert (batch[0][..., 3] == 1).all()
    two, _ = run(app.scene_desc(), app.camera(), 96, 64, 3, 4, 2, **p)
    assert np.array_equal(two[1][..., :3].view(np.uint32), mean_in_frame_order(single[2:]).view(np.uint32))


def test_batch_tree_scene_with_tail_and_motion():
    b, info = scenes.instanced_scene(n_blas=3, tris_per_blas=300, n_groups=4, per_group=6, motion=True)
    cam = scenes.look_at_camera((0.5, 3.0, 9.0), (0, 0, 0), 1.0, shutter_open=0.5, shutter_time=0.05)
    desc = b.build()
    for extra in (dict(), dict(tail_depth=2)):
        single, rays1 = run(desc, cam, 80, 80, 1, 3, 1, spp=1, max_depth=4, **extra)
        batch, rays3 = run(desc, cam, 80, 80, 1, 3, 3, spp=1, max_depth=4, **extra)
        assert rays1 == rays3
        assert np.array_equal(batch[0][..., :3].view(np.uint32), mean_in_frame_order(single).view(np.uint32))
