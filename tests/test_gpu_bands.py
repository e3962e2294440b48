"""Frames rendered as several interleaved row BANDS on concurrent streams (kiraray_b200/csrc/api.cu
WaveState) must be indistinguishable from the same frame rendered as one band: pixels are independent
in the reference (private PCG stream and accumulator per pixel, integrator.cpp:213-220), so the film is
bit-identical and the stage counters add up -- on the fused surface schedule, the reference-order
schedule with participating media, motion-blurred instanced scenes, row partitions and odd sizes."""
import os

import numpy as np
import pytest

import kiraray_b200 as krr
from kiraray_b200 import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COUNTS = ("camera_rays", "closest_rays", "shadow_rays", "scatter_items", "hit_light_items", "miss_items",
          "medium_sample_items", "medium_scatter_items")


def render(desc, cam, w, h, bands, frames=(1, 2), partition=None, **params):
    gpu = krr.Wfpt(params=dict(params, debug_taps=False, bands=bands))
    gpu.set_scene(desc)
    gpu.resize(w, h)
    if partition:
        gpu.set_partition(*partition)
    out = []
    for f in frames:
        gpu.begin_frame(f, cam)
        film = gpu.render_to_host().copy()
        st = gpu.stats()
        out.append((film, st))
    return out


def same(a, b):
    for (fa, sa), (fb, sb) in zip(a, b):
        assert np.array_equal(fa.view(np.uint32), fb.view(np.uint32))
        for k in COUNTS:
            assert sa[k] == sb[k], k
        assert list(sa["closest_by_depth"]) == list(sb["closest_by_depth"])
        assert list(sa["shadow_by_depth"]) == list(sb["shadow_by_depth"])


def app_for(cfg, w, h, **kw):
    app = krr.HostApp(os.path.join(ROOT, "assets", "configs", cfg), asset_root=ROOT)
    app.set_resolution(w, h)
    app.set_wfpt_params(**kw)
    return app


@pytest.mark.parametrize("bands", [2, 3, 4])
def test_cornell_box_film_is_bit_identical(bands):
    app = app_for("cbox.json", 200, 117, spp=3, max_depth=6)  # odd row count: bands of unequal size
    one = render(app.scene_desc(), app.camera(), 200, 117, 1, **app.wfpt_params())
    many = render(app.scene_desc(), app.camera(), 200, 117, bands, **app.wfpt_params())
    same(one, many)
    assert many[0][1]["camera_rays"] == 3 * 200 * 117
    assert many[0][1]["kernel_launches"] > one[0][1]["kernel_launches"]


def test_row_partition_with_bands():
    app = app_for("cbox.json", 160, 90, spp=2, max_depth=5)
    one = render(app.scene_desc(), app.camera(), 160, 90, 1, partition=(31, 64), **app.wfpt_params())
    two = render(app.scene_desc(), app.camera(), 160, 90, 2, partition=(31, 64), **app.wfpt_params())
    same(one, two)
    film = two[0][0]
    # film rows are flipped (row H-1-y); rows outside the partition are cleared
    assert not film[: 90 - 64].any() and not film[90 - 31:].any() and film[90 - 64: 90 - 31, :, :3].any()


def test_more_bands_than_rows():
    app = app_for("cbox.json", 64, 48, spp=1, max_depth=3)
    one = render(app.scene_desc(), app.camera(), 64, 48, 1, partition=(20, 22), **app.wfpt_params())
    four = render(app.scene_desc(), app.camera(), 64, 48, 4, partition=(20, 22), **app.wfpt_params())
    same(one, four)


def test_media_schedule_with_bands():
    app = app_for("cbox_smoke.json", 96, 96, spp=4, max_depth=6)
    one = render(app.scene_desc(), app.camera(), 96, 96, 1, **app.wfpt_params())
    two = render(app.scene_desc(), app.camera(), 96, 96, 2, **app.wfpt_params())
    same(one, two)
    assert two[0][1]["medium_sample_items"] > 0


def test_motion_blur_with_bands():
    b, info = scenes.instanced_scene(n_blas=3, tris_per_blas=300, n_groups=4, per_group=6, motion=True)
    desc = b.build()
    cam = scenes.look_at_camera((0.5, 3.0, 9.0), (0, 0, 0), 1.0, shutter_open=0.5, shutter_time=0.05)
    one = render(desc, cam, 96, 80, 1, spp=2, max_depth=3)
    two = render(desc, cam, 96, 80, 2, spp=2, max_depth=3)
    same(one, two)


def test_taps_and_profiling_fall_back_to_one_band():
    app = app_for("cbox.json", 64, 64, spp=1, max_depth=3)
    gpu = krr.Wfpt(params=dict(app.wfpt_params(), debug_taps=False, bands=2))
    gpu.set_scene(app.scene_desc())
    gpu.resize(64, 64)
    gpu.begin_frame(1, app.camera())
    film2 = gpu.render_to_host().copy()
    with pytest.raises(Exception):
        gpu.pixel_state()  # the frame's state is split over two bands
    gpu.set_profiling(True)
    gpu.begin_frame(1, app.camera())
    film1 = gpu.render_to_host().copy()
    assert gpu.stage_times()["scatter"]["launches"] > 0
    gpu.set_profiling(False)
    assert np.array_equal(film1.view(np.uint32), film2.view(np.uint32))


def test_pipelined_readback_delivers_the_same_frames():
    app = app_for("cbox.json", 120, 68, spp=2, max_depth=4)
    gpu = krr.Wfpt(params=dict(app.wfpt_params(), debug_taps=False))
    gpu.set_scene(app.scene_desc())
    gpu.resize(120, 68)
    want = []
    for f in range(1, 6):
        gpu.begin_frame(f, app.camera())
        want.append(gpu.render_to_host().copy())
    got = [np.zeros((68, 120, 4), np.float32) for _ in want]
    for f in range(1, 6):  # five frames in flight over two internal device films, no host sync in between
        gpu.begin_frame(f, app.camera())
        gpu.render_to_host_async(got[f - 1])
    gpu.wait_host()
    for a, b in zip(want, got):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert not np.array_equal(want[0], want[1])
