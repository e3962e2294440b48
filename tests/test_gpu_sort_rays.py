"""Ray reordering ("sort_rays", k_sort_rays): the trace stage of depth >= 1 takes its rays in (direction octant, origin
Morton code) order inside tiles of the queue.  Pixels are independent in the reference (private PCG stream and
accumulator, integrator.cpp:213-220) and hold one ray per depth, so the order a stage walks its queue in changes no
draw and no addition: films must be bit-identical and the ray counters equal, whatever the key and whichever queues
are sorted."""
import numpy as np
import pytest

import kiraray_b200 as krr
from kiraray_b200 import scenes

pytestmark = pytest.mark.gpu


def render(desc, cam, w, h, **params):
    gpu = krr.Wfpt(params=dict(debug_taps=False, **params))
    gpu.set_scene(desc)
    gpu.resize(w, h)
    gpu.begin_frame(2, cam)
    film = gpu.render_to_host().copy()
    st = gpu.stats()
    return film, (st["closest_rays"], st["shadow_rays"], tuple(st["closest_by_depth"]))


@pytest.mark.parametrize("motion", [False, True])
def test_sorted_trace_gives_the_identical_film(motion):
    b, _ = scenes.instanced_scene(n_blas=3, tris_per_blas=400, n_groups=4, per_group=8, motion=motion)
    cam = scenes.look_at_camera((0.5, 3.0, 9.0), (0, 0, 0), 1.0, shutter_open=0.5, shutter_time=0.05 if motion else 0.0)
    desc = b.build()
    base = dict(spp=2, max_depth=5, frame_batch=2)
    ref, rays = render(desc, cam, 150, 110, sort_rays=0, **base)   # 2 x 16 500 rays: several sort tiles, a ragged last one
    assert np.isfinite(ref).all() and ref[..., :3].mean() > 0
    for extra in (dict(sort_rays=3), dict(sort_rays=1), dict(sort_rays=2), dict(sort_rays=3, sort_key=1), dict(sort_rays=-1),
                  dict(sort_rays=3, bands=2), dict(sort_rays=3, tail_depth=3), dict(sort_rays=7), dict(sort_rays=5), dict(sort_rays=0, l2_persist_mb=4)):
        film, r = render(desc, cam, 150, 110, **base, **extra)
        assert r == rays, extra
        assert np.array_equal(film.view(np.uint32), ref.view(np.uint32)), extra


def test_sorted_trace_on_a_merged_static_tree():
    b = scenes.tessellated_scene(n_objects=6, tris_per_object=2500, n_emissive=40)  # (the builder owns the arrays desc points to)
    desc = b.build()
    cam = scenes.look_at_camera((0.4, 0.5, 3.4), (0, -0.1, 0), 1.5)
    ref, rays = render(desc, cam, 120, 80, sort_rays=0, spp=1, max_depth=6)
    film, r = render(desc, cam, 120, 80, sort_rays=3, spp=1, max_depth=6)
    assert r == rays and np.array_equal(film.view(np.uint32), ref.view(np.uint32))
