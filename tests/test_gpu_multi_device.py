"""MultiDeviceRenderApp (C++ host layer, one process, one handle + one host thread per rank).
On a one-GPU box the ranks share device 0 (films summed on the host): this is the two-handles-two-threads test of
the C ABI's "one handle per GPU, one host thread each" promise.  With >= 2 GPUs the same checks run through NCCL."""
import os

import numpy as np
import pytest

import kiraray_b200 as krr
from kiraray_b200.binding import MultiDeviceApp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 128, 96


def setup():
    app = krr.HostApp(os.path.join(ROOT, "assets", "configs", "cbox.json"), asset_root=ROOT)
    app.set_resolution(W, H)
    app.set_wfpt_params(spp=2, max_depth=5)
    return app, dict(app.wfpt_params(), debug_taps=False)


def single(app, params, frames):
    gpu = krr.Wfpt(params=params)
    gpu.set_scene(app.scene_desc())
    gpu.resize(W, H)
    out, rays = [], 0
    for f in frames:
        gpu.begin_frame(f, app.camera())
        out.append(gpu.render_to_host().copy())
        st = gpu.stats()
        rays += st["closest_rays"] + st["shadow_rays"]
    return out, rays


def devices(n):
    import torch
    have = torch.cuda.device_count()
    return list(range(n)) if have >= n else [0] * n


@pytest.mark.parametrize("n,tiles", [(2, 2), (2, 1), (4, 2)])
def test_tiles_add_up_and_spp_slices_average(n, tiles):
    app, params = setup()
    slices = n // tiles
    multi = MultiDeviceApp(app.scene_desc(), params, W, H, devices(n), tiles)
    film, ms, rays = multi.render(app.camera(), first_frame=3, steps=2)   # last step: frames 3 + slices .. 3 + 2 * slices - 1
    frames = [3 + slices + s for s in range(slices)]
    ref, ref_rays = single(app, params, frames)
    want = ref[0][..., :3].copy()
    for f in ref[1:]:
        want = want + f[..., :3]
    want = want * np.float32(1.0 / slices)
    assert rays == ref_rays
    if multi.uses_nccl and slices > 1 and n > 2:
        assert np.allclose(film[..., :3], want, rtol=1e-6, atol=1e-7)   # NCCL's reduction order over > 2 ranks is its own
    else:
        assert np.array_equal(film[..., :3].view(np.uint32), want.view(np.uint32))
    assert ms > 0
    multi.close()
